# First GPU call of round 2: validate the LayerNorm fold (csrc/ln_fold.cu, SPRC_LN_FOLD=1) and measure it.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_fold.sh r02a'
# Every step runs under its own timeout (a hang in the new epilogues must not cost a strike); the default path is
# measured in the same box right before the folded one so the two numbers share clocks and power state.
O=gpurun_out; T=${1:-r02a}; mkdir -p $O
SPRC_TEST_LN_FOLD=1 timeout 600 python -m pytest tests/test_ln_fold_gpu.py -m gpu -x -q -s > $O/${T}_fold_tests.log 2>&1; echo EXIT=$? >> $O/${T}_fold_tests.log
tail -5 $O/${T}_fold_tests.log
if grep -q "EXIT=0" $O/${T}_fold_tests.log; then
  timeout 400 python bench.py > $O/${T}_bench_default.log 2>$O/${T}_bench_default.err; echo EXIT=$? >> $O/${T}_bench_default.log
  SPRC_LN_FOLD=1 timeout 400 python bench.py > $O/${T}_bench_fold.log 2>$O/${T}_bench_fold.err; echo EXIT=$? >> $O/${T}_bench_fold.log
  # whole GPU suite with the fold switched on (parity gates of the product path must hold with it)
  SPRC_LN_FOLD=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests_fold.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests_fold.log
  SPRC_LN_FOLD=1 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${T}_ncu_qstep_fold.csv python tests/gpu_prof_qstep.py 2368 1 > $O/${T}_ncu_qstep_fold_run.log 2>&1
  tail -c 700 $O/${T}_bench_default.log; tail -c 700 $O/${T}_bench_fold.log; tail -3 $O/${T}_gpu_tests_fold.log
fi
