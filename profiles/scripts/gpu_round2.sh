# second, short round after the default batch moved to 2368: the two new size tests, the default bench, one ncu pass
O=gpurun_out; T=${1:-r01p}
timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_scan_gpu.py -m gpu -x -q -k "batch_2368 or 18944" > $O/${T}_tests.log 2>&1; echo EXIT=$? >> $O/${T}_tests.log
timeout 400 python bench.py > $O/${T}_bench.log 2>$O/${T}_bench.err; echo EXIT=$? >> $O/${T}_bench.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${T}_ncu_qstep.csv python tests/gpu_prof_qstep.py 2368 1 > $O/${T}_ncu_qstep_run.log 2>&1
tail -3 $O/${T}_tests.log; tail -c 600 $O/${T}_bench.log; tail -2 $O/${T}_ncu_qstep_run.log
