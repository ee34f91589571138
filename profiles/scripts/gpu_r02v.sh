# r02v: grouped-output scan (tests), LayerNorm-fold probe: correctness + sustained timing of the two GEMM epilogues
O=gpurun_out; T=${1:-r02v}; mkdir -p $O
timeout 900 python -m pytest tests/test_scan_gpu.py -m gpu -q -x > $O/${T}_scan_tests.log 2>&1; echo EXIT=$? >> $O/${T}_scan_tests.log
grep -E "passed|failed|EXIT|Error|assert" $O/${T}_scan_tests.log | tail -5
timeout 900 python tests/gpu_probe_fold.py 1.0 > $O/${T}_fold_probe.log 2>&1; echo EXIT=$? >> $O/${T}_fold_probe.log
tail -14 $O/${T}_fold_probe.log
