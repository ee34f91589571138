# r02b: full GPU suite on the new defaults (fp16 operands, 1e-3 gates, full-depth pins) + bench with all new rows
O=gpurun_out; T=${1:-r02b}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -s > $O/${T}_gpu_tests.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests.log
grep -E "^\[|passed|failed|EXIT|Error|error" $O/${T}_gpu_tests.log | tail -40
timeout 900 python bench.py > $O/${T}_bench.log 2>$O/${T}_bench.err; echo EXIT=$? >> $O/${T}_bench.log
tail -c 3000 $O/${T}_bench.log; tail -5 $O/${T}_bench.err
