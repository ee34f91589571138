# r03p: clean rebuild of the library, full GPU suite + smoke on the final tree
O=gpurun_out; T=${1:-r03p}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo EXIT=$? >> $O/${T}_gpu_tests.log
grep -E "passed|failed|EXIT" $O/${T}_gpu_tests.log | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo EXIT=$? >> $O/${T}_smoke.log; tail -2 $O/${T}_smoke.log
