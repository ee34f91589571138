# r03q (2 GPUs): retrieval.query_topk over NCCL with the grouped scan writing the exchange buffer (dist test)
O=gpurun_out; T=${1:-r03q}; mkdir -p $O
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -s > $O/${T}_dist_tests.log 2>&1; echo EXIT=$? >> $O/${T}_dist_tests.log
grep -E "rows_equal|passed|failed|EXIT|skipped|Error" $O/${T}_dist_tests.log | tail -5
