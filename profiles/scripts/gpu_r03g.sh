# r03g: 514-key (rerank) cross-attention with K / V halves on separate barriers: op tests, parity, rerank timing
O=gpurun_out; T=${1:-r03g}; mkdir -p $O
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "cross_attention" > $O/${T}_cross_tests.log 2>&1; echo EXIT=$? >> $O/${T}_cross_tests.log
grep -E "passed|failed|EXIT|Error|assert" $O/${T}_cross_tests.log | tail -4
SPRC_CROSS_ATTN_1G=1 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "cross_attention_257" > $O/${T}_cross_tests_1g.log 2>&1; echo EXIT=$? >> $O/${T}_cross_tests_1g.log
grep -E "passed|failed|EXIT" $O/${T}_cross_tests_1g.log | tail -2
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin_gpu.py -m gpu -q -x > $O/${T}_parity.log 2>&1; echo EXIT=$? >> $O/${T}_parity.log
grep -E "passed|failed|EXIT" $O/${T}_parity.log | tail -3
timeout 300 python tests/gpu_bench_rerank.py 8 100 > $O/${T}_rerank.log 2>&1; mv $O/rerank_shapes.csv $O/${T}_rerank_shapes.csv; tail -1 $O/${T}_rerank.log
python tools/show_profile.py $O/${T}_rerank_shapes.csv 3 2>/dev/null | grep -E "cross|total"
