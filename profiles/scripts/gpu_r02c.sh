# r02c: per-shape kernel timings of one index batch + query step (library profiler) and of a rerank call
O=gpurun_out; T=${1:-r02c}; mkdir -p $O
timeout 600 python bench.py --profile-dump $O/${T}_prof --no-cpu-baseline --no-vitg --no-rerank --index-images 4096 --steps 5 > $O/${T}_bench_prof.log 2>&1
timeout 300 python tests/gpu_bench_rerank.py 8 100 > $O/${T}_rerank.log 2>&1; mv $O/rerank_shapes.csv $O/${T}_rerank_shapes.csv
tail -3 $O/${T}_rerank.log; ls -la $O
