# r02i (2 GPUs): NCCL correctness of the sharded retrieval drivers (all-to-all exchange) + bench.py --gpus 2
O=gpurun_out; T=${1:-r02i}; mkdir -p $O
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -x -s > $O/${T}_dist_tests.log 2>&1; echo EXIT=$? >> $O/${T}_dist_tests.log
tail -8 $O/${T}_dist_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${T}_bench_n2.log 2>$O/${T}_bench_n2.err; echo EXIT=$? >> $O/${T}_bench_n2.log
tail -c 2500 $O/${T}_bench_n2.log; tail -5 $O/${T}_bench_n2.err
