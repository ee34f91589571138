# r02y: why is the fold consumer GEMM slower than the default GEMM on the same shape?  ncu --set full of both
O=gpurun_out; T=${1:-r02y}; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 2 -f -o $O/${T}_consumer_vs_default \
  python tests/gpu_prof_fold.py 112184 2304 768 > $O/${T}_ncu.log 2>&1
tail -3 $O/${T}_ncu.log
