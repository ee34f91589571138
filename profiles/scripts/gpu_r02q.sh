# r02q: does sleeping between barrier probes in the GEMM epilogue warps buy clock at the power cap?
O=gpurun_out; T=${1:-r02q}; mkdir -p $O
for b in 0 64 256 1000; do
  echo "== SPRC_GEMM_BACKOFF=$b" >> $O/${T}_backoff.log
  SPRC_GEMM_BACKOFF=$b SPRC_SHAPES="vitL fc1,qf qkv,qf ffn2,kv proj" timeout 600 python tests/gpu_sustained_gemm.py 1.2 >> $O/${T}_backoff.log 2>&1
done
cat $O/${T}_backoff.log
