# r02n: ncu evidence of round 2: one-pass time + DRAM bytes per launch of a 2368-query step, the launch list of bench.py itself,
# --set full of the cross-attention / ragged self-attention / LayerNorm kernels inside that step
O=gpurun_out; T=${1:-r02n}; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/${T}_ncu_qstep.csv python tests/gpu_prof_qstep.py 2368 1 > $O/${T}_ncu_qstep_run.log 2>&1
tail -2 $O/${T}_ncu_qstep_run.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --index-images 128 --no-cpu-baseline --no-vitg --no-eager-gpu --no-rerank > $O/${T}_ncu_launches_run.log 2>&1
tail -c 300 $O/${T}_ncu_launches_run.log
for k in qf_cross_attention_v2 qf_ragged layernorm_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/${T}_full_$k \
    python tests/gpu_prof_qstep.py 2368 1 > $O/${T}_full_${k}_run.log 2>&1
  tail -1 $O/${T}_full_${k}_run.log
done
ls -la $O | grep ${T}
