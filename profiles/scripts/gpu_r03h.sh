# r03h: ncu --set full of the ragged self-attention and the two-group cross-attention inside a 2368-query step
O=gpurun_out; T=${1:-r03h}; mkdir -p $O
for k in qf_self_attention_ragged qf_cross_attention_g2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/${T}_full_$k \
    python tests/gpu_prof_qstep.py 2368 1 > $O/${T}_full_${k}_run.log 2>&1
  tail -1 $O/${T}_full_${k}_run.log
done
ls -la $O | grep ${T}
