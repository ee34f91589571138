#!/bin/bash
# One gpurun call: GPU tests, default bench (+ reference arm), ncu launch list, ncu --set full captures.
# usage: gpurun --timeout 1500 -- 'bash profiles/scripts/gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1; echo "EXIT=$?" >> $O/${TAG}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "EXIT=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.log 2>$O/${TAG}_bench.err; echo "EXIT=$?" >> $O/${TAG}_bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.log 2>&1; echo "EXIT=$?" >> $O/${TAG}_bench_reference.log
# launch list of the same command (short): cold-cache serialised per-launch times
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --index-images 128 --no-cpu-baseline > $O/${TAG}_ncu_launches_run.log 2>&1
# dram bytes of every GEMM launch of the query steps (roofline.traffic = mean over the launches)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16_tcgen05 --csv --log-file $O/${TAG}_ncu_gemm_traffic.csv \
  python tests/gpu_prof_qstep.py 592 1 > $O/${TAG}_ncu_gemm_traffic_run.log 2>&1
# full captures of the top kernels of one query step (ViT depth 1 model, query path only)
for spec in "gemm:regex:gemm_bf16_tcgen05_2cta:30:4" "scan:regex:scan_topk:0:1" "ln:regex:layernorm:20:2" "attnqf:regex:qf_.*attention_tc:4:2"; do
  IFS=: read name kind pat skip cnt <<< "$spec"
  if [ "$name" = "scan" ]; then
    timeout 300 ncu --set full --clock-control none --import-source on -k $kind:$pat -s $skip -c $cnt -f -o $O/${TAG}_full_$name \
      python bench.py --steps 1 --warmup 3 --index-images 128 --no-cpu-baseline > $O/${TAG}_full_${name}_run.log 2>&1
  else
    timeout 300 ncu --set full --clock-control none --import-source on -k $kind:$pat -s $skip -c $cnt -f -o $O/${TAG}_full_$name \
      python tests/gpu_prof_qstep.py 592 1 > $O/${TAG}_full_${name}_run.log 2>&1
  fi
  if [ -f $O/${TAG}_full_$name.ncu-rep ]; then
    ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_full_${name}_raw.csv 2>/dev/null
    ncu -i $O/${TAG}_full_$name.ncu-rep --page details > $O/${TAG}_full_${name}_details.txt 2>/dev/null
  fi
done
ls -la $O
