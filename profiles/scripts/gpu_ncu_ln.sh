#!/bin/bash
O=gpurun_out; mkdir -p $O; T=ncu_ln
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ln768 -s 6 -c 2 -f -o $O/${T} python tests/gpu_prof_qstep.py 592 1 > $O/${T}_run.log 2>&1
ncu -i $O/${T}.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
ncu -i $O/${T}.ncu-rep --page details > $O/${T}_details.txt 2>/dev/null
ncu -i $O/${T}.ncu-rep --page source --csv --print-source sass > $O/${T}_source.csv 2>/dev/null
ls -la $O | tail -5
