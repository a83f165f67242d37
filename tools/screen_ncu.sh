#!/bin/bash
# ncu evidence for the screening path (under gpurun, ONE GPU): launch list of a short bench, one --set full capture each of the
# tcgen05 screening pass and of the update kernel that follows it (launch ~20 of the solve: support size ~ 20).
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 --fp64-steps 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_r02_screen.csv $B > gpurun_out/ncu_screen_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:corr_screen_tf32 -s 20 -c 1 -o gpurun_out/screen_r02 -f $B > gpurun_out/ncu_screen.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:omp_update_kernel -s 20 -c 1 -o gpurun_out/update_r02_screen -f $B > gpurun_out/ncu_update.log 2>&1
ls -la gpurun_out | grep -i "ncu-rep\|launches_r02_screen"
