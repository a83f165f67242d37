#!/bin/bash
# per-kernel times of the update variants (ncu launch list, duration only, under gpurun)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 --fp64-steps 0"
CSB200_UPD_DEFER=4 CSB200_UPD_WARP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'omp_|corr_screen' -s 80 -c 160 --csv --log-file gpurun_out/upd_list_d4w.csv $B > gpurun_out/upd_list_d4w.log 2>&1
ls -la gpurun_out/upd_list*
