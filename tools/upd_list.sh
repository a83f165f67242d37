#!/bin/bash
# per-kernel times of the deferred update variants (ncu launch list, under gpurun)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 --fp64-steps 0"
for v in 4 1; do
CSB200_UPD_DEFER=$v timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex.sum --clock-control none -k regex:'omp_update|omp_residual|corr_screen' -s 70 -c 200 --csv --log-file gpurun_out/upd_list_d$v.csv $B > gpurun_out/upd_list_d$v.log 2>&1
done
ls -la gpurun_out/upd_list*
