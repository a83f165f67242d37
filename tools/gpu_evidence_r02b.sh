#!/bin/bash
# Round-2 final evidence on ONE B200 (under gpurun): full GPU test suite, smoke, the default bench line, the reference arm, ncu launch
# lists of smoke() and of a short bench, one ncu --set full capture each of the FP16 screening pass, the warp-per-signal append and
# the residual sweep.  Outputs under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu_r02_final2.log
tail -3 gpurun_out/pytest_gpu_r02_final2.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) > gpurun_out/smoke_r02_final2.log
cat gpurun_out/smoke_r02_final2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_final2.json 2> gpurun_out/bench_r02_final2.err
tail -c 300 gpurun_out/bench_r02_final2.err
B="python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 --fp64-steps 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02b_smoke.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'omp_|corr_screen|reset_state' -s 100 -c 400 --csv --log-file gpurun_out/launches_r02b_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_screen_tf32 -s 20 -c 1 -o gpurun_out/screen_f16_r02 -f $B > gpurun_out/ncu_screen_f16.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:omp_append_warp -s 20 -c 1 -o gpurun_out/append_warp_r02 -f $B > gpurun_out/ncu_append_warp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:omp_residual_slice -s 40 -c 1 -o gpurun_out/residual_slice_r02 -f $B > gpurun_out/ncu_residual_slice.log 2>&1
ls -la gpurun_out | tail -12
