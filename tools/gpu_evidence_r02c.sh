#!/bin/bash
# Round-2 final evidence at HEAD on ONE B200 (under gpurun): full GPU test suite, smoke, the default bench line, ncu launch lists of
# smoke() and of a short bench, one ncu --set full capture of the FP16 screening pass with the 8-warp epilogue.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu_r02_final4.log
tail -3 gpurun_out/pytest_gpu_r02_final4.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) > gpurun_out/smoke_r02_final4.log
cat gpurun_out/smoke_r02_final4.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_final4.json 2> gpurun_out/bench_r02_final4.err
tail -c 300 gpurun_out/bench_r02_final4.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
B="python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 --fp64-steps 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02d_smoke.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'omp_|corr_screen|reset_state' -s 100 -c 400 --csv --log-file gpurun_out/launches_r02d_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_screen_tf32 -s 20 -c 1 -o gpurun_out/screen_f16_r02d -f $B > gpurun_out/ncu_screen_f16d.log 2>&1
ls -la gpurun_out | tail -8
