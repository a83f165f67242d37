#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): full GPU test suite, smoke, the default bench line, ncu launch lists of
# smoke() and of a short bench, one ncu --set full capture each of the correlation GEMM and of the cooperative whole-solve
# kernel, the FP64 / shared-memory latency microbenchmark.  Outputs under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu_r02_final.log
tail -3 gpurun_out/pytest_gpu_r02_final.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) > gpurun_out/smoke_r02_final.log
cat gpurun_out/smoke_r02_final.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
tail -c 300 gpurun_out/bench_r02_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_smoke.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv \
    python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:corr_gemm_f64 -s 6 -c 1 -o gpurun_out/gemm_r02_final -f \
    python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:persist_solve -s 3 -c 1 -o gpurun_out/persist_r02_final -f \
    python tools/persist_timeline.py 1 > gpurun_out/ncu_persist.log 2>&1
./tools/fp64_latency > gpurun_out/fp64_latency_r02.json
ls -la gpurun_out | tail -12
