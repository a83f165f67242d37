#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): full GPU test suite, the default bench line, ncu launch lists of smoke()
# and of a short bench, one ncu --set full capture of the cooperative whole-solve kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_r02k.log
tail -3 gpurun_out/pytest_r02k.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err
tail -c 300 gpurun_out/bench_r02k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_smoke.csv \
    python __graft_entry__.py smoke > gpurun_out/ncu_smoke.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv \
    python bench.py --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:persist_solve -s 3 -c 1 -o gpurun_out/persist_r02 -f \
    python bench.py --config c2s --secondary none --cpu-signals 0 > gpurun_out/ncu_persist.log 2>&1
ls -la gpurun_out | tail -8
