#!/bin/bash
# last check of the round at HEAD (under gpurun): full GPU test suite, smoke, default bench line
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu_r02_final5.log
tail -2 gpurun_out/pytest_gpu_r02_final5.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) > gpurun_out/smoke_r02_final5.log
cat gpurun_out/smoke_r02_final5.log
timeout 600 python bench.py > gpurun_out/bench_r02_final5.json 2> gpurun_out/bench_r02_final5.err
tail -c 200 gpurun_out/bench_r02_final5.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_final5.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["check"]["oracle_parity"]["pass"], d["check"]["result_digest"], d["clocks"])
PY
