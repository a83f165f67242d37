#!/bin/bash
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_screen.py -m gpu -x -q 2>&1 | tail -2)
for i in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 2 --secondary none --cpu-signals 0 --e2e-steps 3 --fp64-steps 0 > gpurun_out/tail_$i.json 2>/dev/null
python - $i <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/tail_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pass ms", round(d["roofline"]["mean_launch_ms"], 4), "update ms", round(d["roofline_update"]["ms_per_update"], 4), d["check"]["result_digest"], d["clocks"]["sm_mhz"])
PY
done
