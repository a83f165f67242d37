#!/bin/bash
# omp_update_kernel variants at the headline shape (under gpurun): cp.async column ring depth x cache hints.
# The result digest must be the same on every line (the variants are bit-identical by construction).
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --secondary none --cpu-signals 0 --e2e-steps 3 --fp64-steps 0 > gpurun_out/upd_$tag.json 2> gpurun_out/upd_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/upd_{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]; u = d["roofline_update"]
    print(f"{tag:14s} value {d['value']:9.0f}  e2e {d['e2e']['value']:9.0f}  step {d['ms_per_step']:7.2f} ms  pass {r['mean_launch_ms']:.3f} ms  update {u['ms_per_update']:.3f} ms = {u['achieved_tbs']:.2f} TB/s  digest {d['check']['result_digest']} ok {d['check']['support_recovered_frac']} {d['e2e']['bit_identical_to_resident_path']}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/upd_{tag}.err").read()[-600:])
PY
}
run def
run def2
