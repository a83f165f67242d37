#!/bin/bash
# Schedule sweep of the TF32 screening path at the headline shape (under gpurun): serial vs overlapped parts, 3 vs 4 stages.
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --secondary none --cpu-signals 0 --e2e-steps 2 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/sweep_{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{tag:28s} value {d['value']:9.0f}  e2e {d['e2e']['value']:9.0f}  step {d['ms_per_step']:7.2f} ms  pass {r['mean_launch_ms']:.3f} ms x {r['launches']}  share {r['share_of_step']:.3f}  ok {d['check']['support_recovered_frac']}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/sweep_{tag}.err").read()[-400:])
PY
}
run serial_s4 CSB200_SCREEN_PARTS=1
run serial_s3 CSB200_SCREEN_PARTS=1 CSB200_SCREEN_STAGES=3
run parts2_s3 CSB200_SCREEN_PARTS=2
run parts2_s4 CSB200_SCREEN_PARTS=2 CSB200_SCREEN_STAGES=4
run parts3_s3 CSB200_SCREEN_PARTS=3
run parts4_s3 CSB200_SCREEN_PARTS=4
run dmma CSB200_SCREEN=0
