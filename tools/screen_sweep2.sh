#!/bin/bash
# e2e sweep of the screening path: pipeline chunk sizes (under gpurun)
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --secondary none --cpu-signals 0 --e2e-steps 3 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/sweep_{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{tag:28s} value {d['value']:9.0f}  e2e {d['e2e']['value']:9.0f}  step {d['ms_per_step']:7.2f} ms  pass {r['mean_launch_ms']:.3f} ms x {r['launches']}  share {r['share_of_step']:.3f}  ok {d['check']['support_recovered_frac']} {d['e2e']['bit_identical_to_resident_path']}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/sweep_{tag}.err").read()[-400:])
PY
}
run serial CSB200_SCREEN_PARTS=1
run serial_c9472 CSB200_SCREEN_PARTS=1 CSB200_PIPE_CHUNK=9472
run serial_c16384 CSB200_SCREEN_PARTS=1 CSB200_PIPE_CHUNK=16384
run serial_c32768 CSB200_SCREEN_PARTS=1 CSB200_PIPE_CHUNK=32768
run serial_nopipe CSB200_SCREEN_PARTS=1 CSB200_PIPELINE=0
run parts2_s4_c18944 CSB200_SCREEN_PARTS=2 CSB200_SCREEN_STAGES=4
