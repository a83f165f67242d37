#!/bin/bash
# e2e sweep after the chunk-rule change: pipeline chunk sizes (under gpurun)
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --secondary none --cpu-signals 0 --e2e-steps 3 --fp64-steps 0 > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/sweep_{tag}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{tag:20s} value {d['value']:9.0f}  e2e {d['e2e']['value']:9.0f}  step {d['ms_per_step']:7.2f} ms  pass {r['mean_launch_ms']:.3f} ms x {r['launches']}  frac {r['frac']:.3f} share {r['share_of_step']:.3f} l2 {r.get('l2_to_sm_tbs')} upd {d['roofline_update']['achieved_tbs']:.2f} TB/s  ok {d['check']['support_recovered_frac']} {d['e2e']['bit_identical_to_resident_path']} {d['check']['screening']} live {r.get('library_gemm_live')}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/sweep_{tag}.err").read()[-600:])
PY
}
run default
run c8192 CSB200_PIPE_CHUNK=8192
run c9472 CSB200_PIPE_CHUNK=9472
run c16384 CSB200_PIPE_CHUNK=16384
run chunks2 CSB200_SCREEN_CHUNKS=2
run chunks4 CSB200_SCREEN_CHUNKS=4
