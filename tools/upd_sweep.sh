#!/bin/bash
# Variants of the screened omp loop at the headline shape (under gpurun, one B200): pass operand format x update variant x schedule.
# profiles/upd_sweep_r02.log was collected with successive versions of this script during round 2 (each "call" there is one gpurun
# invocation of it with that call's set of `run` lines).  The result digest must be the same on every line: all variants are
# bit-identical by construction (tests/test_gpu_screen.py checks the same on smaller shapes).
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
run default
run tf32 CSB200_SCREEN_F16=0
run cta CSB200_UPD_WARP=0 CSB200_UPD_DEFER=0
run cta_deferred CSB200_UPD_WARP=0 CSB200_UPD_DEFER=4
run warp_1slice CSB200_UPD_DEFER=4
run warp_4slices CSB200_UPD_DEFER=1
run ring4 CSB200_UPD_WARP=0 CSB200_UPD_DEFER=0 CSB200_UPD_RING=4 CSB200_UPD_HINTS=3
run parts2 CSB200_SCREEN_PARTS=2
run stages3 CSB200_SCREEN_STAGES=3
run chunks4 CSB200_SCREEN_CHUNKS=4
run pipe_uniform CSB200_PIPE_CHUNK=18944
run fp64_only CSB200_SCREEN=0
(timeout 600 python -m pytest tests/test_gpu_screen.py -m gpu -x -q 2>&1 | tail -3)
