#!/bin/bash
# fixed cost per update! of the screened omp loop at strong-scaling batch sizes (under gpurun, one GPU)
mkdir -p gpurun_out
for n in 8192 16384 32768; do
  timeout 300 python bench.py --signals $n --steps 5 --warmup 3 --secondary none --cpu-signals 0 --e2e-steps 2 --fp64-steps 0 > gpurun_out/small_$n.json 2> gpurun_out/small_$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/small_{n}.json").read().strip().splitlines()[-1])
print(n, "signals: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "per update us", round(d["ms_per_step"] / 32 * 1e3, 1), "pass us", round(d["roofline"]["mean_launch_ms"] * 1e3, 1))
PY
done
B="python bench.py --signals 8192 --steps 1 --warmup 1 --secondary none --cpu-signals 0 --e2e-steps 1 --fp64-steps 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'omp_|corr_screen' -s 70 -c 40 --csv --log-file gpurun_out/small_8192_list.csv $B > gpurun_out/small_8192_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/small_8192_list.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0].split('::')[-1][:22], float(r[vi].replace(',',''))/1e3) for r in rows[1:]]
print(' | '.join(f"{k} {v:.1f}us" for k,v in seq[:12]))
PY
(timeout 600 python -m pytest tests/test_gpu_screen.py -m gpu -x -q 2>&1 | tail -2)
timeout 300 python bench.py --steps 3 --warmup 2 --secondary none --cpu-signals 0 --e2e-steps 3 --fp64-steps 0 > gpurun_out/small_full.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/small_full.json").read().strip().splitlines()[-1])
print("65536 signals: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pass us", round(d["roofline"]["mean_launch_ms"] * 1e3, 1), d["check"]["result_digest"])
PY
