mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --cpu-signals 0 --e2e-steps 1"
CSB200_GEMM_L2HINT=0 $B > gpurun_out/bench_hint0.json 2>gpurun_out/bench_hint0.err
CSB200_GEMM_L2HINT=1 $B > gpurun_out/bench_hint1.json 2>gpurun_out/bench_hint1.err
python - <<'PY'
import json
for h in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/bench_hint{h}.json").read().strip().splitlines()[-1])
        print("hint",h,"value",d["value"],"gemm ms",d["roofline"]["mean_launch_ms"],"frac",d["roofline"]["frac"])
    except Exception as e: print("hint",h,"failed",e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 --cpu-signals 0 --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
for h in 0 1; do
CSB200_GEMM_L2HINT=$h ncu --set full --clock-control none --import-source on -k regex:corr_gemm -s 10 -c 2 -f -o gpurun_out/corr_gemm_r01b_hint$h python bench.py --steps 1 --warmup 1 --cpu-signals 0 --e2e-steps 1 > gpurun_out/ncu_gemm_hint$h.log 2>&1
ncu -i gpurun_out/corr_gemm_r01b_hint$h.ncu-rep --page raw --csv > gpurun_out/corr_gemm_r01b_hint$h.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:omp_update -s 50 -c 2 -f -o gpurun_out/omp_update_r01b python bench.py --steps 1 --warmup 1 --cpu-signals 0 --e2e-steps 1 > gpurun_out/ncu_update.log 2>&1
ncu -i gpurun_out/omp_update_r01b.ncu-rep --page raw --csv > gpurun_out/omp_update_r01b.csv 2>/dev/null
ls -la gpurun_out | tail -20
