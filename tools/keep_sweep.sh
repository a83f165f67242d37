# Experiment: L2 policy of the L2-regime GEMV loads on the c2s workload (single-signal omp, 1024 x 8192 dictionary).
# CSB200_GEMV_KEEP: -1 = loads without a policy, 0 = evict_normal policy, f > 0 = evict_last on a fraction f (rest evict_first).
# Run on the GPU box:  bash tools/keep_sweep.sh
mkdir -p gpurun_out
for K in -1 0 1 -1 0; do
  echo "KEEP=$K"
  CSB200_GEMV_KEEP=$K timeout 120 python tools/bench_configs.py --config c2s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(d[k]['us_per_solve'],1), round(d[k]['gemv_us_per_launch'],2)) for k in ('f64','f32')})"
done
for K in -1 0 1; do
  for PH in "f64 100" "f32 2500"; do
    set -- $PH
    CSB200_GRAPH=0 CSB200_GEMV_KEEP=$K timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -k regex:corr_gemv -s $2 -c 12 --csv --log-file gpurun_out/keep_${K}_$1.csv python tools/bench_configs.py --config c2s > /dev/null 2>&1
    echo "KEEP=$K $1 gemv ns:" $(grep corr_gemv gpurun_out/keep_${K}_$1.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')
  done
done
CSB200_GRAPH=0 CSB200_GEMV_KEEP=0 timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:corr_gemv -s 100 -c 4 --csv --log-file gpurun_out/keep_0_hit.csv python tools/bench_configs.py --config c2s > /dev/null 2>&1
grep corr_gemv gpurun_out/keep_0_hit.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
