# Experiment: L2 policy of the GEMV's dictionary loads on the c2s workload (single-signal omp, 1024 x 8192 dictionary).
#   CSB200_GEMV_POLICY = none   loads without an L2::cache_hint descriptor
#                        normal createpolicy ... L2::evict_normal   (the default of the L2-regime instantiation)
#                        last   createpolicy ... L2::evict_last
#                        first  createpolicy ... L2::evict_first
# Run on the GPU box:  bash tools/keep_sweep.sh
# The committed profiles/c2s_keep_{-1,0,1}_*.csv were written by the first version of this script, whose hook took
# -1 / 0 / 1 for what is now none / normal / last (and fractions in between for "evict_last on a fraction of the lines,
# evict_first on the rest", which never beat evict_normal and was removed).
mkdir -p gpurun_out
for K in none normal last none normal first; do
  echo "POLICY=$K"
  CSB200_GEMV_POLICY=$K timeout 120 python tools/bench_configs.py --config c2s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(d[k]['us_per_solve'],1), round(d[k]['gemv_us_per_launch'],2)) for k in ('f64','f32')})"
done
for K in none normal last; do
  for PH in "f64 100" "f32 2500"; do
    set -- $PH
    CSB200_GRAPH=0 CSB200_GEMV_POLICY=$K timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -k regex:corr_gemv -s $2 -c 12 --csv --log-file gpurun_out/policy_${K}_$1.csv python tools/bench_configs.py --config c2s > /dev/null 2>&1
    echo "POLICY=$K $1 gemv ns:" $(grep corr_gemv gpurun_out/policy_${K}_$1.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')
  done
done
CSB200_GRAPH=0 CSB200_GEMV_POLICY=normal timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:corr_gemv -s 100 -c 4 --csv --log-file gpurun_out/policy_normal_hit.csv python tools/bench_configs.py --config c2s > /dev/null 2>&1
grep corr_gemv gpurun_out/policy_normal_hit.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
