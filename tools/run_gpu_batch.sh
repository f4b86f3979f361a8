# One consolidated GPU call: validate HEAD (full -m gpu suite), the default bench line, and the ncu launch list.
# Usage:  gpurun --timeout 900 -- 'bash tools/run_gpu_batch.sh'
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "=== bench (default: posenet_bs64, graph)"
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cat gpurun_out/bench_default.json | cut -c1-1500
el "=== all gpu tests (no -x)"
timeout 480 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 | cut -c1-400
el "=== ncu launch list (2 eager steps)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 330 --csv --log-file gpurun_out/launches_raw.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"
el "=== bench mapnet_n32t3"
timeout 200 python bench.py --workload mapnet_n32t3 --no-cpu-baseline > gpurun_out/bench_mapnet.json 2> gpurun_out/bench_mapnet.err; echo "rc=$?"; cat gpurun_out/bench_mapnet.json | cut -c1-600
el "=== done"
