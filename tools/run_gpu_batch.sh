# Round-end validation call: full -m gpu suite, smoke(), the bench lines, the ncu launch list and ONE ncu --set full
# capture of the conv engines.  The .ncu-rep files stay on the box (gpurun brings back at most 64 MiB -- two reports
# of 16 launches were 73 MB and the whole gpurun_out/ of that call was dropped): they are exported to CSV there and
# only the CSV travels; summarise it here with  python tools/ncu_summary.py --csv gpurun_out/conv_full_raw.csv <tag>.
# Usage:  gpurun --timeout 480 -- 'bash tools/run_gpu_batch.sh'
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "=== all gpu tests (no -x)"
timeout 240 python -m pytest tests -m gpu -q -s --durations=6 2>&1 | grep -E "graph-vs-eager|merged-vs|tc-vs-simt|fused-vs|passed|failed|FAILED|Error|^E " | cut -c1-600
el "=== staged kernels (first GPU run of the f2 input pipeline; opt-in test)"
MAPNET_STAGED_TESTS=1 timeout 120 python -m pytest tests/test_gpu_preprocess.py -m gpu -q 2>&1 | tail -5 | cut -c1-400
el "=== smoke"
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -4
el "=== bench default (posenet_bs64) with cpu baseline"
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cat gpurun_out/bench_final.json
el "=== bench mapnet_n32t3 / mapnetpp_n16t10"
timeout 100 python bench.py --no-cpu-baseline --workload mapnet_n32t3 | tee gpurun_out/bench_mapnet.json
timeout 100 python bench.py --no-cpu-baseline --workload mapnetpp_n16t10 | tee gpurun_out/bench_mapnetpp.json
el "=== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 330 --csv --log-file gpurun_out/launches_raw.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"
el "=== ncu --set full, conv engines (second eager step: 12 forward launches, 12 backward launches)"
for part in "fwd 107" "bwd 150"; do
  set -- $part
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s $2 -c 12 -f -o /tmp/conv_full_$1 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1; echo "rc=$?"
  ncu -i /tmp/conv_full_$1.ncu-rep --page raw --csv > gpurun_out/conv_full_$1_raw.csv 2>/dev/null
done
# source-level stall samples of a short slice (3 launches: one report per direction would be tens of MB of CSV)
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 120 -c 3 -f -o /tmp/conv_src \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_src.log 2>&1; echo "rc=$?"
ncu -i /tmp/conv_src.ncu-rep --page source --csv > gpurun_out/conv_source.csv 2>/dev/null    # -> python tools/ncu_stalls.py
ls -la gpurun_out/*.csv; du -sh gpurun_out
el "=== done"
