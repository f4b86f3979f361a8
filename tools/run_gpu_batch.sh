# Final validation call of the round: full -m gpu suite, smoke(), bench lines, one ncu --set full capture of the
# conv engines, conv microbench with the step's epilogue variants.
# Usage:  gpurun --timeout 420 -- 'bash tools/run_gpu_batch.sh'
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "=== all gpu tests (no -x)"
timeout 240 python -m pytest tests -m gpu -q -s --durations=6 2>&1 | grep -vE "^\s*$|^tests/.*UserWarning|losses.append|Consider using|Docs:" | tail -45 | cut -c1-600
el "=== smoke"
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -4
el "=== bench default (posenet_bs64) with cpu baseline"
timeout 200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_final.json
el "=== bench mapnet_n32t3 / mapnetpp_n16t10"
timeout 100 python bench.py --no-cpu-baseline --workload mapnet_n32t3 > gpurun_out/bench_mapnet.json 2> gpurun_out/bench_mapnet.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_mapnet.json
timeout 100 python bench.py --no-cpu-baseline --workload mapnetpp_n16t10 > gpurun_out/bench_mapnetpp.json 2> gpurun_out/bench_mapnetpp.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_mapnetpp.json
el "=== ncu --set full, conv engines (second eager step: stem + layer1/2 fprop, then a backward slice)"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 107 -c 16 -f -o gpurun_out/conv_full_fwd \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1; echo "rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 150 -c 16 -f -o gpurun_out/conv_full_bwd \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bwd.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
el "=== conv microbench: plain / step epilogues / step epilogues + isolated launches"
timeout 60 python tools/bench_conv.py 64 2>&1 | tail -6
MAPNET_BENCH_EPI=1 timeout 60 python tools/bench_conv.py 64 2>&1 | tail -6
MAPNET_BENCH_EPI=2 timeout 60 python tools/bench_conv.py 64 2>&1 | tail -6
el "=== done"
