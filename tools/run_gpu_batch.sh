# One consolidated GPU call (<= 5 min): merged stride-2 dgrad validation + A/B bench + launch list.
# Usage:  gpurun --timeout 600 -- 'bash tools/run_gpu_batch.sh'
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "=== new unit tests (default tile choice)"
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "merged or conv_engines" 2>&1 | tail -8 | cut -c1-300
el "=== same, forced CTA pairs"
MAPNET_TC_2CTA=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "merged or (conv_engines and bf16-)" 2>&1 | tail -6 | cut -c1-300
el "=== same, no CTA pairs"
MAPNET_TC_2CTA=0 timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "merged or (conv_engines and bf16-)" 2>&1 | tail -6 | cut -c1-300
el "=== all gpu tests (no -x)"
timeout 300 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "merged-vs|tc-vs-simt|fused-vs|passed|failed|FAILED|Error|assert" | cut -c1-400
el "=== bench A: merged + fold (default)"
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "rc=$?"; cut -c1-260 gpurun_out/bench_A.json
el "=== bench B: per-class launches, no fold"
MAPNET_TC_DGRAD_MERGE=0 MAPNET_TC_DS_FOLD=0 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "rc=$?"; cut -c1-260 gpurun_out/bench_B.json
el "=== bench C: default + 8 elementwise blocks per SM"
MAPNET_EW_BLOCKS_PER_SM=8 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; echo "rc=$?"; cut -c1-260 gpurun_out/bench_C.json
el "=== ncu launch list (default)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 330 --csv --log-file gpurun_out/launches_raw.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"
el "=== done"
