# One consolidated GPU call: 8-warp epilogue variant (libmapnet_b200_e8.so) vs the default library.
# Usage:  gpurun --timeout 600 -- 'bash tools/run_gpu_batch.sh'
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "=== default library: all gpu tests"
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300
el "=== e8 library: kernel + step tests"
MAPNET_LIB_VARIANT=e8 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -8 | cut -c1-300
el "=== e8 library, forced CTA pairs: kernel tests"
MAPNET_LIB_VARIANT=e8 MAPNET_TC_2CTA=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "merged or (conv_engines and bf16-)" 2>&1 | tail -3 | cut -c1-300
el "=== bench A: default"
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_A.json
el "=== bench E: e8"
MAPNET_LIB_VARIANT=e8 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_E.json 2> gpurun_out/bench_E.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_E.json
el "=== bench C: default + 8 elementwise blocks per SM"
MAPNET_EW_BLOCKS_PER_SM=8 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_C.json
el "=== bench F: e8 + 8 elementwise blocks per SM"
MAPNET_LIB_VARIANT=e8 MAPNET_EW_BLOCKS_PER_SM=8 timeout 120 python bench.py --no-cpu-baseline > gpurun_out/bench_F.json 2> gpurun_out/bench_F.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_F.json
el "=== bench E2: e8 mapnet_n32t3"
MAPNET_LIB_VARIANT=e8 timeout 120 python bench.py --no-cpu-baseline --workload mapnet_n32t3 > gpurun_out/bench_E2.json 2> gpurun_out/bench_E2.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_E2.json
el "=== conv microbench e8 vs default (layer shapes)"
timeout 100 python tools/bench_conv.py 64 2>&1 | tail -8
MAPNET_LIB_VARIANT=e8 timeout 100 python tools/bench_conv.py 64 2>&1 | tail -8
el "=== done"
