mkdir -p gpurun_out
echo "=== stem unit test"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stem" 2>&1 | tail -15
echo "=== tc vs simt"; timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -s -k "tensor_core_path_matches" 2>&1 | grep -E "tc-vs-simt|assert|passed|failed|Error" | cut -c1-600
echo "=== tc vs simt, im2col stem"; MAPNET_STEM_S2D=0 timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -s -k "tensor_core_path_matches" 2>&1 | grep -E "tc-vs-simt|assert|passed|failed|Error" | cut -c1-600
echo "=== all gpu tests (no -x)"; timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -6
