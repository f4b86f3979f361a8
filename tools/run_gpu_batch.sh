mkdir -p gpurun_out
echo "=== 2cta conv unit tests"; MAPNET_TC_2CTA=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_engines and bf16 and not simt" 2>&1 | tail -4
echo "=== 1cta conv unit tests"; timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_engines and bf16 and not simt" 2>&1 | tail -3
echo "=== 2cta bench"; MAPNET_TC_2CTA=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_2cta.json 2> gpurun_out/bench7_2cta.err; tail -n 4 gpurun_out/bench7_2cta.err; python -c "
import json; d=json.load(open('gpurun_out/bench7_2cta.json')); print('2cta', d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== graph parity"; timeout 200 python -m pytest tests/test_gpu_graph.py -m gpu -q 2>&1 | grep -E "^E|passed|failed" | head -12
echo "=== 2cta step parity"; MAPNET_TC_2CTA=1 timeout 200 python -m pytest tests/test_gpu_step.py -m gpu -q -k "bf16_tensor_core" 2>&1 | tail -4
