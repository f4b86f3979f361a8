mkdir -p gpurun_out
echo "=== graph bench (1 GPU)"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_graph.json 2> gpurun_out/bench7_graph.err; tail -n 6 gpurun_out/bench7_graph.err; python -c "
import json; d=json.load(open('gpurun_out/bench7_graph.json')); print('graph', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['conv_ms_per_step'])"
echo "=== 2cta conv unit tests"; MAPNET_TC_2CTA=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_engines and bf16 and not simt" 2>&1 | tail -12
echo "=== 2cta bench"; MAPNET_TC_2CTA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_2cta.json 2> gpurun_out/bench7_2cta.err; tail -n 4 gpurun_out/bench7_2cta.err; python -c "
import json; d=json.load(open('gpurun_out/bench7_2cta.json')); print('2cta', d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== graph parity"; timeout 200 python -m pytest tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -5
