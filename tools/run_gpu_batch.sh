mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== bench (default)"; timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; tail -n 3 gpurun_out/bench12.err; python -c "
import json; d=json.load(open('gpurun_out/bench12.json')); print('default', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()}, d['cpu_baseline'])"
echo "=== wgrad 1cta debug3"; MAPNET_TC_2CTA=0 MAPNET_TC_DEBUG=3 timeout 100 python tools/bench_conv.py 64 2>&1 | tail -7 | cut -c60-100
echo "=== wgrad 1cta"; MAPNET_TC_2CTA=0 timeout 100 python tools/bench_conv.py 64 2>&1 | tail -7 | cut -c60-100
echo "=== default microbench"; timeout 100 python tools/bench_conv.py 64 2>&1 | tail -7 | cut -c1-100
