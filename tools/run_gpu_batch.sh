mkdir -p gpurun_out
echo "=== conv unit tests"; timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_engines and bf16 and not simt" 2>&1 | tail -3
for cfg in "HALO=0 2CTA=0" "HALO=1 2CTA=0" "HALO=0 2CTA=1" "HALO=0 2CTA=0 DEBUG=1" "HALO=0 2CTA=0 DEBUG=2"; do
  echo "=== $cfg"; env $(echo $cfg | sed 's/\([A-Z0-9]*\)=/MAPNET_TC_\1=/g') timeout 100 python tools/bench_conv.py 64 2>&1 | tail -7 | cut -c1-100
done
echo "=== bench (default)"; timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -n 3 gpurun_out/bench11.err; python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print('default', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
