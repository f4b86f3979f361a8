mkdir -p gpurun_out
echo "=== conv kernel tests"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
echo "=== bench_conv"; timeout 120 python tools/bench_conv.py 64 2>&1 | tail -7
echo "=== bench_conv WGRAD_OCC=2"; MAPNET_TC_WGRAD_OCC=2 timeout 120 python tools/bench_conv.py 64 2>&1 | tail -7
echo "=== gpu step tests"; timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -3
echo "=== bench fused"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err; tail -n 3 gpurun_out/bench24.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench24.json')); r=d['roofline']
print('fused', d['value'], d['ms_per_step'], 'eager', d['config'].get('eager_ms_per_step'), 'e2e', d['e2e']['value'], d['gpu_launches'], r['conv_ms_per_step'], {k:round(v['tflops']) for k,v in r['per_class'].items()})
PY
echo "=== bench fused occ2"; MAPNET_TC_WGRAD_OCC=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench24o.json 2> gpurun_out/bench24o.err; tail -n 3 gpurun_out/bench24o.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench24o.json')); r=d['roofline']
print('fused occ2', d['value'], d['ms_per_step'], 'eager', d['config'].get('eager_ms_per_step'), 'e2e', d['e2e']['value'], d['gpu_launches'], r['conv_ms_per_step'], {k:round(v['tflops']) for k,v in r['per_class'].items()})
PY
