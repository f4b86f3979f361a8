mkdir -p gpurun_out
echo "=== gpu tests (s2d stem)"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for sd in 1 0; do
echo "=== bench STEM_S2D=$sd"; MAPNET_STEM_S2D=$sd timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench28_$sd.json 2> gpurun_out/bench28_$sd.err; tail -n 3 gpurun_out/bench28_$sd.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench28_$sd.json')); r=d['roofline']
print('s2d$sd', d['value'], d['ms_per_step'], 'eager', d['config'].get('eager_ms_per_step'), 'e2e', d['e2e']['value'], d['gpu_launches'], r['conv_ms_per_step'], {k:round(v['tflops']) for k,v in r['per_class'].items()})
PY
done
