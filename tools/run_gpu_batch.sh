mkdir -p gpurun_out
echo "=== gpu tests"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "=== bench_conv"; timeout 120 python tools/bench_conv.py 64 2>&1 | tail -7
for sf in 1 0; do
echo "=== bench STEM_FUSE=$sf"; MAPNET_STEM_FUSE=$sf timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench26_$sf.json 2> gpurun_out/bench26_$sf.err; tail -n 3 gpurun_out/bench26_$sf.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench26_$sf.json')); r=d['roofline']
print('stemfuse$sf', d['value'], d['ms_per_step'], 'eager', d['config'].get('eager_ms_per_step'), 'e2e', d['e2e']['value'], d['gpu_launches'], r['conv_ms_per_step'], {k:round(v['tflops']) for k,v in r['per_class'].items()})
PY
done
