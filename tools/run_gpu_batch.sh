mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== bench (default)"; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; tail -n 3 gpurun_out/bench13.err; python -c "
import json; d=json.load(open('gpurun_out/bench13.json')); print('default', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], 'e2e', d['e2e']['value'], d['gpu_launches'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== bench mapnet"; timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload mapnet_n32t3 > gpurun_out/bench13_mapnet.json 2> gpurun_out/bench13m.err; tail -n 3 gpurun_out/bench13m.err; python -c "
import json; d=json.load(open('gpurun_out/bench13_mapnet.json')); print('mapnet', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'])"
echo "=== bench mapnet++"; timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload mapnetpp_n16t10 > gpurun_out/bench13_pp.json 2> gpurun_out/bench13p.err; tail -n 3 gpurun_out/bench13p.err; python -c "
import json; d=json.load(open('gpurun_out/bench13_pp.json')); print('mapnet++', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['frac'])"
