mkdir -p gpurun_out
for bo in 1 0; do
echo "=== halo conv unit tests baseoff=$bo"; MAPNET_TC_HALO_BASEOFF=$bo timeout 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_engines and bf16 and not simt" 2>&1 | grep -E "passed|failed|FAILED|Error" | head -14
done
echo "=== bench halo (baseoff=1)"; timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -n 3 gpurun_out/bench10.err; python -c "
import json; d=json.load(open('gpurun_out/bench10.json')); print('halo', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== bench halo (baseoff=0)"; MAPNET_TC_HALO_BASEOFF=0 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench10b.json 2> gpurun_out/bench10b.err; tail -n 3 gpurun_out/bench10b.err; python -c "
import json; d=json.load(open('gpurun_out/bench10b.json')); print('halo0', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
