mkdir -p gpurun_out
echo "=== conv unit tests (auto tiles)"; timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_engines and bf16 and not simt" 2>&1 | tail -3
echo "=== bench auto"; timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -n 4 gpurun_out/bench9.err; python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print('auto', d['value'], d['ms_per_step'], 'eager', d['config']['eager_ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== step parity bf16"; timeout 200 python -m pytest tests/test_gpu_step.py -m gpu -q -k "bf16" 2>&1 | tail -3
echo "=== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1150 -c 300 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1; wc -l gpurun_out/launches_r1c.csv
