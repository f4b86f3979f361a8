mkdir -p gpurun_out
echo "=== 2cta bench"; MAPNET_TC_2CTA=1 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench8_2cta.json 2> gpurun_out/bench8_2cta.err; tail -n 4 gpurun_out/bench8_2cta.err; python -c "
import json; d=json.load(open('gpurun_out/bench8_2cta.json')); print('2cta', d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== 1cta bench"; timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench8_1cta.json 2> gpurun_out/bench8_1cta.err; tail -n 4 gpurun_out/bench8_1cta.err; python -c "
import json; d=json.load(open('gpurun_out/bench8_1cta.json')); print('1cta', d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'], {k:round(v['tflops']) for k,v in d['roofline']['per_class'].items()})"
echo "=== graph parity"; timeout 200 python -m pytest tests/test_gpu_graph.py -m gpu -q 2>&1 | grep -E "^E|passed|failed" | head -12
echo "=== 2cta ncu"; MAPNET_TC_2CTA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_conv2 -s 30 -c 6 -o gpurun_out/prof_conv2_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1; ls -la gpurun_out/prof_conv2_r1.ncu-rep
