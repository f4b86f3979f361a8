#!/bin/bash
# KRSC master-weight layout: GPU parity tests + bench.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/bench_a.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches'],d.get('precision_modes'))"
