#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity_r2.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0 1 0; do
MAPNET_BN_LAZY_FIN=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-modes > gpurun_out/bench_lazy$v.json 2> gpurun_out/bench_lazy$v.err; echo "bench lazy=$v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/bench_lazy$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches'])"
done
