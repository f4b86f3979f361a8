#!/bin/bash
# Full GPU validation + transposes-on-side-stream A/B.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
for v in 1 0 1 0; do
MAPNET_TR_ASYNC=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-modes > gpurun_out/bench_tr$v.json 2> gpurun_out/bench_tr$v.err; echo "bench tr=$v rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/bench_tr$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['gpu_launches'])"
done
