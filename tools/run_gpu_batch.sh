mkdir -p gpurun_out
echo "=== ddp test"; timeout 200 python -m pytest tests/test_gpu_ddp.py -m gpu -q 2>&1 | tail -2
echo "=== 2-GPU bench"; timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -n 4 gpurun_out/bench_n2.err | cut -c1-200; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print('n2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['n_gpus'], d['config']['eager_ms_per_step'])"
echo "=== 2-GPU reference arm"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
