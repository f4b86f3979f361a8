#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x > gpurun_out/pytest_ddp.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_ddp.log
