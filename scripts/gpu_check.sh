#!/bin/bash
# First-contact GPU run: environment probe, smoke, parity tests.  Everything under `timeout`.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  nproc; free -g | head -2; ls /root/reference 2>&1 | head -2
} > gpurun_out/env.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log; tail -40 gpurun_out/pytest_gpu.log
