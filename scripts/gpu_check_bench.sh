#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -4 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^\[|passed|failed|Error|error|assert" gpurun_out/pytest_gpu.log | tail -30
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"
tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
