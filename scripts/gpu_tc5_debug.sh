#!/bin/bash
# Stage-by-stage check of the tensor-memory round kernel, then the GPU test suite and a short bench.
mkdir -p gpurun_out
timeout 300 python tests/tools/debug_tc5.py > gpurun_out/debug_tc5.log 2>&1; echo "debug_tc5 exit $?"; tail -40 gpurun_out/debug_tc5.log
if [ "$1" != "quick" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench exit $?"; tail -3 gpurun_out/bench_tc5.err; tail -1 gpurun_out/bench_tc5.json | cut -c1-1500
fi
