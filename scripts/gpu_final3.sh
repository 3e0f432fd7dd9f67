#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
