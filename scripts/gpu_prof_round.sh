#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_round -s 8 -c 1 -o gpurun_out/prof_round_mma \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_round.log 2>&1; echo "ncu round exit $?"
tail -3 gpurun_out/ncu_round.log
