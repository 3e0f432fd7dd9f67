#!/bin/bash
mkdir -p gpurun_out
HMY_TIMING=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/e2e_breakdown.txt 2>&1; echo "e2e exit $?"; tail -40 gpurun_out/e2e_breakdown.txt
WRITE_R=0 timeout 300 python scripts/trace_tc5.py syn1m 2>&1 | grep -E "^grid|per block|block 5 tile|MMA warp|tile phase per block: m" 
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_round_tc5 -s 3 -c 1 -o gpurun_out/prof_tc5 python scripts/trace_tc5.py syn1m > gpurun_out/ncu_tc5.log 2>&1; echo "ncu exit $?"
