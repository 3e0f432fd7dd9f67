#!/bin/bash
mkdir -p gpurun_out
timeout 60 experiments/_bin/mma_rate > gpurun_out/mma_rate.txt 2>&1; echo "mma_rate exit $?"; cat gpurun_out/mma_rate.txt
WRITE_R=0 timeout 300 python scripts/trace_tc5.py syn1m > gpurun_out/trace_tc5_w0.txt 2>&1; echo "trace exit $?"; cat gpurun_out/trace_tc5_w0.txt
