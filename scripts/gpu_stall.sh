#!/bin/bash
# which wait absorbs the rare long tile?  10 traced rounds on one GPU
mkdir -p gpurun_out
WRITE_R=0 TRACE_ROUNDS=${TRACE_ROUNDS:-12} timeout 300 python scripts/trace_tc5.py syn1m > gpurun_out/trace_stall.txt 2>&1; echo "trace exit $?"
tail -25 gpurun_out/trace_stall.txt
