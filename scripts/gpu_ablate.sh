#!/bin/bash
# Ablation timings + one ncu capture of the tile-pipeline experiment (experiments/tcgen05_tile_pipeline_bench.cu)
mkdir -p gpurun_out
for a in 0 1 2 3 4 8 12 15 16 32 64 127; do
  for m in 2 3; do timeout 60 experiments/_bin/pb_$a --time-only $m 2>&1 | grep ABL; done
done | tee gpurun_out/ablate.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_pipeline -s 1 -c 1 -o gpurun_out/prof_pipe3 experiments/_bin/pb_0 --time-only 3 > gpurun_out/ncu_pipe3.log 2>&1; echo "ncu exit $?"
