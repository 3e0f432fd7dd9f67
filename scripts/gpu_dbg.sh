#!/bin/bash
# timing experiments on the round kernel (engine option dbg: results are wrong on purpose)
for d in 0 1 2 3 4 8 12 15; do
  echo "=== dbg=$d"
  HMY_ENGINE_OPTS="dbg=$d" WRITE_R=0 timeout 120 python scripts/trace_tc5.py syn1m 2>&1 | grep -E "^grid|per block: tiles|per block: penalty|block 5 tile [012]:|block 5 tile 1 \(" | cut -c1-230
done
