#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/trace_round.py syn1m > gpurun_out/trace_syn1m.txt 2>&1; echo "trace exit $?"
tail -22 gpurun_out/trace_syn1m.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for o in "mma_wn=0" "mma_wn=2"; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --engine-opt $o 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o value',d['value'],'round ms',d['roofline']['avg_launch_ms'],'frac',d['roofline']['frac'],'ridge ms',d['roofline']['ridge']['avg_pass_ms'])"
done
