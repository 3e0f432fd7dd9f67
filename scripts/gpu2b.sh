#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -s 2>&1 | grep -E "Z_corr|passed|failed|rror" | tail -12
for o in "relaxed=0" "relaxed=1"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --engine-opt $o 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o', 'value', d['value'], 'ms/step', d['ms_per_step'], 'round ms', d['roofline']['avg_launch_ms'], 'rounds/step', d['config']['rounds_per_step'])"
done
