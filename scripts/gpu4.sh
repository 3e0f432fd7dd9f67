#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
n=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err; echo "bench $n exit $?"; tail -3 gpurun_out/bench_g$n.err
python -c "import json; d=json.loads(open('gpurun_out/bench_g$n.json').read().strip().splitlines()[-1]); print($n, 'value', d['value'], 'ms/step', d['ms_per_step'], 'round ms', d['roofline']['avg_launch_ms'], 'e2e', d['e2e'] and d['e2e']['value'], d['config']['mode'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 1 --warmup 0 --cpu-sample 20000 2>/dev/null | tail -1 | cut -c1-200
