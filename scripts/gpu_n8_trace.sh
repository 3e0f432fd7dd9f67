#!/bin/bash
# 8-GPU box: timeline of the sharded round kernel on every rank, then a short bench line
mkdir -p gpurun_out
N=${N:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/trace_tc5_dist.py syn1m > gpurun_out/trace_n$N.txt 2> gpurun_out/trace_n$N.err; echo "trace exit $?"; tail -3 gpurun_out/trace_n$N.err; grep -c . gpurun_out/trace_n$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/bench_n${N}_async.json 2> gpurun_out/bench_n${N}_async.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n${N}_async.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_async.json').read().strip().splitlines()[-1])
print('N=%d value %.1f M cells/s  ms/step %.2f  round %.3f ms  ridge %.3f ms' % (d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['ridge']['avg_pass_ms']))
PY
grep "wall" gpurun_out/trace_n$N.txt
