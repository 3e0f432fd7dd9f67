#!/bin/bash
# 8-GPU box: one bench line per workload given (default syn1m)
mkdir -p gpurun_out
N=${N:-8}
for WL in ${WLS:-syn1m}; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --workload $WL > gpurun_out/bench_n${N}_$WL.json 2> gpurun_out/bench_n${N}_$WL.err; echo "bench $WL exit $?"; tail -2 gpurun_out/bench_n${N}_$WL.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_$WL.json').read().strip().splitlines()[-1])
print('N=%d %s value %.1f M cells/s  ms/step %.2f  round %.3f ms  ridge %.3f ms e2e %.1f M' % (d['n_gpus'], '$WL', d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6))
print(' parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal'], v['n_ranks']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
done
