#!/bin/bash
# N-GPU box (N=2|4|8, default 8): one bench line per workload in WLS (default syn1m); TRACE=1 adds the per-rank
# timeline of the sharded round kernel (scripts/trace_tc5_dist.py); TESTS=1 runs the sharded parity tests first.
#   scripts/gpurun_retry.sh --gpus 8 --timeout 600 -- 'WLS="syn1m syn10m8" bash scripts/gpu_multi.sh'
mkdir -p gpurun_out
N=${N:-8}
if [ "${TESTS:-0}" = 1 ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_n$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_n$N.log
fi
if [ "${TRACE:-0}" = 1 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/trace_tc5_dist.py syn1m > gpurun_out/trace_n$N.txt 2> gpurun_out/trace_n$N.err; echo "trace exit $?"; grep "wall\|block step" gpurun_out/trace_n$N.txt | cut -c1-300
fi
for WL in ${WLS:-syn1m}; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --workload $WL > gpurun_out/bench_n${N}_$WL.json 2> gpurun_out/bench_n${N}_$WL.err; echo "bench $WL exit $?"; tail -2 gpurun_out/bench_n${N}_$WL.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_$WL.json').read().strip().splitlines()[-1])
print('N=%d %s value %.1f M cells/s  ms/step %.2f  round %.3f ms  ridge %.3f ms e2e %.1f M' % (d['n_gpus'], '$WL', d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6))
print(' parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal'], v['n_ranks']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
done
