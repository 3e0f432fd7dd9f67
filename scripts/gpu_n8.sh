#!/bin/bash
# 8-GPU runs: sharded parity tests (4 and 8 ranks), weak-scaling bench, BASELINE configs 4 and 5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -s -k "(4-1 or 8-1) and (pbmc or ircolitis)" > gpurun_out/pytest_dist_n8.log 2>&1; echo "pytest dist exit $?"; grep -E "GPUs|passed|failed" gpurun_out/pytest_dist_n8.log | tail -8
for wl in syn1m syn10m8 syn5m8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu --workload $wl > gpurun_out/bench_n8_$wl.json 2> gpurun_out/bench_n8_$wl.err; echo "bench $wl exit $?"; tail -2 gpurun_out/bench_n8_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n8_$wl.json').read().strip().splitlines()[-1])
    print('$wl N=8 value %.1f M cells/s  ms/step %.2f  round %.3f ms frac %.3f  ridge %.3f ms  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6))
    print('  mode', d['run']['mode'], '| rounds/step', d['run']['rounds_per_step']); print('  parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal'], v['n_ranks'], v['round_kernel']) for k,v in d['parity'].items() if isinstance(v,dict)})
except Exception as e: print('no line', e)
PY
done
