#!/bin/bash
# 2-GPU bench (fused tc5 kernel) and the sharded parity block
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.1f M cells/s  ms/step %.2f  round %.3f ms  ridge %.3f ms  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6))
print('mode', d['run']['mode']); print('parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal'], v['n_ranks'], v['round_kernel']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
