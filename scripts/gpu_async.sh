#!/bin/bash
# 2-GPU box: parity tests, then the bench at N=1 and N=2 (rounds enqueued without host round trips)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_async.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_async.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ('bench_n1','bench_n2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f,'value %.1f M cells/s  ms/step %.2f  round %.3f ms frac %.3f ridge %.3f ms  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6))
    print(' run', d['run']); print(' parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
WRITE_R=0 TRACE_ROUNDS=12 timeout 300 python scripts/trace_tc5.py syn1m > gpurun_out/trace_stall.txt 2>&1; echo "trace exit $?"
tail -24 gpurun_out/trace_stall.txt | cut -c1-220
