#!/bin/bash
# last call of the round on one B200: bench line first, then the GPU test suite
mkdir -p gpurun_out
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1_last.json 2> gpurun_out/bench_n1_last.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1_last.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_last.json').read().strip().splitlines()[-1])
print('value %.1f M cells/s  ms/step %.2f  round %.3f ms frac %.3f ridge %.3f ms frac %.3f e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['ridge']['avg_pass_ms'], d['roofline']['ridge']['frac'], d['e2e']['value']/1e6))
print(' parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
timeout 280 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_last.log
