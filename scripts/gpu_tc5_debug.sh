#!/bin/bash
# Stage-by-stage check of the tensor-memory round kernel, then the GPU test suite, a timeline and a short bench.
mkdir -p gpurun_out
timeout 300 python tests/tools/debug_tc5.py > gpurun_out/debug_tc5.log 2>&1; echo "debug_tc5 exit $?"; grep -A4 "tc5': 1" gpurun_out/debug_tc5.log | tail -20
if [ "$1" != "quick" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
WRITE_R=0 timeout 300 python scripts/trace_tc5.py syn1m > gpurun_out/trace_tc5_w0.txt 2>&1; echo "trace exit $?"; cat gpurun_out/trace_tc5_w0.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_tc5.json 2> gpurun_out/bench_tc5.err; echo "bench exit $?"; tail -3 gpurun_out/bench_tc5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tc5.json').read().strip().splitlines()[-1])
print('value %.1f M cells/s  ms/step %.2f  round %.3f ms frac %.3f  ridge %.3f ms  e2e %.1f M (%.0f ms)' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6, 1e3*d['e2e']['seconds']))
print('parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
fi
