#!/bin/bash
# End-of-round evidence on one B200: GPU test suite, bench line (+ reference arm), launch list, ncu capture, timeline.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm exit $?"; tail -1 gpurun_out/bench_ref.json | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value %.1f M cells/s  ms/step %.2f  round %.3f ms frac %.3f  ridge %.3f ms  e2e %.1f M (%.0f ms)  cpu %s' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6, 1e3*d['e2e']['seconds'], d['cpu_baseline']))
print('parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal']) for k,v in d['parity'].items() if isinstance(v,dict)})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
WRITE_R=0 TRACE_ROUNDS=12 timeout 300 python scripts/trace_tc5.py syn1m > gpurun_out/trace_tc5_w0.txt 2>&1; echo "trace exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_round_tc5 -s 3 -c 1 -o gpurun_out/prof_tc5 python scripts/trace_tc5.py syn1m > gpurun_out/ncu_tc5.log 2>&1; echo "ncu exit $?"
