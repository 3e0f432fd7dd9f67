#!/bin/bash
# last call of the round on one B200: bench line (parity block + e2e check), then the tests that touch the new host paths
mkdir -p gpurun_out
timeout 85 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1_pinned.json 2> gpurun_out/bench_n1_pinned.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1_pinned.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n1_pinned.json').read().strip().splitlines()[-1])
    print('value %.1f M cells/s  ms/step %.2f  round %.3f ms ridge %.3f ms e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['ridge']['avg_pass_ms'], d['e2e']['value']/1e6), d['e2e'].get('seconds_all'), d['e2e'].get('max_rel_diff_vs_resident_run'), d['e2e'].get('dma_direct'))
    print(' parity', {k:(v['vs_reference_fp32'], v['kmeans_rounds_equal']) for k,v in d['parity'].items() if isinstance(v,dict)})
except Exception as e: print('no bench line', e)
PY
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "page_locked or golden_ircolitis or end_to_end or hundreds or fp64_oracle" > gpurun_out/pytest_pinned.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_pinned.log
