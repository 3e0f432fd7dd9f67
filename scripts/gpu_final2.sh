#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"; tail -2 gpurun_out/bench_ours.err
for wl in syn10m8 syn5m8; do
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --workload $wl > gpurun_out/bench_$wl.json 2>/dev/null; echo "bench $wl exit $?"
python -c "import json; d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', 'value', d['value'], 'round ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'ridge ms', d['roofline']['ridge']['avg_pass_ms'], 'rounds/step', d['config']['rounds_per_step'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_round_mma -s 8 -c 1 -o gpurun_out/prof_round_final \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_round.log 2>&1; echo "ncu round exit $?"
timeout 900 ncu --set full --clock-control none -k regex:k_ridge -s 4 -c 3 -o gpurun_out/prof_ridge_final \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_ridge.log 2>&1; echo "ncu ridge exit $?"
timeout 600 python scripts/trace_round.py syn1m > gpurun_out/trace_syn1m.txt 2>&1
cat gpurun_out/bench_ours.json
