#!/bin/bash
# round-1 final single-GPU measurements: bench (both arms), launch list, ncu full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"; tail -2 gpurun_out/bench_ours.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_round_mma -s 8 -c 1 -o gpurun_out/prof_round_final \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_round.log 2>&1; echo "ncu round exit $?"
timeout 900 ncu --set full --clock-control none -k regex:k_ridge -s 4 -c 3 -o gpurun_out/prof_ridge_final \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_ridge.log 2>&1; echo "ncu ridge exit $?"
timeout 600 python scripts/trace_round.py syn1m > gpurun_out/trace_syn1m.txt 2>&1
cat gpurun_out/bench_ours.json
