#!/bin/bash
# First GPU call of round 2: hardware verdict on the tcgen05 work written blind at the end of round 1.
# Everything runs under `timeout` (a hung cooperative kernel must not cost a strike); outputs -> gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- bash scripts/gpu_r2_first.sh
mkdir -p gpurun_out
A="-gencode arch=compute_100a,code=sm_100a -O2"
for p in tcgen05_score_probe tcgen05_transposed_probe tcgen05_tile_step_probe; do
  nvcc $A -o /tmp/$p experiments/$p.cu > gpurun_out/$p.build.log 2>&1 && timeout 60 /tmp/$p > gpurun_out/$p.log 2>&1
  echo "$p exit $?"; tail -4 gpurun_out/$p.log
done
# tile-step throughput under three feeding / overlap schemes (design input for the next tc5 version)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_bench experiments/tcgen05_tile_pipeline_bench.cu > gpurun_out/pipe_bench.build.log 2>&1 \
  && timeout 180 /tmp/pipe_bench > gpurun_out/pipe_bench.log 2>&1; echo "pipe_bench exit $?"; tail -8 gpurun_out/pipe_bench.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -o /tmp/gbar experiments/grid_barrier_bench.cu > gpurun_out/gbar.build.log 2>&1 \
  && timeout 60 /tmp/gbar > gpurun_out/gbar.log 2>&1; echo "grid barrier bench exit $?"; cat gpurun_out/gbar.log
# the product path must still be green before anything else is looked at
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
# stage-by-stage numbers first (tells WHERE it is wrong), then the opt-in parity tests
timeout 300 python tests/tools/debug_tc5.py > gpurun_out/debug_tc5.log 2>&1; echo "debug_tc5 exit $?"; tail -30 gpurun_out/debug_tc5.log
HMY_TEST_TC5=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -k tc5 -x -q -s > gpurun_out/pytest_tc5.log 2>&1; echo "pytest tc5 exit $?"
grep -E "^\[|passed|failed|Error|error|assert" gpurun_out/pytest_tc5.log | tail -20
for o in "tc5=0" "tc5=1"; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --engine-opt $o > gpurun_out/bench_$o.json 2> gpurun_out/bench_$o.err; echo "bench $o exit $?"
  python -c "import json; d=json.loads(open('gpurun_out/bench_$o.json').read().strip().splitlines()[-1]); print('$o value', d['value'], 'round ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'parity', d.get('parity'))"
done
HMY_ENGINE_OPTS=tc5=1 timeout 300 python scripts/trace_round.py syn1m > gpurun_out/trace_tc5.txt 2>&1; echo "trace tc5 exit $?"; tail -24 gpurun_out/trace_tc5.txt
HMY_TEST_LISI=1 timeout 300 python -m pytest tests/test_gpu_lisi.py -m gpu -x -q > gpurun_out/pytest_lisi.log 2>&1; echo "pytest lisi exit $?"; tail -5 gpurun_out/pytest_lisi.log
HMY_TEST_KMINIT=1 timeout 300 python -m pytest tests/test_gpu_kmeans_init.py -m gpu -x -q -s > gpurun_out/pytest_kminit.log 2>&1; echo "pytest kminit exit $?"; grep -E "^\[|passed|failed|Error|assert" gpurun_out/pytest_kminit.log | tail -12
HMY_TEST_DEVPERM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -k replayed_through_the_oracle -x -q -s > gpurun_out/pytest_devperm.log 2>&1; echo "pytest devperm exit $?"; grep -E "^\[|passed|failed|Error|assert" gpurun_out/pytest_devperm.log | tail -6
