#!/usr/bin/env python
"""Per-rank timeline of the sharded tensor-memory round kernel (k_round_tc5<NC, MULTI>) under torchrun:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/trace_tc5_dist.py [workload]
Every rank dumps its stamps to gpurun_out/trace_rank<r>.npy; rank 0 prints the per-block summary of every rank."""
import sys, os, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from harmonypy_b200.harmony import Harmony, Comm
from harmonypy_b200.synthetic import make_synthetic_arrays

wl = sys.argv[1] if len(sys.argv) > 1 else "syn1m"
w = bench.WORKLOADS[wl]
world, rank, local_rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dist.barrier()
N_total = w["per_gpu"] * world
lo, hi = N_total * rank // world, N_total * (rank + 1) // world
Z, codes = make_synthetic_arrays(N_total, w["d"], w["levels"], seed=0, lo=lo, hi=hi)
Pr_b = bench.global_level_probs(w, N_total, lo, hi, codes, dist)
Y0 = bench.init_centroids(w, N_total) if rank == 0 else np.zeros((w["K"], w["d"]), np.float32)
comm = Comm(None)
Y0 = comm.broadcast_array(Y0, 0)
prob = bench.make_problem(w, Z, codes, Pr_b, N_total, lo)
ho = Harmony(prob, 0.2, 10, 20, 1e-5, 1e-4, 0.05, False, 0, local_rank, perm_mode="device", comm=comm, run=False)
eng = ho._engine
assert eng.counter("tc5") == 1 and eng.counter("fused") == 1, (eng.counter("tc5"), eng.counter("fused"))
ho.init_cluster(0, Y0)
eng.set_option("write_r", 0)

def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

for _ in range(3):
    eng.kmeans_round(None)
out = []
for mode in ("wait", "nowait", "wait", "nowait"):
    sync()
    r0, m0 = eng.counter("rounds"), eng.timer_ms("ms_round")
    t0 = time.perf_counter()
    if mode == "wait":
        for _ in range(8):
            eng.kmeans_round(None)
    else:
        for _ in range(8):
            eng.kmeans_round(None, wait=False)
        eng.objectives(8)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 8 * 1e3
    ker = (eng.timer_ms("ms_round") - m0) / max(1, eng.counter("rounds") - r0)
    out.append((mode, wall, ker))
sync()
eng.set_option("trace", 1)
eng.kmeans_round(None)
torch.cuda.synchronize()
G = eng.counter("grid")
buf = np.zeros((G + 1, 256), dtype=np.uint64)
eng._ck(eng.lib.hmy_get(eng.h, 9, buf.ctypes.data_as(C.c_void_p), buf.nbytes), "trace")
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/trace_rank{rank}.npy", buf)
lines = [f"rank {rank}: " + "  ".join(f"{m}: wall {a:.3f} ms/round, kernel {b:.3f} ms" for m, a, b in out)]
try:
    t = buf[:G].astype(np.int64)
    nblk = eng.counter("nblk")
    t0 = t[:, 0].min()
    lines.append(f"rank {rank}: grid {G}, nblk {nblk}, span start -> last barrier {(t[:, 6 + 5 * (nblk - 1)].max() - t0) / 1e3:.1f} us")
    names = ["wait first tile staged", "penalty rows (LL sums)", "tiles", "o-done wait + flush", "grid barrier + push"]
    acc = [[] for _ in names]
    for b in range(nblk):
        prev = t[:, 1] if b == 0 else t[:, 6 + 5 * (b - 1)]
        s = [t[:, 2 + 5 * b + i] for i in range(5)]
        acc[0].append(s[0] - prev); acc[1].append(s[1] - s[0]); acc[2].append(s[2] - s[1]); acc[3].append(s[3] - s[2]); acc[4].append(s[4] - s[3])
    for n, a in zip(names, acc):
        a = np.array(a)[1:-1]
        lines.append(f"rank {rank}: per block: {n:24s} min {a.min()/1e3:7.2f}  mean {a.mean()/1e3:7.2f}  max {a.max()/1e3:7.2f} us | max over CTAs, mean over blocks {a.max(axis=1).mean()/1e3:6.2f}")
    rel = np.array([t[:, 6 + 5 * b].max() for b in range(nblk)])
    lines.append(f"rank {rank}: block step (release to release) us: " + str(np.round(np.diff(rel) / 1e3, 2).tolist()))
    pen = np.array([t[:, 3 + 5 * b] - t[:, 2 + 5 * b] for b in range(nblk)])
    lines.append(f"rank {rank}: penalty rows per block, mean over CTAs: " + str(np.round(pen.mean(axis=1) / 1e3, 1).tolist()))
    lines.append(f"rank {rank}: penalty rows per block, min over CTAs:  " + str(np.round(pen.min(axis=1) / 1e3, 1).tolist()))
except Exception as e:          # the raw stamps are on disk either way
    lines.append(f"rank {rank}: analysis failed: {e!r}")
gathered = [None] * world
dist.all_gather_object(gathered, lines)
if rank == 0:
    for ls in gathered:
        print("\n".join(ls))
dist.barrier()
dist.destroy_process_group()
