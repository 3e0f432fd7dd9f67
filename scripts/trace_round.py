#!/usr/bin/env python
"""Timeline of the persistent round kernel from its per-CTA globaltimer stamps (option trace)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from harmonypy_b200.harmony import Harmony
from harmonypy_b200.synthetic import make_synthetic_arrays

wl = sys.argv[1] if len(sys.argv) > 1 else "syn1m"
w = bench.WORKLOADS[wl]
N = w["per_gpu"]
Z, codes = make_synthetic_arrays(N, w["d"], w["levels"], seed=0)
Pr_b = bench.global_level_probs(w, N, 0, N, codes)
Y0 = bench.init_centroids(w, N)
prob = bench.make_problem(w, Z, codes, Pr_b, N, 0)
# engine options from the environment, e.g. HMY_ENGINE_OPTS="tc5=1"
opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("HMY_ENGINE_OPTS", "").split(",") if kv)}
ho = Harmony(prob, 0.2, 10, 20, 1e-5, 1e-4, 0.05, False, 0, 0, perm_mode="device", run=False, engine_options=opts or None)
eng = ho._engine
ho.init_cluster(0, Y0)
for _ in range(3):
    ho.kmeans_round()
eng.set_option("trace", 1)
ho.kmeans_round()
G = eng.counter("grid")
buf = np.zeros((G + 1, 256), dtype=np.uint64)
eng._ck(eng.lib.hmy_get(eng.h, 9, buf.ctypes.data_as(C.c_void_p), buf.nbytes), "trace")
ser = buf[G].astype(np.int64).reshape(32, 4)
t = buf[:G].astype(np.int64)
t0 = t[:, 0].min()
nblk = eng.counter("nblk")
def stat(x): return f"min {x.min()/1e3:8.1f}  mean {x.mean()/1e3:8.1f}  max {x.max()/1e3:8.1f} us"
print(f"grid {G}, nblk {nblk}, kernel span {(t[:, 5+3*(nblk-1)].max()-t0)/1e3:.1f} us")
print("start skew      ", stat(t[:, 0] - t0))
print("phase0          ", stat(t[:, 1] - t[:, 0]))
print("barrier0 (wait) ", stat(t[:, 2] - t[:, 1]))
pl, pr, bw = [], [], []
for b in range(nblk):
    prev = t[:, 2] if b == 0 else t[:, 5 + 3 * (b - 1)]
    pl.append(t[:, 3 + 3 * b] - prev); pr.append(t[:, 4 + 3 * b] - t[:, 3 + 3 * b]); bw.append(t[:, 5 + 3 * b] - t[:, 4 + 3 * b])
pl, pr, bw = np.array(pl), np.array(pr), np.array(bw)
print("per block: penalty load  ", stat(pl))
print("per block: process       ", stat(pr), " (per-block max over CTAs, mean over blocks: %.1f us)" % (pr.max(axis=1).mean() / 1e3))
print("per block: barrier wait  ", stat(bw), " (min over CTAs = serial+handshake, mean over blocks: %.1f us)" % (bw.min(axis=1).mean() / 1e3))
step = np.array([t[:, 5 + 3 * b].max() for b in range(nblk)]); step = np.diff(np.concatenate([[t[:, 2].max()], step]))
print("block step time (release to release): mean %.1f us" % (step.mean() / 1e3), np.round(step / 1e3, 1))
ok = ser[:, 0] > 0
for name, a, b in (("fence", 0, 1), ("body+sync", 1, 2), ("reset+fence", 2, 3), ("total", 0, 3)):
    d = (ser[ok, b] - ser[ok, a])
    print(f"serial {name:12s}", stat(d))
# per-tile stamps of block 5 (thread 0 = warp 0): tile start, ids ready, Z ready, scores done, epilogue done, sync, Y-GEMM done
tt = t[:, 64:125]
names = ["ids+lev load", "Z gather", "score MMA", "epilogue", "sync wait", "Y-GEMM"]
nst = 7
if eng.counter("tc5") == 1:
    # tensor-memory kernel: 5 stamps per tile (stage begin, operands published, scores ready, epilogue done, R published)
    names = ["(pre-staged) gather+publish", "barrier + tables + score wait", "epilogue", "publish R", "next: stage begin"]
    nst = 6
rows = []
for c in range(G):
    x = tt[c]; x = x[x > 0]
    if len(x) >= nst + 1: rows.append(x[:nst] - x[0])
rows = np.array(rows)
print("first tile of block 5, cumulative us (mean over CTAs):", np.round(rows.mean(axis=0) / 1e3, 2), names)
pr5 = t[:, 4 + 15] - t[:, 3 + 15]
slow = np.argsort(-pr5)[:8]
print("block 5 slowest CTAs:", [(int(c), round(pr5[c] / 1e3, 1)) for c in slow], " fastest:", round(pr5.min() / 1e3, 1))
allp = np.array([t[:, 4 + 3 * b] - t[:, 3 + 3 * b] for b in range(nblk)])
print("per-CTA mean process time over blocks: min %.1f max %.1f ; corr of slow set between blocks 5,6: %.2f" % (
    allp.mean(axis=0).min() / 1e3, allp.mean(axis=0).max() / 1e3, np.corrcoef(allp[5], allp[6])[0, 1]))
for c in slow[:3]:
    x = tt[c]; x = x[x > 0]
    print("CTA", int(c), "stamps (us):", np.round((x - x[0]) / 1e3, 1))
