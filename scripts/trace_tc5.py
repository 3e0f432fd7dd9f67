#!/usr/bin/env python
"""Timeline of the tensor-memory round kernel (k_round_tc5) from its per-CTA globaltimer stamps (option trace).
Slots: see T5_STAMP in harmonypy_b200/csrc/hmy_round_tc5.cuh."""
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
opts = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("HMY_ENGINE_OPTS", "").split(",") if kv)}
ho = Harmony(prob, 0.2, 10, 20, 1e-5, 1e-4, 0.05, False, 0, 0, perm_mode="device", run=False, engine_options=opts or None)
eng = ho._engine
assert eng.counter("tc5") == 1
ho.init_cluster(0, Y0)
eng.set_option("write_r", int(os.environ.get("WRITE_R", "0")))
for _ in range(3):
    ho.kmeans_round()
eng.set_option("trace", 1)
for _ in range(int(os.environ.get("TRACE_ROUNDS", "1"))):
    ho.kmeans_round()
G = eng.counter("grid")
buf = np.zeros((G + 1, 256), dtype=np.uint64)
eng._ck(eng.lib.hmy_get(eng.h, 9, buf.ctypes.data_as(C.c_void_p), buf.nbytes), "trace")
t = buf[:G].astype(np.int64)
t0 = t[:, 0].min()
nblk = eng.counter("nblk")
def stat(x): return f"min {x.min()/1e3:7.2f}  mean {x.mean()/1e3:7.2f}  max {x.max()/1e3:7.2f} us"
print(f"grid {G}, nblk {nblk}, write_r {os.environ.get('WRITE_R', '0')}, span start -> last barrier {(t[:, 6 + 5 * (nblk - 1)].max() - t0) / 1e3:.1f} us; ms_round timer {eng.timer_ms('ms_round'):.3f} over {eng.counter('rounds')} rounds")
print("prologue        ", stat(t[:, 1] - t[:, 0]))
names = ["wait first tile staged", "penalty rows", "tiles", "o-done wait + flush", "grid barrier"]
acc = [[] for _ in names]
for b in range(nblk):
    prev = t[:, 1] if b == 0 else t[:, 6 + 5 * (b - 1)]
    s = [t[:, 2 + 5 * b + i] for i in range(5)]
    acc[0].append(s[0] - prev); acc[1].append(s[1] - s[0]); acc[2].append(s[2] - s[1]); acc[3].append(s[3] - s[2]); acc[4].append(s[4] - s[3])
for n, a in zip(names, acc):
    a = np.array(a)[1:-1]       # blocks 1 .. nblk-2
    print(f"per block: {n:24s}", stat(a), f"| max over CTAs, mean over blocks {a.max(axis=1).mean() / 1e3:6.2f}")
rel = np.array([t[:, 6 + 5 * b].max() for b in range(nblk)])
print("block step (release to release) us:", np.round(np.diff(rel) / 1e3, 2))
tl = t[:, 102:126].reshape(G, 4, 6)
ok = (tl[:, :, 5] > 0)
for i in range(4):
    m = ok[:, i]
    if m.sum() == 0: continue
    x = tl[m, i, :] - t[m, 3 + 5 * 5][:, None]
    print(f"block 5 tile {i} ({m.sum()} CTAs), us after penalty rows: scores ready {x[:,0].mean()/1e3:6.2f}  in regs {x[:,1].mean()/1e3:6.2f}  pass1 {x[:,2].mean()/1e3:6.2f}  sums met {x[:,3].mean()/1e3:6.2f}  operand free {x[:,4].mean()/1e3:6.2f}  written {x[:,5].mean()/1e3:6.2f}")

base = t[:, 3 + 5 * 5][:, None]
mm = t[:, 128:144].reshape(G, 4, 4); pp = t[:, 144:160].reshape(G, 4, 4)
for i in range(4):
    m = mm[:, i, 0] > 0
    if m.sum() == 0: continue
    x = (mm[m, i, :] - base[m]) / 1e3; y = (pp[m, i, :3] - base[m]) / 1e3
    print(f"block 5 tile {i}: MMA thread: score issue {x[:,0].mean():7.2f} committed {x[:,1].mean():7.2f} | acc issue {x[:,2].mean():7.2f} committed {x[:,3].mean():7.2f}"
          f" || producer: stage free {y[:,0].mean():7.2f} described {y[:,1].mean():7.2f} copies issued {y[:,2].mean():7.2f}   (us relative to block 5's penalty rows)")

# ---- who is slow?  per-CTA mean of the tile phase, and whether the slow CTAs repeat
tiles = np.array([t[:, 4 + 5 * b] - t[:, 3 + 5 * b] for b in range(1, nblk - 1)])      # [blocks][G]
pen = np.array([t[:, 3 + 5 * b] - t[:, 2 + 5 * b] for b in range(1, nblk - 1)])
m = tiles.mean(axis=0) / 1e3
print("per-CTA mean tile phase (us): min %.2f  p50 %.2f  p90 %.2f  max %.2f" % (m.min(), np.median(m), np.percentile(m, 90), m.max()))
order = np.argsort(-m)
print("slowest CTAs:", [(int(c), round(float(m[c]), 2)) for c in order[:10]])
arg = tiles.argmax(axis=1)
print("argmax CTA per block:", arg.tolist())
print("tile phase per block: mean over CTAs", np.round(tiles.mean(axis=1) / 1e3, 1).tolist())
print("tile phase per block: max over CTAs ", np.round(tiles.max(axis=1) / 1e3, 1).tolist())
print("penalty phase per block: max over CTAs", np.round(pen.max(axis=1) / 1e3, 1).tolist(), "argmax", pen.argmax(axis=1).tolist())
# distribution within one block
b = 5
x = (t[:, 4 + 5 * b] - t[:, 3 + 5 * b]) / 1e3
print("block 5 tile phase histogram (us):", np.histogram(x, bins=[0, 9, 10, 11, 12, 13, 14, 16, 20, 40])[0].tolist())

cy = t[:, 160:166]
m = cy[:, 0] > 0
if m.sum():
    d = np.diff(cy[m], axis=1).mean(axis=0)
    print("MMA warp, accumulation of block 5 tile 0, SM cycles: header reads %.0f | slot-sum MMAs %.0f | (o-done commit) | centroid MMAs %.0f | commit %.0f | one globaltimer stamp %.0f" % (d[0], d[1], d[2], d[3], d[4]))

# ---- the slowest tile of every CTA over all traced launches (slots 166..183), and the longest gaps of the producer / MMA warp
gap = t[:, 174] / 1e3
order = np.argsort(-gap)
print("slowest tile per CTA (us): min %.2f  p50 %.2f  p90 %.2f  max %.2f" % (gap.min(), np.median(gap), np.percentile(gap, 90), gap.max()))
print("CTA  blk tile   gap | after the previous tile's end: described  scores ready  in regs  pass 1  sums met  operand free  end")
for c in order[:12]:
    prev = t[c, 183]
    rel = [(t[c, 176 + i] - prev) / 1e3 for i in range(6)] + [(t[c, 182] - prev) / 1e3]
    print("%3d  %3d %4d %6.2f | " % (c, t[c, 175] >> 8, t[c, 175] & 255, gap[c]) + "  ".join("%8.2f" % x for x in rel))
for name, base, n, sites in (("producer", 184, 3, ["stage free", "described", "copies issued"]), ("MMA warp", 192, 4, ["score issue", "score commit", "acc issue", "acc commit"])):
    for i, sn in enumerate(sites):
        g = t[:, base + 1 + i] / 1e3
        c = int(np.argmax(g))
        print(f"{name}: longest gap ending at '{sn}': p50 {np.median(g):.2f}  p99 {np.percentile(g, 99):.2f}  max {g.max():.2f} us (CTA {c}, tag {int(t[c, base + 1 + n + i])})")
