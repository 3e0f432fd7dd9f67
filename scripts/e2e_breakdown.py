import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from harmonypy_b200 import _cabi
from harmonypy_b200.harmony import Harmony
from harmonypy_b200.synthetic import make_synthetic_arrays
w = bench.WORKLOADS["syn1m"]; N = w["per_gpu"]
Z, codes = make_synthetic_arrays(N, w["d"], w["levels"], seed=0)
Pr_b = bench.global_level_probs(w, N, 0, N, codes)
Y0 = bench.init_centroids(w, N)
prob = bench.make_problem(w, Z, codes, Pr_b, N, 0)
import torch; torch.cuda.synchronize()
for rep in range(3):
    t = [time.perf_counter()]
    eng = _cabi.Engine(0, N, N, 0, w["d"], w["K"], np.asarray(w["levels"], np.int32)); t.append(time.perf_counter())
    eng.set_params(prob.Pr_b, prob.theta, prob.sigma, prob.lamb, False, 0.2, 0.05); t.append(time.perf_counter())
    eng.set_data(prob.Z, prob.codes); t.append(time.perf_counter())
    eng.close()
    ho = Harmony(prob, 0.2, 10, 20, 1e-5, 1e-4, 0.05, False, 0, 0, perm_mode="device", init_centroids=Y0, run=False); t.append(time.perf_counter())
    ho.init_cluster(0, Y0); ho.harmonize(10, False); ho._engine.synchronize(); t.append(time.perf_counter())
    out = ho.result_local(); t.append(time.perf_counter())
    del ho
    names = ["create", "set_params", "set_data", "Harmony ctor (create+params+data)", "init+harmonize", "get Z_corr"]
    print(rep, {n: round(1e3 * (b - a), 1) for n, a, b in zip(names, t[:-1], t[1:])})
