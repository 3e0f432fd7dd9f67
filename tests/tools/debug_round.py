import sys, os
import numpy as np, pandas as pd
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from harmonypy_b200.harmony import Harmony, prepare_problem
from harmonypy_b200.synthetic import make_synthetic
from oracle.harmony_oracle import HarmonyOracle, onehot_from_codes, torch_perm_source

def rel(a, b): return float(np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / np.abs(b).max())

for (N, d, levels, K) in [(9000, 50, [30, 4], 200), (9000, 50, [30], 200), (9000, 50, [30, 4], 100)]:
    Z, meta = make_synthetic(N, d, levels, seed=11)
    prob, _ = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), nclust=K)
    Y0 = Z[np.random.default_rng(1).choice(N, K, replace=False)]
    for opts in ({"persistent": 1}, {"persistent": 0}, {"persistent": 1, "mma": 0}):
        ho = Harmony(prob, 0.2, 1, 3, 1e-5, 1e-4, 0.05, False, 3, 0, run=False, engine_options=opts)
        orc = HarmonyOracle(prob.Z.T, onehot_from_codes(prob.codes, prob.levels, np.float64), prob.Pr_b, prob.sigma,
                            prob.theta, prob.lamb, dtype=np.float64)
        ho.init_cluster(3, Y0); orc.init_from_centroids(Y0.T)
        print(N, opts, "init  R %.2e O %.2e" % (rel(ho.R, orc.R.T), rel(ho.O, orc.O)))
        src = torch_perm_source(N, 3)
        for r in range(3):
            ho.kmeans_round(); orc.kmeans_round(src())
            print("   round", r, "Y %.2e R %.2e O %.2e obj %.3e vs %.3e" % (rel(ho.Y.T, orc.Y.T), rel(ho.R, orc.R.T), rel(ho.O, orc.O),
                  ho.objective_kmeans[-1], orc.objective_kmeans[-1]), "Osum", float(ho.O.sum()), float(orc.O.sum()))
