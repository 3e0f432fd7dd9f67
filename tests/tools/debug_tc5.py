"""Stage-by-stage comparison of the tensor-memory round kernel (engine option tc5) with the fp64 oracle and
with the mma.sync kernel: init assignment, then three rounds.  A debugging aid for a B200 box, not a test:

    timeout 300 python tests/tools/debug_tc5.py            # prints one line per stage and kernel

Reading the output: `init` exercises scoring + epilogue + both accumulations without penalty or block
lists; a wrong R with a right O-sum points at the scoring/epilogue, a right R with wrong Y/O at the
MN-major accumulation operands (see experiments/tcgen05_tile_step_probe.cu for the LBO/SBO arbitration).
"""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from harmonypy_b200.harmony import Harmony, prepare_problem          # noqa: E402
from harmonypy_b200.synthetic import make_synthetic                  # noqa: E402
from oracle.harmony_oracle import HarmonyOracle, onehot_from_codes, torch_perm_source   # noqa: E402


def rel(a, b):
    return float(np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / np.abs(b).max())


def main():
    cases = [(3000, 20, [3], 40), (9000, 50, [5, 3], 100), (20000, 50, [11], 128)]
    for (N, d, levels, K) in cases:
        Z, meta = make_synthetic(N, d, levels, seed=11)
        prob, _ = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), nclust=K)
        Y0 = Z[np.random.default_rng(1).choice(N, K, replace=False)]
        for opts in ({"tc5": 1}, {"tc5": 0}):
            ho = Harmony(prob, 0.2, 1, 3, 1e-5, 1e-4, 0.05, False, 3, 0, run=False, engine_options=opts)
            orc = HarmonyOracle(prob.Z.T, onehot_from_codes(prob.codes, prob.levels, np.float64), prob.Pr_b,
                                prob.sigma, prob.theta, prob.lamb, dtype=np.float64)
            ho.init_cluster(3, Y0)
            orc.init_from_centroids(Y0.T)
            print(f"N={N} d={d} levels={levels} K={K} {opts} grid={ho._engine.counter('grid')} "
                  f"smem={ho._engine.counter('smem_round')}", flush=True)
            print("   init     R %.2e O %.2e Osum %.6f vs %.6f" % (rel(ho.R, orc.R.T), rel(ho.O, orc.O),
                  float(ho.O.sum()), float(orc.O.sum())), flush=True)
            src = torch_perm_source(N, 3)
            for r in range(3):
                ho.kmeans_round()
                orc.kmeans_round(src())
                print("   round %d  Y %.2e R %.2e O %.2e obj %.6e vs %.6e" % (
                    r, rel(ho.Y.T, orc.Y.T), rel(ho.R, orc.R.T), rel(ho.O, orc.O),
                    ho.objective_kmeans[-1], orc.objective_kmeans[-1]), flush=True)


if __name__ == "__main__":
    main()
