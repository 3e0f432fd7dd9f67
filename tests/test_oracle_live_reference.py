"""Differential test of the CPU oracle against the LIVE reference over randomised configurations.

The committed golden fixtures pin the oracle on three datasets; this test widens the pin across the argument space
(per-covariate / per-level theta and lambda, dynamic lambda, tau, sigma, block sizes, 1-3 covariates, small K) by
running the reference itself -- its fp64 restatement, i.e. harmonypy/harmony.py with float32 -> float64 substituted in a
temp copy exactly as tests/golden/make_golden.py does -- and the fp64 oracle on the same inputs, the same sklearn
centroids and the same permutation stream.  Agreement is at rounding level (1e-9), not at fp32 noise level.

Runs only where /root/reference exists (the build container); skipped on the GPU box.  CPU only.
"""
import os
import sys

import numpy as np
import pandas as pd
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "harmonypy", "harmony.py")),
                                reason="the reference checkout is only present in the build container")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def ref64():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import make_golden
    return make_golden, make_golden.load_reference_module(True)


def _problem(rng, N, d, levels):
    ntype = 5
    centers = rng.normal(size=(ntype, d)) * 2.0
    t = rng.integers(0, ntype, N)
    Z = centers[t] + rng.normal(size=(N, d))
    meta = {}
    for v, L in enumerate(levels):
        b = rng.integers(0, L, N)
        Z = Z + (rng.normal(size=(L, d)) * 0.8)[b]
        meta[f"v{v}"] = pd.Categorical([f"l{v}_{x}" for x in b])
    return Z.astype(np.float32), pd.DataFrame(meta)


CASES = [
    # N, d, levels, kwargs
    (700, 8, [3], dict(nclust=9)),
    (900, 6, [4, 2], dict(nclust=12, theta=[1.5, 0.5], lamb=[0.5, 2.0], sigma=0.15)),
    (800, 10, [2, 3, 2], dict(nclust=10, theta=[2.0, 1.0, 0.7, 1.2, 0.3, 2.5, 1.1], lamb=-1, alpha=0.35)),
    (650, 5, [5], dict(nclust=8, tau=7, theta=3.0, block_size=0.2, max_iter_kmeans=6)),
    (1000, 7, [3, 3], dict(nclust=14, lamb=[0.1, 0.2, 0.3, 1.0, 2.0, 3.0], block_size=0.013, epsilon_cluster=1e-3)),
    (600, 4, [2], dict(nclust=5, sigma=0.3, block_size=0.5, epsilon_harmony=1e-2)),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_oracle_fp64_equals_live_reference_fp64(ref64, case):
    from oracle.harmony_oracle import HarmonyOracle, onehot_from_codes
    make_golden, mod64 = ref64
    N, d, levels, kw = CASES[case]
    rng = np.random.default_rng(100 + case)
    Z, meta = _problem(rng, N, d, levels)
    kw = dict(kw, max_iter_harmony=kw.get("max_iter_harmony", 3))
    stage_cells = np.arange(0, N, 7)
    # the reference draws its own centroids and permutations; both are captured and replayed into the oracle
    ho, cap = make_golden.run_captured(mod64, pd.DataFrame(Z.astype(np.float64)), meta, list(meta.columns), kw, stage_cells)

    codes = np.stack([pd.Categorical(meta[v]).codes for v in meta.columns]).astype(np.int32)
    lv = np.array([len(pd.Categorical(meta[v]).categories) for v in meta.columns], dtype=np.int32)
    phi = onehot_from_codes(codes, lv, np.float64)
    assert np.array_equal(phi, ho.Phi.T)
    orc = HarmonyOracle(Z.astype(np.float64).T, phi, ho.Pr_b, ho.sigma, ho.theta, ho.lamb, alpha=float(ho.alpha),
                        lambda_estimation=bool(ho.lambda_estimation), block_size=float(ho.block_size),
                        epsilon_kmeans=float(ho.epsilon_kmeans), epsilon_harmony=float(ho.epsilon_harmony),
                        max_iter_kmeans=int(ho.max_iter_kmeans), dtype=np.float64)
    perms = iter(cap["perms"])
    orc.init_from_centroids(np.asarray(cap["Y0"], dtype=np.float64).T)
    orc.harmonize(int(ho.max_iter_harmony), lambda: next(perms))

    assert list(orc.kmeans_rounds) == list(ho.kmeans_rounds)
    scale = np.abs(ho.Z_corr).max()
    assert np.abs(orc.Z_corr.T - ho.Z_corr).max() / scale < 1e-9
    np.testing.assert_allclose(orc.objective_harmony, ho.objective_harmony, rtol=1e-10)
    np.testing.assert_allclose(orc.objective_kmeans, ho.objective_kmeans, rtol=1e-10)
    np.testing.assert_allclose(orc.R.T, ho.R, atol=1e-11)
    np.testing.assert_allclose(orc.O, ho.O, rtol=1e-9, atol=1e-10)
