"""Deterministic synthetic workloads for the benchmark configs.

``n_types`` cell types with a decaying PC spectrum, plus an additive per-level shift
for every batch covariate.  The generator is host-side NumPy; it is input plumbing,
not part of the hot path.
"""
from __future__ import annotations

import numpy as np


def make_synthetic_arrays(N, d, levels_per_var, seed=0, n_types=25, chunk=1 << 20):
    """Return (Z float32 N x d, codes int32 V x N).  Generated in chunks so 10M cells
    do not need a float64 N x d temporary."""
    rng = np.random.default_rng(seed)
    s = 10.0 / np.sqrt(np.arange(1, d + 1))
    mu = rng.standard_normal((n_types, d)) * s
    V = len(levels_per_var)
    probs = [rng.dirichlet(5.0 * np.ones(int(b))) for b in levels_per_var]
    shifts = [3.0 * rng.normal(0.0, 0.5, size=(int(b), d)) * (s / s[0]) for b in levels_per_var]
    Z = np.empty((N, d), dtype=np.float32)
    codes = np.empty((V, N), dtype=np.int32)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        n = hi - lo
        t = rng.integers(0, n_types, size=n)
        z = mu[t] + rng.standard_normal((n, d))
        for v in range(V):
            b = rng.choice(int(levels_per_var[v]), size=n, p=probs[v])
            codes[v, lo:hi] = b
            z += shifts[v][b]
        Z[lo:hi] = z.astype(np.float32)
    return Z, codes


def make_synthetic(N, d, levels_per_var, seed=0, n_types=25):
    """Same data as ``make_synthetic_arrays`` with a pandas meta_data frame whose
    columns ``var0, var1, ...`` hold string labels ``v{v}_{b:03d}`` (zero padded so the
    lexical level order used by pd.get_dummies equals the numeric code order)."""
    import pandas as pd
    Z, codes = make_synthetic_arrays(N, d, levels_per_var, seed=seed, n_types=n_types)
    cols = {}
    for v in range(codes.shape[0]):
        labels = np.array([f"v{v}_{b:03d}" for b in range(int(levels_per_var[v]))])
        cols[f"var{v}"] = labels[codes[v]]
    return Z, pd.DataFrame(cols)
