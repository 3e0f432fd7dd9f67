"""Deterministic synthetic workloads for the benchmark configs.

``n_types`` cell types with a decaying PC spectrum, plus an additive per-level shift
for every batch covariate.  The generator is host-side NumPy; it is input plumbing,
not part of the hot path.  Cells are generated in fixed chunks that are seeded
individually, so any rank can materialise any cell range of the same global dataset.
"""
from __future__ import annotations

import numpy as np

CHUNK = 1 << 18


def synthetic_params(d, levels_per_var, seed=0, n_types=25):
    rng = np.random.default_rng([int(seed), 0])
    s = 10.0 / np.sqrt(np.arange(1, d + 1))
    mu = rng.standard_normal((n_types, d)) * s
    probs = [rng.dirichlet(5.0 * np.ones(int(b))) for b in levels_per_var]
    shifts = [3.0 * rng.normal(0.0, 0.5, size=(int(b), d)) * (s / s[0]) for b in levels_per_var]
    return mu, probs, shifts


def make_synthetic_arrays(N, d, levels_per_var, seed=0, n_types=25, lo=0, hi=None):
    """Cells [lo, hi) of the N-cell dataset: (Z float32 (hi-lo) x d, codes int32 V x (hi-lo))."""
    hi = N if hi is None else hi
    mu, probs, shifts = synthetic_params(d, levels_per_var, seed, n_types)
    V = len(levels_per_var)
    Z = np.empty((hi - lo, d), dtype=np.float32)
    codes = np.empty((V, hi - lo), dtype=np.int32)
    for c in range(lo // CHUNK, (max(hi, lo + 1) - 1) // CHUNK + 1):
        c0, c1 = c * CHUNK, min(N, (c + 1) * CHUNK)
        n = c1 - c0
        rng = np.random.default_rng([int(seed), c + 1])
        t = rng.integers(0, n_types, size=n)
        z = mu[t] + rng.standard_normal((n, d))
        cb = np.empty((V, n), dtype=np.int32)
        for v in range(V):
            b = rng.choice(int(levels_per_var[v]), size=n, p=probs[v])
            cb[v] = b
            z += shifts[v][b]
        a, b_ = max(lo, c0), min(hi, c1)
        if a < b_:
            Z[a - lo:b_ - lo] = z[a - c0:b_ - c0].astype(np.float32)
            codes[:, a - lo:b_ - lo] = cb[:, a - c0:b_ - c0]
    return Z, codes


def level_labels(v, n_levels):
    """Zero padded so the lexical level order of pd.get_dummies equals the numeric code order."""
    return np.array([f"v{v}_{b:03d}" for b in range(int(n_levels))])


def make_synthetic(N, d, levels_per_var, seed=0, n_types=25):
    """Whole dataset plus a pandas meta_data frame with string columns ``var0, var1, ...``."""
    import pandas as pd
    Z, codes = make_synthetic_arrays(N, d, levels_per_var, seed=seed, n_types=n_types)
    cols = {f"var{v}": level_labels(v, levels_per_var[v])[codes[v]] for v in range(codes.shape[0])}
    return Z, pd.DataFrame(cols)
