"""CPU oracle for the Local Inverse Simpson Index (test infrastructure only -- see oracle/harmony_oracle.py).

Restates harmonypy/lisi.py of the reference (slowkow/harmonypy @ 4d2b63f) in NumPy for the "next" row
SURVEY.md section 8(f) rank 4 (``compute_lisi``):

  * ``knn_exact``        -- the neighbour search of lisi.py:53-57 (sklearn kd_tree there; exact brute force here,
                            same Euclidean metric, self excluded afterwards like lisi.py:56-57);
  * ``compute_simpson``  -- lisi.py:68-133, the per-cell bisection on beta until the entropy of the Gaussian
                            kernel weights equals log(perplexity), then Simpson's index of the labels under
                            those weights.  Vectorised over cells with per-cell masks; the control flow of each
                            cell (doubling / halving / bisecting, stop at |Hdiff| < tol, at most 50 tries, the
                            P_sum == 0 branch and the H == 0 -> -1 default) is the reference's, step for step;
  * ``compute_lisi``     -- lisi.py:24-65.

Pinned by tests/test_lisi_oracle.py to (a) the reference's own known-answer files data/lisi_{x,metadata,lisi}.tsv.gz
(tests/test_lisi.py:5-17, committed as tests/golden/lisi_kat.npz) and (b) outputs of the reference run in the build
container on pbmc PCs (tests/golden/make_golden_lisi.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def knn_exact(X, k, chunk=1024):
    """k nearest neighbours of every row of X among the rows of X (Euclidean, the row itself included, sorted by
    distance then index).  Returns (distances [N, k], indices [N, k]) like NearestNeighbors.kneighbors (lisi.py:53-54)."""
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    N = X.shape[0]
    dist = np.empty((N, k), dtype=np.float64)
    idx = np.empty((N, k), dtype=np.int64)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        diff = X[s:e, None, :] - X[None, :, :]
        d2 = np.einsum("ijk,ijk->ij", diff, diff)
        part = np.argpartition(d2, k - 1, axis=1)[:, :k]
        pd2 = np.take_along_axis(d2, part, axis=1)
        order = np.lexsort((part, pd2), axis=1)
        idx[s:e] = np.take_along_axis(part, order, axis=1)
        dist[s:e] = np.sqrt(np.take_along_axis(pd2, order, axis=1))
    return dist, idx


def _entropy_step(D, beta):
    """lisi.py:85-93 / :113-121 for all columns at once.  D: [k, n] distances, beta: [n].  Returns H [n], P [k, n]."""
    P = np.exp(-D * beta[None, :])
    P_sum = P.sum(axis=0)
    zero = P_sum == 0
    safe = np.where(zero, 1.0, P_sum)
    H = np.where(zero, 0.0, np.log(safe) + beta * (D * P).sum(axis=0) / safe)
    P = np.where(zero[None, :], 0.0, P / safe[None, :])
    return H, P


def compute_simpson(distances, indices, codes, n_categories, perplexity, tol=1e-5):
    """lisi.py:68-133.  distances, indices: [k, n] (neighbour-major like the reference's transposed arguments,
    lisi.py:63); codes: integer category of every cell (pd.Categorical codes)."""
    D = np.asarray(distances, dtype=np.float64)
    n = D.shape[1]
    logU = np.log(perplexity)
    beta = np.ones(n)
    betamin = np.full(n, -np.inf)
    betamax = np.full(n, np.inf)
    H, P = _entropy_step(D, beta)                              # :85-93
    Hdiff = H - logU
    for _ in range(50):                                        # :95-96
        active = np.abs(Hdiff) >= tol                          # :98-99 (a cell that stopped stays as it is)
        if not active.any():
            break
        up = active & (Hdiff > 0)                              # :101-106
        dn = active & ~(Hdiff > 0)                             # :107-112
        betamin = np.where(up, beta, betamin)
        betamax = np.where(dn, beta, betamax)
        with np.errstate(invalid="ignore"):
            new_up = np.where(np.isfinite(betamax), (beta + betamax) / 2, beta * 2)
            new_dn = np.where(np.isfinite(betamin), (beta + betamin) / 2, beta / 2)
        beta = np.where(up, new_up, np.where(dn, new_dn, beta))
        H2, P2 = _entropy_step(D, beta)                        # :113-121
        H = np.where(active, H2, H)
        P = np.where(active[None, :], P2, P)
        Hdiff = np.where(active, H - logU, Hdiff)
    simpson = np.where(H == 0, -1.0, 0.0)                      # :123-125
    lab = np.asarray(codes)[np.asarray(indices)]               # [k, n] category of every neighbour (:128-129)
    for c in range(n_categories):                              # :127-132
        s = np.where(lab == c, P, 0.0).sum(axis=0)
        simpson = simpson + s * s
    return simpson


def compute_lisi(X, metadata, label_colnames, perplexity=30):
    """lisi.py:24-65.  metadata: pandas DataFrame (or mapping of column name -> labels)."""
    import pandas as pd
    X = np.asarray(X, dtype=np.float64)
    n_cells = X.shape[0]
    distances, indices = knn_exact(X, int(perplexity * 3))     # :53-54
    indices = indices[:, 1:]                                   # :56-57
    distances = distances[:, 1:]
    out = np.zeros((n_cells, len(label_colnames)))
    for i, label in enumerate(label_colnames):
        labels = pd.Categorical(metadata[label])               # :61
        simpson = compute_simpson(distances.T, indices.T, labels.codes, len(labels.categories), perplexity)
        out[:, i] = 1 / simpson                                # :64
    return out
