"""CPU oracle for the device-side centroid initialisation (test infrastructure only).

NumPy restatement of harmonypy_b200/csrc/hmy_kmeans_init.cuh -- the OPTIONAL replacement of the reference's sklearn
call (harmony.py:369-373: ``KMeans(n_clusters=K, init='k-means++', n_init=1, max_iter=25)`` on the unit-length cells).
There is no reference output to pin this to (sklearn's random stream cannot be reproduced on the device, which is why
the parity configurations keep sklearn); tests/test_kmeans_init_oracle.py checks that it is a faithful k-means++ /
Lloyd (inertia on par with sklearn's on the reference's own data) and the GPU tests compare the kernels with it.

Random numbers: counter-based, u(seed, step, cell id) from three rounds of splitmix64, identical to the CUDA code.
The centre of step c is argmin_i E_i / D2_i with E_i = -log(u) (an exponential race = sampling proportional to D2_i).
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed, step, ids):
    """(0, 1] uniforms keyed by (seed, step, id): hmy_kmi_uniform."""
    with np.errstate(over="ignore"):
        h = splitmix64(splitmix64(splitmix64(np.uint64(seed)) + np.uint64(step)) ^ np.asarray(ids, dtype=np.uint64))
    return ((h >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)


def _d2(Z, c):
    diff = Z - c[None, :]
    return np.einsum("ij,ij->i", diff, diff, dtype=np.float32)


def kmeanspp_seed(Zc, K, seed):
    """Indices of the K seeding cells (k_kmpp_pass)."""
    Zc = np.asarray(Zc, dtype=np.float32)
    N = Zc.shape[0]
    ids = np.arange(N, dtype=np.uint64)
    chosen = [int(splitmix64(np.uint64(seed)) % np.uint64(N))]
    mind2 = _d2(Zc, Zc[chosen[0]])
    for c in range(1, K):
        if c > 1:
            mind2 = np.minimum(mind2, _d2(Zc, Zc[chosen[-1]]))
        pos = mind2 > 0
        if not pos.any():
            raise ValueError("fewer distinct cells than clusters")
        key = np.full(N, np.inf, dtype=np.float32)
        key[pos] = (-np.log(uniform(seed, c, ids[pos])) / mind2[pos].astype(np.float64)).astype(np.float32)
        chosen.append(int(np.argmin(key)))                     # first index among exact ties
    return np.array(chosen, dtype=np.int64)


def lloyd(Zc, C0, max_iter=25, tol=1e-4):
    """k_lloyd_assign / k_lloyd_update.  Returns (centres, iterations, inertia of the last assignment, last shift^2)."""
    Zc = np.asarray(Zc, dtype=np.float32)
    C = np.array(C0, dtype=np.float32)
    N, d = Zc.shape
    mean = Zc.astype(np.float64).mean(axis=0)
    var = (1.0 - float(mean @ mean)) / d                        # rows have unit length
    it, inertia, shift = 0, 0.0, 0.0
    for it in range(1, max_iter + 1):
        score = (C * C).sum(axis=1, dtype=np.float32)[None, :] - np.float32(2.0) * (Zc @ C.T)
        lab = np.argmin(score, axis=1)
        inertia = float((score[np.arange(N), lab].astype(np.float64) + (Zc.astype(np.float64) ** 2).sum(axis=1)).sum())
        sums = np.zeros((C.shape[0], d), dtype=np.float64)
        np.add.at(sums, lab, Zc.astype(np.float64))
        cnt = np.bincount(lab, minlength=C.shape[0])
        newC = C.copy()
        nz = cnt > 0
        newC[nz] = (sums[nz] / cnt[nz, None]).astype(np.float32)
        shift = float(((newC.astype(np.float64) - C.astype(np.float64)) ** 2).sum())
        C = newC
        if shift <= tol * var:
            break
    return C, it, inertia, shift


def kmeans_init(Z, K, seed, max_iter=25, tol=1e-4):
    """Cells (any scale) -> unit rows -> k-means++ seeding -> Lloyd: what Harmony(init_mode='device') starts from."""
    Z = np.asarray(Z, dtype=np.float32)
    Zc = Z / np.linalg.norm(Z, axis=1, keepdims=True)
    seeds = kmeanspp_seed(Zc, K, seed)
    if max_iter == 0:
        return Zc[seeds].copy(), dict(iterations=0, inertia=0.0, last_shift2=0.0, seeds=seeds)
    C, it, inertia, shift = lloyd(Zc, Zc[seeds], max_iter, tol)
    return C, dict(iterations=it, inertia=inertia, last_shift2=shift, seeds=seeds)
