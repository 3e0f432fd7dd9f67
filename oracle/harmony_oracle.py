"""CPU oracle for the Harmony inner loop (TEST INFRASTRUCTURE -- not a product path).

This module is a NumPy restatement of the algorithm that slowkow/harmonypy v0.2.0
implements with torch ops in ``harmonypy/harmony.py``.  It exists only so that the
CUDA engine in ``harmonypy_b200`` can be checked against an independent statement of
the same arithmetic.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; nothing
under ``harmonypy_b200/`` does, and the product path raises if its CUDA library is
missing instead of falling back to this file.

Parity pin: ``tests/golden/make_golden.py`` runs the real reference (imported from
``/root/reference``, CPU fp32, 8 threads) with stage capture and commits the
snapshots under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays the same
inputs / centroids / permutations through this oracle and compares every stage.
The numbers that pin came out at are recorded in DESIGN.md ("Oracle pin").

Every function cites the reference lines (relative to /root/reference/) it follows.

Layout follows the reference (so BLAS sees the same shapes): Z is d x N, R is K x N,
Phi is B x N one-hot, O/E are K x B, Y is d x K.  ``dtype`` selects fp32 (the
reference's arithmetic) or fp64 (the arbiter run described in SURVEY.md section 8c).
"""
from __future__ import annotations

import numpy as np


def onehot_from_codes(codes: np.ndarray, levels_per_var, dtype=np.float32) -> np.ndarray:
    """Dense B x N indicator from V x N integer level codes (B = sum(levels)).

    Row order is covariate-major then level, which is the column order of
    ``pd.get_dummies(meta[vars_use])`` (harmonypy/harmony.py:133).
    """
    codes = np.atleast_2d(np.asarray(codes))
    V, N = codes.shape
    B = int(np.sum(levels_per_var))
    phi = np.zeros((B, N), dtype=dtype)
    off = 0
    cols = np.arange(N)
    for v in range(V):
        phi[off + codes[v], cols] = 1
        off += int(levels_per_var[v])
    return phi


def block_bounds(N: int, block_size: float):
    """Block boundaries of update_R (harmonypy/harmony.py:474-475, :483-484)."""
    n_blocks = int(np.ceil(1.0 / block_size))
    cells_per_block = int(N * block_size)
    bounds = []
    for blk in range(n_blocks):
        lo = blk * cells_per_block
        hi = N if blk == n_blocks - 1 else (blk + 1) * cells_per_block
        bounds.append((lo, hi))
    return bounds


class HarmonyOracle:
    """State + stages of the Harmony loop, one method per reference method.

    Parameters mirror ``Harmony.__init__`` (harmonypy/harmony.py:224-271) except that
    the initial centroids are passed in (``init_from_centroids``) instead of being
    computed by sklearn here, and the block permutations are supplied by the caller
    (``perm_source()`` must return the next ``randperm(N)`` as an integer array) so the
    RNG stream is owned by the test, not by the oracle.
    """

    def __init__(self, Z_dxN, phi_BxN, Pr_b, sigma, theta, lamb, alpha=0.2,
                 lambda_estimation=False, block_size=0.05, epsilon_kmeans=1e-5,
                 epsilon_harmony=1e-4, max_iter_kmeans=20, dtype=np.float32):
        ft = np.dtype(dtype).type
        self.ft = ft
        self.Z_orig = np.array(Z_dxN, dtype=ft)
        self.Z_corr = self.Z_orig.copy()
        # harmony.py:238 -- unit-length cells
        self.Z_cos = self.Z_orig / np.linalg.norm(self.Z_orig, axis=0)
        self.Phi = np.array(phi_BxN, dtype=ft)
        self.Pr_b = np.array(Pr_b, dtype=ft)
        self.d, self.N = self.Z_orig.shape
        self.B = self.Phi.shape[0]
        # harmony.py:249-256 -- per-level cell lists and the design with intercept
        self.level_cells = [np.nonzero(self.Phi[b] > 0)[0] for b in range(self.B)]
        self.Phi_moe = np.vstack([np.ones((1, self.N), dtype=ft), self.Phi])
        self.sigma = np.array(sigma, dtype=ft)
        self.theta = np.array(theta, dtype=ft)
        self.lamb = np.array(lamb, dtype=ft)
        self.alpha = alpha
        self.lambda_estimation = lambda_estimation
        self.block_size = block_size
        self.K = self.sigma.shape[0]
        self.window_size = 3                      # harmony.py:258
        self.epsilon_kmeans = epsilon_kmeans
        self.epsilon_harmony = epsilon_harmony
        self.max_iter_kmeans = max_iter_kmeans
        self.objective_harmony = []
        self.objective_kmeans = []
        self.objective_kmeans_dist = []
        self.objective_kmeans_entropy = []
        self.objective_kmeans_cross = []
        self.kmeans_rounds = []
        self.R = None
        self.Y = None
        self.O = None
        self.E = None
        self.dist = None

    # ------------------------------------------------------------------ helpers
    def _cosine_dist(self):
        """dist = 2 (1 - Y^T Z_cos)  (harmony.py:380, :447)."""
        return (2 * (1 - self.Y.T @ self.Z_cos)).astype(self.ft, copy=False)

    @staticmethod
    def _unit_columns(M):
        """Column L2 normalisation (harmony.py:377, :444, :569)."""
        return M / np.linalg.norm(M, axis=0)

    # ------------------------------------------------------------------ a2
    def init_from_centroids(self, Y_dxK):
        """Tail of init_cluster after the sklearn call (harmony.py:373-392)."""
        self.Y = self._unit_columns(np.array(Y_dxK, dtype=self.ft))
        self.dist = self._cosine_dist()
        A = np.exp(-self.dist / self.sigma[:, None])          # :383-384
        self.R = A / A.sum(axis=0)                             # :385
        self.E = np.outer(self.R.sum(axis=1), self.Pr_b)      # :388
        self.O = self.R @ self.Phi.T                           # :389
        self.compute_objective()                               # :391
        self.objective_harmony.append(self.objective_kmeans[-1])   # :392

    # ------------------------------------------------------------------ a5
    def compute_objective(self):
        """Three-term objective (harmony.py:394-417)."""
        ft = self.ft
        c0 = 2000.0 / self.N                                   # :396
        # the three sums are accumulated in float64 (the values are `ft`): convergence is a 1e-5
        # threshold on these sums and must not depend on NumPy's fp32 summation order
        err = float(np.sum(self.R * self.dist, dtype=np.float64))                # :399
        with np.errstate(divide="ignore", invalid="ignore"):
            h = self.R * np.log(self.R)                        # :572-576
        h = np.where(np.isfinite(h), h, ft(0))
        ent = float(np.sum(h * self.sigma[:, None], dtype=np.float64))           # :402
        Oc = np.maximum(self.O, ft(1e-8))                      # :407
        Ec = np.maximum(self.E, ft(1e-8))                      # :408
        tl = self.theta[None, :] * np.log((Oc + Ec) / Ec)      # :409-410
        cross = float(np.sum((self.R * self.sigma[:, None]) * (tl @ self.Phi), dtype=np.float64))  # :405, :411
        self.objective_kmeans.append((err + ent + cross) * c0)
        self.objective_kmeans_dist.append(err * c0)
        self.objective_kmeans_entropy.append(ent * c0)
        self.objective_kmeans_cross.append(cross * c0)

    # ------------------------------------------------------------------ a4
    def update_R(self, perm):
        """Blockwise Gauss-Seidel assignment update (harmony.py:464-513).

        ``perm`` is the ``randperm(N)`` the reference draws at :471.  Instead of
        physically permuting the matrices (:478-480, :512-513) the blocks index the
        cells directly; the arithmetic per block is unchanged.
        """
        ft = self.ft
        S = np.exp(-self.dist / self.sigma[:, None])           # :466-467
        S = S / S.sum(axis=0)                                  # :468
        perm = np.asarray(perm)
        for lo, hi in block_bounds(self.N, self.block_size):
            cells = perm[lo:hi]
            Rb = self.R[:, cells]
            Pb = self.Phi[:, cells]
            # take the block out of the running statistics (:491-492)
            self.E -= np.outer(Rb.sum(axis=1), self.Pr_b)
            self.O -= Rb @ Pb.T
            # diversity penalty (:495-499, :579-584)
            denom = np.maximum(self.O + self.E, ft(1e-8))
            ratio = np.clip(self.E / denom, ft(1e-8), ft(1.0))
            pen = np.power(ratio, self.theta[None, :]).astype(ft, copy=False)
            # re-assign (:500-503)
            Rn = S[:, cells] * (pen @ Pb)
            Rn = Rn / np.maximum(Rn.sum(axis=0), ft(1e-8))
            # put the block back (:506-507, :509)
            self.E += np.outer(Rn.sum(axis=1), self.Pr_b)
            self.O += Rn @ Pb.T
            self.R[:, cells] = Rn

    # ------------------------------------------------------------------ a6
    def check_convergence(self, i_type):
        """harmony.py:515-533 (window of 3 for k-means; signed test for harmony)."""
        if i_type == 0:
            w = self.window_size
            if len(self.objective_kmeans) <= w + 1:
                return False
            old = sum(self.objective_kmeans[-w - 1:-1])
            new = sum(self.objective_kmeans[-w:])
            return abs(old - new) / abs(old) < self.epsilon_kmeans
        if i_type == 1:
            if len(self.objective_harmony) < 2:
                return False
            old = self.objective_harmony[-2]
            new = self.objective_harmony[-1]
            return (old - new) / abs(old) < self.epsilon_harmony
        return True

    # ------------------------------------------------------------------ a3
    def kmeans_round(self, perm):
        """One pass of the loop body of cluster() (harmony.py:443-453)."""
        self.Y = self._unit_columns(self.Z_cos @ self.R.T)     # :443-444
        self.dist = self._cosine_dist()                        # :447
        self.update_R(perm)                                    # :450
        self.compute_objective()                               # :453

    def cluster(self, perm_source):
        """harmony.py:437-462.  (:438 is dead work and is skipped.)"""
        rounds = 0
        for i in range(self.max_iter_kmeans):
            self.kmeans_round(perm_source())
            rounds = i + 1
            if i > self.window_size and self.check_convergence(0):   # :455-458
                break
        self.kmeans_rounds.append(rounds)                      # :461
        self.objective_harmony.append(self.objective_kmeans[-1])   # :462

    # ------------------------------------------------------------------ a7
    def moe_correct_ridge(self):
        """Per-cluster ridge regression and correction (harmony.py:535-569)."""
        ft = self.ft
        Zc = self.Z_orig.copy()                                # :537
        for k in range(self.K):
            if self.lambda_estimation:                         # :541-544, :587-591
                lam = np.concatenate([np.zeros(1, ft), self.E[k] * ft(self.alpha)]).astype(ft)
            else:
                lam = self.lamb
            Phi_Rk = self.Phi_moe * self.R[k]                  # :547
            cov = Phi_Rk @ self.Phi_moe.T + np.diag(lam)       # :550
            inv = np.linalg.inv(cov).astype(ft, copy=False)    # :553
            Zk = self.Z_orig * self.R[k]                       # :556
            W = np.outer(inv[:, 0], Zk.sum(axis=1))            # :559
            for b in range(self.B):                            # :561-563
                W = W + np.outer(inv[:, b + 1], Zk[:, self.level_cells[b]].sum(axis=1))
            W[0, :] = 0                                        # :565
            Zc = Zc - W.T @ Phi_Rk                             # :566
        self.Z_corr = Zc.astype(ft, copy=False)
        self.Z_cos = self._unit_columns(self.Z_corr)           # :569

    # ------------------------------------------------------------------ driver
    def harmonize(self, max_iter_harmony, perm_source, on_stage=None):
        """harmony.py:419-435.  ``on_stage(name, it, self)`` is a test hook."""
        converged = False
        for it in range(1, max_iter_harmony + 1):
            self.cluster(perm_source)
            if on_stage:
                on_stage("cluster", it, self)
            self.moe_correct_ridge()
            if on_stage:
                on_stage("ridge", it, self)
            converged = self.check_convergence(1)
            if converged:
                break
        return converged


def torch_perm_source(N, seed):
    """The reference's permutation stream: ``torch.manual_seed(seed)`` once
    (harmony.py:200) then one CPU ``torch.randperm(N)`` per update_R (:471)."""
    import torch
    gen_state = {"seeded": False}

    def nxt():
        if not gen_state["seeded"]:
            torch.manual_seed(seed)
            gen_state["seeded"] = True
        return torch.randperm(N).numpy()
    return nxt
