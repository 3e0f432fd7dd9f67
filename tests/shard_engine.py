"""TEST-ONLY engine: a cell-sharded fp64 NumPy restatement of the staged multi-GPU algorithm
(local partial sums + all-reduce of the small tables), implementing the
same engine interface the Python host drives (harmonypy_b200._cabi.Engine).  It lets the
world_size-2 gloo tests exercise the host's N>1 path (sharding, permutation stream,
convergence decisions, gathers) on CPU and prove that "shard + sum the tables" reproduces the
single-process oracle.  Never imported by the product."""
import numpy as np

from harmonypy_b200 import _cabi


class ShardOracleEngine:
    def __init__(self, problem, lo, hi, device, comm, options):
        self.p, self.lo, self.hi, self.comm = problem, lo, hi, comm
        self.N = problem.N
        self.K, self.d = problem.K, problem.d
        self.levels = np.asarray(problem.levels)
        self.B = int(self.levels.sum())
        self.offs = np.concatenate([[0], np.cumsum(self.levels)[:-1]])
        self.calls = {"allreduce": 0}

    # ---- plumbing
    def _sum(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if self.comm is not None and self.comm.world > 1:
            self.comm.allreduce_numpy(a)
            self.calls["allreduce"] += 1
        return a

    def set_option(self, k, v):
        pass

    def counter(self, name):
        return self.calls.get(name, 0)

    def set_params(self, Pr_b, theta, sigma, lamb, lambda_estimation, alpha, block_size):
        f = np.float64
        self.Pr_b, self.theta, self.sigma = np.asarray(Pr_b, f), np.asarray(theta, f), np.asarray(sigma, f)
        self.lamb = None if lamb is None else np.asarray(lamb, f)
        self.lambda_estimation, self.alpha, self.block_size = bool(lambda_estimation), float(alpha), float(block_size)

    def set_data(self, Z, codes):
        self.Z = np.asarray(Z, np.float64)
        self.codes = np.asarray(codes)
        self.n = self.Z.shape[0]
        self.Zcorr = self.Z.copy()
        self.Zcos = self.Z / np.linalg.norm(self.Z, axis=1, keepdims=True)
        self.phi = np.zeros((self.n, self.B))
        for v in range(len(self.levels)):
            self.phi[np.arange(self.n), self.offs[v] + self.codes[v]] = 1

    # ---- stages
    def _E(self):
        return np.outer(self.O[:, :self.levels[0]].sum(axis=1), self.Pr_b)

    def _objective(self, R, dist):
        err = np.sum(R * dist)
        with np.errstate(divide="ignore", invalid="ignore"):
            h = R * np.log(R)
        ent = np.sum(np.where(np.isfinite(h), h, 0.0) * self.sigma[None, :])
        err, ent = self._sum(np.array([err, ent]))
        Oc, Ec = np.maximum(self.O, 1e-8), np.maximum(self._E(), 1e-8)
        cross = np.sum(self.sigma[:, None] * self.theta[None, :] * np.log((Oc + Ec) / Ec) * self.O)
        return float(err), float(ent), float(cross)

    def _centroids(self):
        Y = self._sum(self.R.T @ self.Zcos)
        self.Ynext = Y / np.linalg.norm(Y, axis=1, keepdims=True)

    def init_from_centroids(self, Y0):
        Y = np.asarray(Y0, np.float64)
        self.Yuse = Y / np.linalg.norm(Y, axis=1, keepdims=True)
        dist = 2 * (1 - self.Zcos @ self.Yuse.T)
        A = np.exp(-dist / self.sigma[None, :])
        self.R = A / A.sum(axis=1, keepdims=True)
        self.O = self._sum(self.R.T @ self.phi)
        self._centroids()
        return self._objective(self.R, dist)

    def kmeans_round(self, perm):
        assert perm is not None, "the test engine only supports the reference permutation stream"
        self.Yuse = self.Ynext
        dist = 2 * (1 - self.Zcos @ self.Yuse.T)
        S = np.exp(-dist / self.sigma[None, :])
        S = S / S.sum(axis=1, keepdims=True)
        n_blocks = int(np.ceil(1.0 / self.block_size))
        cpb = int(self.N * self.block_size)
        members = []
        for b in range(n_blocks):
            g = perm[b * cpb:(self.N if b == n_blocks - 1 else (b + 1) * cpb)]
            g = g[(g >= self.lo) & (g < self.hi)] - self.lo
            members.append(g)
        Told = self._sum(np.stack([self.R[m].T @ self.phi[m] for m in members]))    # one all-reduce
        for b, m in enumerate(members):
            self.O = self.O - Told[b]
            E = self._E()
            ratio = np.clip(E / np.maximum(self.O + E, 1e-8), 1e-8, 1.0)
            pen = ratio ** self.theta[None, :]
            Rn = S[m] * (self.phi[m] @ pen.T)
            Rn = Rn / np.maximum(Rn.sum(axis=1, keepdims=True), 1e-8)
            self.R[m] = Rn
            self.O = self.O + self._sum(Rn.T @ self.phi[m])                           # one per block
        self._centroids()
        return self._objective(self.R, dist)

    def ridge_correct(self):
        phim = np.hstack([np.ones((self.n, 1)), self.phi])
        gram = self._sum(np.einsum("nk,ni,nj->kij", self.R, phim, phim))
        mom = self._sum(np.einsum("nk,ni,nd->kid", self.R, phim, self.Z))
        E = self._E()
        Zc = self.Z.copy()
        for k in range(self.K):
            lam = np.concatenate([[0.0], self.alpha * E[k]]) if self.lambda_estimation else self.lamb
            W = np.linalg.solve(gram[k] + np.diag(lam), mom[k])
            W[0] = 0
            Zc -= (phim * self.R[:, k:k + 1]) @ W
        self.Zcorr = Zc
        self.Zcos = Zc / np.linalg.norm(Zc, axis=1, keepdims=True)
        self._centroids()

    def get(self, which):
        f = np.float32
        if which == _cabi.Z_CORR: return self.Zcorr.astype(f)
        if which == _cabi.Z_COS: return self.Zcos.astype(f)
        if which == _cabi.Z_ORIG: return self.Z.astype(f)
        if which == _cabi.R: return self.R.astype(f)
        if which == _cabi.Y: return self.Yuse.astype(f)
        if which == _cabi.O: return self.O.copy()
        if which == _cabi.E: return self._E()
        raise KeyError(which)
