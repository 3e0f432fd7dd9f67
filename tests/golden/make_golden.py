#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [pbmc] [ircolitis] [synth]

What it does, per dataset:
  * imports slowkow/harmonypy from /root/reference (unmodified) and runs
    ``run_harmony(..., device='cpu')`` with torch pinned to 8 threads (SURVEY.md 8c:
    never 1 thread), capturing -- without touching the reference's code -- the sklearn
    centroids (harmony.py:370-373), every ``torch.randperm`` (harmony.py:471), and the
    state after ``init_cluster`` / each ``cluster()`` / each ``moe_correct_ridge()``;
  * runs the fp64 arbiter: the same source text with float32->float64 substituted
    on the fly into a temp dir (never written into this repo), same inputs rounded
    through fp32, same centroids and the same permutations;
  * writes ``<name>_input.npz`` (inputs the tests feed to the oracle / CUDA engine) and
    ``<name>_golden.npz`` (what the reference produced).

The fixtures are small on purpose: per-iteration snapshots are subsampled
(``stage_cells``), the final Z_corr is stored on ``final_cells`` plus full-matrix
column sums.  pbmc final Z_corr is stored in full.
"""
import hashlib
import importlib.util
import os
import re
import sys
import tempfile

sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import numpy as np
import pandas as pd
import torch

torch.set_num_threads(8)


def perm_digest(p):
    return hashlib.sha1(np.asarray(p, dtype=np.int64).tobytes()).hexdigest()[:16]


def load_dataset(name):
    if name == "pbmc":
        meta = pd.read_csv(f"{REF}/data/pbmc_3500_meta.tsv.gz", sep="\t")
        pcs = pd.read_csv(f"{REF}/data/pbmc_3500_pcs.tsv.gz", sep="\t")
        return pcs, meta[["donor"]], ["donor"], {}
    if name == "ircolitis":
        pcs = pd.read_csv(f"{REF}/data/ircolitis_blood_cd8_pcs.tsv.gz", sep="\t").iloc[:, 1:]
        meta = pd.read_csv(f"{REF}/data/ircolitis_blood_cd8_obs.tsv.gz", sep="\t", usecols=["batch"])
        return pcs, meta, ["batch"], {}
    if name == "synth":
        # small two-covariate case with non-default theta/lambda/tau: exercises V>1,
        # per-covariate theta, dynamic lambda (lamb=-1) and tau (harmony.py:137-173)
        from harmonypy_b200.synthetic import make_synthetic
        Z, meta = make_synthetic(6000, 20, [5, 3], seed=7)
        return pd.DataFrame(Z), meta, list(meta.columns), dict(
            theta=[2.0, 1.0], lamb=-1, tau=5, nclust=40, max_iter_harmony=4, sigma=0.12)
    raise SystemExit(f"unknown dataset {name}")


def load_reference_module(fp64):
    """Import the reference's harmony.py; for fp64 substitute dtypes in a temp copy."""
    if not fp64:
        import harmonypy.harmony as mod
        return mod
    src = open(f"{REF}/harmonypy/harmony.py").read()
    src = src.replace("torch.float32", "torch.float64").replace("np.float32", "np.float64")
    src = src.replace("model.fit(Z_cos_np.T)", 'model.fit(Z_cos_np.T.astype("float32"))')
    tmp = tempfile.mkdtemp(prefix="hmy64_")
    path = os.path.join(tmp, "harmony64.py")
    with open(path, "w") as f:
        f.write(src)
    spec = importlib.util.spec_from_file_location("harmony64", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_captured(mod, pcs, meta, vars_use, kwargs, stage_cells, replay=None):
    """Run mod.run_harmony with capture.  If ``replay`` is given (dict with 'Y0' and
    'perms'), force those centroids / permutations instead of drawing new ones."""
    cap = {"perms": [], "stages": [], "Y0": None}
    H = mod.Harmony
    orig_cluster, orig_ridge, orig_init = H.cluster, H.moe_correct_ridge, H.init_cluster
    orig_randperm = torch.randperm
    OrigKMeans = mod.KMeans

    class CapKMeans(OrigKMeans):
        def fit(self, X, *a, **k):
            if replay is not None:
                self.cluster_centers_ = np.asarray(replay["Y0"], dtype=np.float64)
                return self
            out = super().fit(X, *a, **k)
            cap["Y0"] = np.array(self.cluster_centers_, dtype=np.float32)
            return out

    def randperm(n, *a, **k):
        if replay is not None:
            p = torch.as_tensor(replay["perms"][len(cap["perms"])])
        else:
            p = orig_randperm(n, *a, **k)
        cap["perms"].append(p.numpy().copy())
        return p

    def snap(self, name):
        s = {"name": name,
             "Y": self._Y.numpy().copy(), "O": self._O.numpy().copy(), "E": self._E.numpy().copy(),
             "R_sub": self._R[:, stage_cells].numpy().T.copy(),
             "Zcorr_sub": self._Z_corr[:, stage_cells].numpy().T.copy(),
             "n_obj": len(self.objective_kmeans)}
        cap["stages"].append(s)

    def init_cluster(self, rs):
        orig_init(self, rs); snap(self, "init")

    def cluster(self):
        orig_cluster(self); snap(self, "cluster")

    def ridge(self):
        orig_ridge(self); snap(self, "ridge")

    H.init_cluster, H.cluster, H.moe_correct_ridge = init_cluster, cluster, ridge
    mod.KMeans = CapKMeans
    torch.randperm = randperm
    try:
        ho = mod.run_harmony(pcs, meta, vars_use, verbose=False, device="cpu", **kwargs)
    finally:
        H.init_cluster, H.cluster, H.moe_correct_ridge = orig_init, orig_cluster, orig_ridge
        mod.KMeans = OrigKMeans
        torch.randperm = orig_randperm
    return ho, cap


def main(names):
    for name in names:
        pcs, meta, vars_use, kwargs = load_dataset(name)
        N = meta.shape[0]
        step_stage = max(1, N // 512)
        step_final = 1 if N <= 8000 else 8
        stage_cells = np.arange(0, N, step_stage)
        final_cells = np.arange(0, N, step_final)

        mod32 = load_reference_module(False)
        ho, cap = run_captured(mod32, pcs, meta, vars_use, kwargs, stage_cells)
        print(name, "fp32 rounds", ho.kmeans_rounds, "obj_h[-1]", ho.objective_harmony[-1])

        # Inputs exactly as the reference's front end derived them (harmony.py:116-173).
        Z = np.asarray(pcs.values if hasattr(pcs, "values") else pcs, dtype=np.float32)
        if Z.shape[0] != N:
            Z = Z.T
        codes = np.stack([pd.Categorical(meta[v]).codes for v in vars_use]).astype(np.int32)
        levels = np.array([len(pd.Categorical(meta[v]).categories) for v in vars_use], dtype=np.int32)
        # sanity: codes reproduce the reference's one-hot (pd.get_dummies column order)
        from oracle.harmony_oracle import onehot_from_codes
        assert np.array_equal(onehot_from_codes(codes, levels), ho.Phi.T), "code/one-hot order mismatch"

        np.savez_compressed(
            os.path.join(HERE, f"{name}_input.npz"),
            Z=Z, codes=codes, levels=levels, Y0=cap["Y0"],
            Pr_b=ho.Pr_b, theta=ho.theta, sigma=ho.sigma, lamb=ho.lamb,
            alpha=np.float64(ho.alpha), lambda_estimation=np.bool_(ho.lambda_estimation),
            block_size=np.float64(ho.block_size), K=np.int32(ho.K),
            max_iter_harmony=np.int32(ho.max_iter_harmony), max_iter_kmeans=np.int32(ho.max_iter_kmeans),
            epsilon_kmeans=np.float64(ho.epsilon_kmeans), epsilon_harmony=np.float64(ho.epsilon_harmony),
            random_state=np.int32(0),
            run_kwargs=np.array(repr(kwargs)),
            perm_digests=np.array([perm_digest(p) for p in cap["perms"]]),
            perm0_head=cap["perms"][0][:32].astype(np.int64),
        )

        # fp64 arbiter on identical centroids and permutations
        mod64 = load_reference_module(True)
        pcs64 = pd.DataFrame(Z.astype(np.float64))
        ho64, cap64 = run_captured(mod64, pcs64, meta, vars_use, kwargs, stage_cells,
                                   replay={"Y0": cap["Y0"], "perms": cap["perms"]})
        print(name, "fp64 rounds", ho64.kmeans_rounds)
        Zc32, Zc64 = ho.Z_corr, ho64.Z_corr
        rel = np.abs(Zc32 - Zc64).max() / np.abs(Zc64).max()
        print(name, "ref fp32(8T) vs fp64 arbiter: max|d|/max|Z| = %.3e" % rel)

        g = {}
        for i, s in enumerate(cap["stages"]):
            for key in ("Y", "O", "E", "Zcorr_sub"):
                g[f"s{i}_{key}"] = s[key].astype(np.float32)
            if s["name"] == "init" or i in (1, len(cap["stages"]) - 2):
                g[f"s{i}_R_sub"] = s["R_sub"].astype(np.float32)
        for i, s in enumerate(cap64["stages"]):
            g[f"s{i}_O_f64"] = s["O"].astype(np.float64)
            g[f"s{i}_Y_f64"] = s["Y"].astype(np.float64)
        np.savez_compressed(
            os.path.join(HERE, f"{name}_golden.npz"),
            stage_names=np.array([s["name"] for s in cap["stages"]]),
            stage_n_obj=np.array([s["n_obj"] for s in cap["stages"]], dtype=np.int32),
            stage_cells=stage_cells.astype(np.int64), final_cells=final_cells.astype(np.int64),
            kmeans_rounds=np.array(ho.kmeans_rounds, dtype=np.int32),
            kmeans_rounds_f64=np.array(ho64.kmeans_rounds, dtype=np.int32),
            objective_harmony=np.array(ho.objective_harmony, dtype=np.float64),
            objective_kmeans=np.array(ho.objective_kmeans, dtype=np.float64),
            objective_kmeans_dist=np.array(ho.objective_kmeans_dist, dtype=np.float64),
            objective_kmeans_entropy=np.array(ho.objective_kmeans_entropy, dtype=np.float64),
            objective_kmeans_cross=np.array(ho.objective_kmeans_cross, dtype=np.float64),
            objective_kmeans_f64=np.array(ho64.objective_kmeans, dtype=np.float64),
            Zcorr_final=Zc32[final_cells].astype(np.float32),
            Zcorr_final_f64=Zc64[final_cells].astype(np.float64),
            Zcorr_colsum=Zc32.astype(np.float64).sum(axis=0),
            Zcorr_absmax=np.float64(np.abs(Zc32).max()),
            ref_f32_vs_f64=np.float64(rel),
            # parameters as the fp64 arbiter saw them (sigma=0.1 is not an fp32 number)
            f64_Pr_b=ho64.Pr_b.astype(np.float64), f64_theta=ho64.theta.astype(np.float64),
            f64_sigma=ho64.sigma.astype(np.float64), f64_lamb=ho64.lamb.astype(np.float64),
            **g,
        )
        for f in (f"{name}_input.npz", f"{name}_golden.npz"):
            print("  wrote", f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or ["pbmc", "synth", "ircolitis"])
