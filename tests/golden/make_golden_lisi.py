#!/usr/bin/env python
"""Golden fixtures for compute_lisi (harmonypy/lisi.py), generated from the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_lisi.py

Writes
  * ``lisi_kat.npz``  -- the reference's own known-answer test (tests/test_lisi.py:5-17): data/lisi_x.tsv.gz,
    data/lisi_metadata.tsv.gz as integer codes + category names, and the expected values of data/lisi_lisi.tsv.gz
    (produced by the R package), plus what the reference's Python code returns on them here;
  * ``lisi_pbmc.npz`` -- compute_lisi of the unmodified reference on the first 10 PCs of data/pbmc_3500_pcs.tsv.gz with
    the ``donor`` labels, perplexity 30 (3500 cells: the size the oracle and a GPU kernel are compared at).
"""
import os
import sys

sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import numpy as np
import pandas as pd

from harmonypy.lisi import compute_lisi      # the reference, unmodified


def codes_of(meta, cols):
    cats = [pd.Categorical(meta[c]) for c in cols]
    return (np.stack([c.codes.astype(np.int32) for c in cats]),
            np.array([len(c.categories) for c in cats], dtype=np.int32))


def main():
    X = pd.read_csv(f"{REF}/data/lisi_x.tsv.gz", sep="\t")
    meta = pd.read_csv(f"{REF}/data/lisi_metadata.tsv.gz", sep="\t")
    want = pd.read_csv(f"{REF}/data/lisi_lisi.tsv.gz", sep="\t").iloc[:, -2:].to_numpy()
    got = compute_lisi(X, meta, meta.columns, 30)
    assert np.allclose(got, want)                      # the reference's own assertion (tests/test_lisi.py:17)
    codes, ncat = codes_of(meta, list(meta.columns))
    np.savez_compressed(os.path.join(HERE, "lisi_kat.npz"), X=X.to_numpy(np.float64), codes=codes, n_categories=ncat,
                        columns=np.array(list(meta.columns)), perplexity=np.float64(30),
                        lisi_expected=want, lisi_reference_python=got)
    print("lisi_kat: reference python vs R golden max abs diff", float(np.abs(got - want).max()))

    pcs = pd.read_csv(f"{REF}/data/pbmc_3500_pcs.tsv.gz", sep="\t").iloc[:, :10]
    pm = pd.read_csv(f"{REF}/data/pbmc_3500_meta.tsv.gz", sep="\t")
    X32 = pcs.to_numpy(np.float32)                     # the fixture stores fp32; the reference sees exactly these values
    got = compute_lisi(X32.astype(np.float64), pm, ["donor"], 30)
    codes, ncat = codes_of(pm, ["donor"])
    np.savez_compressed(os.path.join(HERE, "lisi_pbmc.npz"), X=X32, codes=codes, n_categories=ncat,
                        columns=np.array(["donor"]), perplexity=np.float64(30), lisi_reference_python=got)
    print("lisi_pbmc:", got.shape, "mean LISI", float(got.mean()))


if __name__ == "__main__":
    main()
