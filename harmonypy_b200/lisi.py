"""compute_lisi with the reference's signature (harmonypy/lisi.py:24-65), evaluated by the CUDA library.

    lisi = harmonypy_b200.compute_lisi(X, metadata, label_colnames, perplexity=30)      # n_cells x n_labels

Host side: category codes per label column (``pd.Categorical(metadata[label])``, lisi.py:61) and one call of
``hmy_lisi_compute`` (include/harmony_b200.h): exact fp64 nearest neighbours, the per-cell perplexity bisection and
the inverse Simpson index all run on the GPU.  No CPU fallback: without the library or a GPU this raises.

The kernels were written at the end of round 1 without GPU time left (see csrc/hmy_lisi.cu); until they have been
validated against oracle/lisi_oracle.py on hardware their GPU tests are opt-in (HMY_TEST_LISI=1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


def compute_lisi(X, metadata, label_colnames, perplexity=30, device=0):
    import pandas as pd
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim != 2:
        raise ValueError("X must be a cells x features matrix")
    n, d = X.shape
    label_colnames = list(label_colnames)
    if metadata.shape[0] != n:
        raise ValueError("metadata must have one row per row of X")
    codes = np.empty((len(label_colnames), n), dtype=np.int32)
    for i, label in enumerate(label_colnames):
        cat = pd.Categorical(metadata[label])                  # lisi.py:61
        if (cat.codes < 0).any():
            raise ValueError(f"label column {label!r} has missing values")
        codes[i] = cat.codes
    out = np.empty((n, len(label_colnames)), dtype=np.float64)
    lib = _cabi.load()
    rc = lib.hmy_lisi_compute(int(device), n, d, X.ctypes.data_as(C.c_void_p), len(label_colnames),
                              codes.ctypes.data_as(C.c_void_p), float(perplexity), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise _cabi.EngineError(lib.hmy_last_error(None).decode())
    return out
