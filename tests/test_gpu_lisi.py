"""compute_lisi on the GPU (harmonypy_b200/lisi.py -> hmy_lisi_compute) against the reference's known-answer test,
reference outputs and the CPU oracle (validated on a B200 in round 2; runs with the plain `-m gpu` suite).
"""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = [pytest.mark.gpu]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _meta(g):
    return pd.DataFrame({str(c): pd.Categorical.from_codes(g["codes"][i], categories=list(range(int(g["n_categories"][i]))))
                         for i, c in enumerate(g["columns"])})


def test_known_answer_test_of_the_reference():
    """tests/test_lisi.py:5-17 of the reference, same criterion (np.allclose against the R package's values)."""
    import harmonypy_b200 as hm
    g = np.load(os.path.join(GOLDEN, "lisi_kat.npz"))
    got = hm.compute_lisi(g["X"], _meta(g), [str(c) for c in g["columns"]], float(g["perplexity"]))
    assert got.shape == (400, 2)
    assert np.allclose(got, g["lisi_expected"])
    np.testing.assert_allclose(got, g["lisi_reference_python"], rtol=1e-9)


def test_pbmc_pcs_against_reference_output():
    import harmonypy_b200 as hm
    g = np.load(os.path.join(GOLDEN, "lisi_pbmc.npz"))
    got = hm.compute_lisi(g["X"].astype(np.float64), _meta(g), ["donor"], 30)
    np.testing.assert_allclose(got, g["lisi_reference_python"], rtol=1e-9)


@pytest.mark.parametrize("n,d,perp", [(5000, 50, 30), (1201, 3, 10), (257, 128, 42)])
def test_against_oracle_on_random_data(n, d, perp):
    import harmonypy_b200 as hm
    from oracle.lisi_oracle import compute_lisi as oracle_lisi
    rng = np.random.default_rng(n)
    centers = rng.normal(size=(7, d)) * 3
    t = rng.integers(0, 7, n)
    X = centers[t] + rng.normal(size=(n, d))
    meta = pd.DataFrame({"type": pd.Categorical(t.astype(str)), "batch": pd.Categorical(rng.integers(0, 3, n).astype(str))})
    got = hm.compute_lisi(X, meta, ["type", "batch"], perp)
    want = oracle_lisi(X, meta, ["type", "batch"], perp)
    np.testing.assert_allclose(got, want, rtol=1e-8)


def test_argument_errors_are_loud():
    import harmonypy_b200 as hm
    from harmonypy_b200._cabi import EngineError
    X = np.random.default_rng(0).normal(size=(50, 4))
    meta = pd.DataFrame({"a": pd.Categorical(["x"] * 50)})
    with pytest.raises(EngineError, match="fewer cells"):
        hm.compute_lisi(X, meta, ["a"], 30)            # 90 neighbours of 50 cells
    with pytest.raises(EngineError, match="perplexity"):
        hm.compute_lisi(np.zeros((500, 4)), pd.DataFrame({"a": pd.Categorical(["x"] * 500)}), ["a"], 60)
