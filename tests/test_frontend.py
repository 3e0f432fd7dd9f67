"""Host-side argument normalisation against what the reference's front end produced
(fixtures come from the real run_harmony, harmony.py:116-173)."""
import numpy as np
import pandas as pd
import pytest

from conftest import load_case
from harmonypy_b200.harmony import prepare_problem, Harmony, get_device
from harmonypy_b200.synthetic import make_synthetic


def test_synth_two_covariates_tau_and_dynamic_lambda():
    inp, _ = load_case("synth")
    Z, meta = make_synthetic(6000, 20, [5, 3], seed=7)
    prob, vars_use = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), theta=[2.0, 1.0], lamb=-1,
                                     tau=5, nclust=40, sigma=0.12)
    assert vars_use == ["var0", "var1"]
    for k in ("Pr_b", "theta", "sigma", "lamb"):
        np.testing.assert_array_equal(getattr(prob, k), inp[k])
    np.testing.assert_array_equal(prob.codes, inp["codes"])
    np.testing.assert_array_equal(prob.Z, inp["Z"])
    assert prob.lambda_estimation and prob.K == 40 and prob.B == 8


def _meta(n=90):
    rng = np.random.default_rng(0)
    return pd.DataFrame({"b": rng.choice(["x", "y", "z"], n), "c": rng.choice(["p", "q"], n)})


def test_orientation_and_defaults():
    meta = _meta()
    Z = np.random.default_rng(1).standard_normal((90, 7))
    p1, _ = prepare_problem(Z, meta, "b")
    p2, _ = prepare_problem(Z.T, meta, ["b"])                 # PCs x cells is transposed (harmony.py:116-118)
    np.testing.assert_array_equal(p1.Z, p2.Z)
    assert p1.Z.dtype == np.float32 and p1.Z.shape == (90, 7)
    assert p1.K == int(min(round(90 / 30.0), 100))            # harmony.py:123-124
    np.testing.assert_array_equal(p1.theta, np.full(3, 2, np.float32))
    np.testing.assert_array_equal(p1.lamb, np.array([0, 1, 1, 1], np.float32))
    assert abs(p1.Pr_b.sum() - 1) < 1e-6


def test_theta_lambda_expansion_rules():
    meta = _meta()
    Z = np.zeros((90, 4)) + 1.0
    p, _ = prepare_problem(Z, meta, ["b", "c"], theta=[1.0, 3.0], lamb=[0.5, 2.0], nclust=5)
    np.testing.assert_array_equal(p.theta, np.array([1, 1, 1, 3, 3], np.float32))
    np.testing.assert_array_equal(p.lamb, np.array([0, .5, .5, .5, 2, 2], np.float32))
    p, _ = prepare_problem(Z, meta, ["b", "c"], theta=[1, 2, 3, 4, 5], lamb=[1, 2, 3, 4, 5], nclust=5)
    np.testing.assert_array_equal(p.theta, np.arange(1, 6, dtype=np.float32))
    np.testing.assert_array_equal(p.lamb, np.arange(0, 6, dtype=np.float32))
    p, _ = prepare_problem(Z, meta, ["b", "c"], theta=0.5, lamb=3, nclust=5)
    assert p.theta.tolist() == [0.5] * 5 and p.lamb.tolist() == [0, 3, 3, 3, 3, 3]
    assert abs(p.Pr_b.sum() - 2) < 1e-6                       # sums to V (harmony.py:169-170)
    with pytest.raises(AssertionError):
        prepare_problem(Z, meta, ["b", "c"], theta=[1, 2, 3], nclust=5)
    with pytest.raises(AssertionError):
        prepare_problem(np.zeros((50, 4)), meta, "b")         # cell-count mismatch (harmony.py:120)


def test_level_order_is_get_dummies_order():
    meta = pd.DataFrame({"b": ["b10", "b2", "b1", "b10", "b2", "b1"]})
    p, _ = prepare_problem(np.ones((6, 3)), meta, "b", nclust=2)
    want = pd.get_dummies(meta[["b"]]).to_numpy().astype(np.float32)   # harmony.py:133
    got = np.zeros_like(want)
    got[np.arange(6), p.codes[0]] = 1
    np.testing.assert_array_equal(got, want)
    assert p.level_names == list(pd.get_dummies(meta[["b"]]).columns)


def test_numeric_covariate_raises_like_reference():
    meta = pd.DataFrame({"b": [0, 1, 0, 1]})
    with pytest.raises(KeyError):
        prepare_problem(np.ones((4, 2)), meta, "b", nclust=2)


def test_tau_rescales_theta():
    meta = _meta(300)
    p, _ = prepare_problem(np.ones((300, 3)), meta, "b", nclust=10, tau=20)
    N_b = meta["b"].value_counts().sort_index().to_numpy()
    want = (np.float32(2) * (1 - np.exp(-(N_b / (10 * 20)) ** 2))).astype(np.float32)
    np.testing.assert_allclose(p.theta, want, rtol=1e-6)


def test_get_device_cuda_only():
    assert get_device(None) == 0 and get_device("cuda:3") == 3 and get_device("cuda") == 0
    with pytest.raises(ValueError):
        get_device("cpu")


def test_mode_arguments_are_validated_before_any_device_work():
    import pandas as pd
    from harmonypy_b200.harmony import Harmony, prepare_problem
    rng = np.random.default_rng(0)
    Z = rng.normal(size=(60, 5)).astype(np.float32)
    meta = pd.DataFrame({"b": pd.Categorical(rng.integers(0, 2, 60).astype(str))})
    prob, _ = prepare_problem(pd.DataFrame(Z), meta, ["b"], nclust=4)
    for kw in (dict(perm_mode="bogus"), dict(init_mode="bogus")):
        with pytest.raises(ValueError, match="must be"):
            Harmony(prob, 0.2, 1, 1, 1e-5, 1e-4, 0.05, False, 0, 0, run=False, **kw)
