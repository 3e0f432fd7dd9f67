"""The oracle of the device-side centroid initialisation (oracle/kmeans_init_oracle.py): is it a faithful
k-means++ / Lloyd?  Compared with sklearn -- the reference's initialiser, harmony.py:369-373 -- on the reference's own
pbmc PCs.  CPU only."""
import numpy as np

from conftest import load_case
from oracle.kmeans_init_oracle import kmeans_init, kmeanspp_seed, splitmix64, uniform


def _zcos(name="pbmc"):
    inp, _ = load_case(name)
    Z = inp["Z"].astype(np.float32)
    return Z / np.linalg.norm(Z, axis=1, keepdims=True), int(inp["K"])


def test_splitmix64_reference_values():
    # first outputs of the published splitmix64 generator started at 0 (state advanced by the golden-ratio constant)
    s = np.uint64(0)
    out = []
    for _ in range(3):
        out.append(int(splitmix64(s)))
        with np.errstate(over="ignore"):
            s = s + np.uint64(0x9E3779B97F4A7C15)
    assert out == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    u = uniform(7, 3, np.arange(100000))
    assert 0.0 < u.min() and u.max() <= 1.0 and abs(u.mean() - 0.5) < 5e-3


def test_seeding_is_deterministic_spread_out_and_seed_dependent():
    Zc, K = _zcos()
    a, b, c = kmeanspp_seed(Zc, K, 0), kmeanspp_seed(Zc, K, 0), kmeanspp_seed(Zc, K, 1)
    assert (a == b).all() and len(set(a.tolist())) == K
    assert len(set(a.tolist()) & set(c.tolist())) < K // 2
    # plain D^2 sampling (no greedy local trials as in sklearn): on these diffuse 30-d cells its potential is about
    # that of K uniformly drawn cells (sklearn's greedy variant is ~8 % lower); what matters is the Lloyd result below
    def potential(idx):
        d2 = ((Zc[:, None, :] - Zc[idx][None, :, :]) ** 2).sum(-1)
        return d2.min(axis=1).sum()
    rnd = np.random.default_rng(0).choice(len(Zc), K, replace=False)
    assert potential(a) < 1.05 * potential(rnd)


def test_inertia_on_par_with_sklearn():
    from sklearn.cluster import KMeans
    Zc, K = _zcos()
    ours = np.array([kmeans_init(Zc, K, s)[1]["inertia"] for s in range(3)])
    theirs = np.array([KMeans(n_clusters=K, init="k-means++", n_init=1, max_iter=25, random_state=s).fit(Zc).inertia_ for s in range(3)])
    print("inertia ours", ours, "sklearn", theirs)
    assert ours.mean() < 1.05 * theirs.mean()


def test_max_iter_zero_returns_the_seed_rows():
    Zc, K = _zcos()
    C, info = kmeans_init(Zc, K, 5, max_iter=0)
    np.testing.assert_allclose(C, Zc[info["seeds"]], atol=2e-7)      # kmeans_init re-normalises the rows
