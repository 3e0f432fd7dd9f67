"""Host protocol of engines that run one round ahead (C-ABI counter "lookahead" = 1: the tensor-memory round
kernel), on CPU.  The CUDA engine is replaced by a test engine that enforces the contract of include/harmony_b200.h:

  * the permutation of round r is handed over one call early (hmy_queue_perm before init, then every
    hmy_kmeans_round carries the NEXT round's),
  * a stage stores R only while option write_r = 1; hmy_ridge_correct and hmy_get(R) fail otherwise,
  * hmy_kmeans_round(ctx, perm, NULL) enqueues a round without returning its objective; hmy_objectives(n)
    returns the last n stages' sums, oldest first.

What is under test is harmonypy_b200.harmony (cluster(), init_cluster(), the write_r toggling and the order in
which permutations and objectives travel): the run must reproduce, bit for bit, the same engine driven through the
plain protocol (permutation with its own round, objective returned by every call)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from shard_engine import ShardOracleEngine          # noqa: E402
from harmonypy_b200 import _cabi                    # noqa: E402
from harmonypy_b200.harmony import run_harmony      # noqa: E402
from harmonypy_b200.synthetic import make_synthetic  # noqa: E402


class LookaheadEngine(ShardOracleEngine):
    lookahead = True

    def __init__(self, *a):
        super().__init__(*a)
        self.queue, self.objs = [], []
        self.write_r, self.r_valid, self.have_init = 1, True, False
        self.log = []                    # (stage, write_r, waited)

    def set_option(self, k, v):
        if k == "write_r":
            self.write_r = int(v)

    def queue_perm(self, perm=None):
        assert not self.have_init and not self.queue, "hmy_queue_perm: the next round already has its permutation"
        self.queue.append(np.array(perm, copy=True))

    def init_from_centroids(self, Y0):
        assert len(self.queue) == 1, "reference permutation mode: the first round's permutation comes before init"
        obj = super().init_from_centroids(Y0)
        self.have_init, self.r_valid = True, bool(self.write_r)
        self.objs.append(obj)
        self.log.append(("init", self.write_r, True))
        return obj

    def kmeans_round(self, perm, wait=True):
        assert self.have_init and len(self.queue) == 1, "permutation queue out of step"
        cur = self.queue.pop(0)
        self.queue.append(np.array(perm, copy=True))       # the NEXT round's
        obj = super().kmeans_round(cur)
        self.r_valid = bool(self.write_r)
        self.objs.append(obj)
        self.log.append(("round", self.write_r, bool(wait)))
        return obj if wait else None

    def objectives(self, n):
        assert 1 <= n <= 16 and n <= len(self.objs)
        return self.objs[-n:]

    def ridge_correct(self):
        assert self.r_valid, "hmy_ridge_correct: the last stage did not store R (option write_r = 0)"
        super().ridge_correct()

    def get(self, which):
        if which == _cabi.R:
            assert self.r_valid, "hmy_get(R): the last stage did not store R"
        return super().get(which)


def _run(factory, case):
    Z, meta = make_synthetic(case["N"], case["d"], case["levels"], seed=4)
    Y0 = Z[np.random.default_rng(2).choice(case["N"], case["K"], replace=False)]
    return run_harmony(Z, meta, list(meta.columns), nclust=case["K"], max_iter_harmony=case["iters"],
                       max_iter_kmeans=case["mk"], verbose=False, random_state=9, init_centroids=Y0,
                       engine_factory=factory, **case.get("kw", {}))


@pytest.mark.parametrize("case", [
    dict(N=1500, d=6, levels=[3], K=8, iters=3, mk=8),
    dict(N=1200, d=5, levels=[4, 2], K=6, iters=2, mk=3),                 # fewer rounds than the window
    dict(N=1000, d=5, levels=[2], K=5, iters=2, mk=1),
    dict(N=1300, d=6, levels=[3], K=7, iters=2, mk=20, kw=dict(epsilon_cluster=1e-2)),   # stops on the rule
])
def test_lookahead_protocol_reproduces_plain_protocol(case):
    a = _run(LookaheadEngine, case)
    b = _run(ShardOracleEngine, case)
    assert list(a.kmeans_rounds) == list(b.kmeans_rounds)
    for name in ("objective_kmeans", "objective_kmeans_dist", "objective_kmeans_entropy", "objective_kmeans_cross",
                 "objective_harmony"):
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=name)
    np.testing.assert_array_equal(a.Z_corr, b.Z_corr)
    np.testing.assert_array_equal(a.R, b.R)
    log = a._engine.log
    w = a.window_size
    # rounds after which cluster() cannot stop (harmony.py:455) travel without a host round trip and without R
    i = 0
    for n_rounds in a.kmeans_rounds:
        while log[i][0] == "init":
            i += 1
        for r in range(n_rounds):
            stage, write_r, waited = log[i + r]
            can_stop = r > w or r == case["mk"] - 1
            assert stage == "round" and write_r == int(can_stop), (r, log[i + r])
            if not can_stop:
                assert not waited
        i += n_rounds
    assert i == len(log)


def test_stages_called_one_by_one_always_store_R():
    from harmonypy_b200.harmony import Harmony, prepare_problem
    import pandas as pd
    Z, meta = make_synthetic(900, 5, [3], seed=1)
    prob, _ = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), nclust=6)
    Y0 = Z[:6]
    ho = Harmony(prob, 0.2, 10, 20, 1e-5, 1e-4, 0.05, False, 3, 0, engine_factory=LookaheadEngine, run=False)
    ho.init_cluster(3, Y0)
    assert ho.R.shape == (6, 900) or ho.R.shape == (900, 6)
    ho.kmeans_round()
    ho.update_R()
    _ = ho.R                              # valid after every stage called directly
    ho.moe_correct_ridge()
    assert [x[1] for x in ho._engine.log] == [1, 1, 1]
