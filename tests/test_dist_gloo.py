"""world_size-2 tests of the host's multi-GPU path on CPU (gloo).  The CUDA engine is replaced
by tests/shard_engine.py; what is under test is harmonypy_b200.harmony: contiguous sharding,
the shared permutation stream, table all-reduces through Comm, identical convergence decisions
on every rank, and the gathers behind the NumPy properties."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_max


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, case):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from shard_engine import ShardOracleEngine
        from harmonypy_b200.harmony import run_harmony
        from harmonypy_b200.synthetic import make_synthetic
        Z, meta = make_synthetic(case["N"], case["d"], case["levels"], seed=4)
        Y0 = Z[np.random.default_rng(2).choice(case["N"], case["K"], replace=False)]
        ho = run_harmony(Z, meta, list(meta.columns), nclust=case["K"], max_iter_harmony=case["iters"],
                         max_iter_kmeans=6, verbose=False, random_state=9, init_centroids=Y0,
                         comm=True, engine_factory=ShardOracleEngine, **case.get("kw", {}))
        out = dict(rank=rank, lo=ho._lo, hi=ho._hi, rounds=list(ho.kmeans_rounds),
                   obj=list(ho.objective_kmeans), Z=ho.Z_corr, R=ho.R, O=ho.O,
                   allreduces=ho._engine.counter("allreduce"))
        q.put(out)
    finally:
        dist.destroy_process_group()


def _run(case, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, case)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(outs, key=lambda o: o["rank"])


def _single(case):
    from oracle.harmony_oracle import HarmonyOracle, onehot_from_codes, torch_perm_source
    from harmonypy_b200.harmony import prepare_problem
    from harmonypy_b200.synthetic import make_synthetic
    Z, meta = make_synthetic(case["N"], case["d"], case["levels"], seed=4)
    Y0 = Z[np.random.default_rng(2).choice(case["N"], case["K"], replace=False)]
    prob, _ = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), nclust=case["K"], **case.get("kw", {}))
    orc = HarmonyOracle(prob.Z.T, onehot_from_codes(prob.codes, prob.levels, np.float64), prob.Pr_b, prob.sigma,
                        prob.theta, prob.lamb, lambda_estimation=prob.lambda_estimation, max_iter_kmeans=6,
                        dtype=np.float64)
    orc.init_from_centroids(Y0.T)
    orc.harmonize(case["iters"], torch_perm_source(case["N"], 9))
    return orc


@pytest.mark.parametrize("case", [
    dict(N=3001, d=8, levels=[3], K=12, iters=2),
    dict(N=2500, d=6, levels=[4, 2], K=9, iters=2, kw=dict(lamb=-1, theta=[1.5, 0.5])),
])
def test_two_ranks_reproduce_single_process_oracle(case):
    outs = _run(case)
    orc = _single(case)
    a, b = outs
    assert (a["lo"], a["hi"]) == (0, case["N"] // 2) and (b["lo"], b["hi"]) == (case["N"] // 2, case["N"])
    # every rank takes the same convergence decisions and sees the same gathered result
    assert a["rounds"] == b["rounds"] == list(orc.kmeans_rounds)
    np.testing.assert_array_equal(a["Z"], b["Z"])
    np.testing.assert_allclose(a["obj"], b["obj"], rtol=0, atol=0)
    np.testing.assert_allclose(a["obj"], orc.objective_kmeans, rtol=1e-9)
    assert rel_max(a["Z"], orc.Z_corr.T) < 1e-6          # fp32 property cast only
    assert rel_max(a["R"], orc.R.T) < 1e-6
    assert rel_max(a["O"], orc.O) < 1e-6
    # exact mode: 1 (removed sums) + n_blocks all-reduces of K x B per round
    assert a["allreduces"] == b["allreduces"] > 0


def _worker_presharded(rank, world, port, q, case):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        import dataclasses
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from shard_engine import ShardOracleEngine
        from harmonypy_b200.harmony import Harmony, prepare_problem
        from harmonypy_b200.synthetic import make_synthetic
        N = case["N"]
        Z, meta = make_synthetic(N, case["d"], case["levels"], seed=4)
        Y0 = Z[np.random.default_rng(2).choice(N, case["K"], replace=False)]
        full, _ = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), nclust=case["K"])
        lo, hi = case["edges"][rank], case["edges"][rank + 1]
        # what a caller that already holds one shard per rank passes: its own rows, the global size and its offset
        mine = dataclasses.replace(full, Z=full.Z[lo:hi].copy(), codes=full.codes[:, lo:hi].copy(), n_global=N, shard_lo=lo)
        ho = Harmony(mine, 0.2, case["iters"], 6, 1e-5, 1e-4, 0.05, False, 9, 0, init_centroids=Y0, comm=True,
                     engine_factory=ShardOracleEngine)
        q.put(dict(rank=rank, lo=ho._lo, hi=ho._hi, rounds=list(ho.kmeans_rounds), obj=list(ho.objective_kmeans),
                   Z=ho.Z_corr, R=ho.R, Phi=ho.Phi, local=ho.result_local()))
    finally:
        dist.destroy_process_group()


def test_uneven_presharded_input_lands_at_its_own_offsets():
    """Problem.shard_lo / n_global: ranks hold unequal row ranges that differ from the even split; gathers (Z_corr, R,
    Phi) must place every rank's rows at ITS offset, and the run must equal the single-process oracle."""
    case = dict(N=3001, d=8, levels=[3], K=12, iters=2, edges=[0, 1100, 3001])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_presharded, args=(r, 2, port, q, case)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in procs], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = outs
    assert (a["lo"], a["hi"], b["lo"], b["hi"]) == (0, 1100, 1100, 3001)
    orc = _single(case)
    assert a["rounds"] == b["rounds"] == list(orc.kmeans_rounds)
    np.testing.assert_array_equal(a["Z"], b["Z"])
    assert rel_max(a["Z"], orc.Z_corr.T) < 1e-6 and rel_max(a["R"], orc.R.T) < 1e-6
    np.testing.assert_array_equal(a["Z"][:1100], a["local"])
    np.testing.assert_array_equal(a["Z"][1100:], b["local"])
    from harmonypy_b200.synthetic import make_synthetic
    from harmonypy_b200.harmony import prepare_problem
    Z, meta = make_synthetic(case["N"], case["d"], case["levels"], seed=4)
    full, _ = prepare_problem(pd.DataFrame(Z), meta, list(meta.columns), nclust=case["K"])
    want = np.zeros((case["N"], 3), np.float32)
    want[np.arange(case["N"]), full.codes[0]] = 1
    np.testing.assert_array_equal(a["Phi"], want)
    np.testing.assert_array_equal(b["Phi"], want)


def test_comm_shard_covers_all_cells():
    from harmonypy_b200.harmony import Comm

    class Fake(Comm):
        def __init__(self, rank, world):
            self.rank, self.world = rank, world
    for N, W in ((10, 3), (1000003, 8), (7, 8)):
        edges = [Fake(r, W).shard(N) for r in range(W)]
        assert edges[0][0] == 0 and edges[-1][1] == N
        assert all(edges[i][1] == edges[i + 1][0] for i in range(W - 1))
