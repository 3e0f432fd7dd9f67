"""Cells sharded over 2 / 4 / 8 GPUs (one process per GPU): the fused mode (in-kernel peer exchange; the tensor-memory
round kernel where its shape limits hold) and the staged NCCL mode must give the single-GPU / reference result.
Needs as many CUDA devices as ranks (skipped otherwise)."""
import os
import socket

import numpy as np
import pytest

from conftest import load_case, rel_max

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, name, fused, relaxed=0):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from harmonypy_b200.harmony import Harmony, Problem
        inp, gold = load_case(name)
        prob = Problem(Z=inp["Z"], codes=inp["codes"], levels=inp["levels"], level_names=[], Pr_b=inp["Pr_b"],
                       theta=inp["theta"], lamb=inp["lamb"], lambda_estimation=bool(inp["lambda_estimation"]),
                       sigma=inp["sigma"], K=int(inp["K"]))
        ho = Harmony(prob, float(inp["alpha"]), int(inp["max_iter_harmony"]), int(inp["max_iter_kmeans"]),
                     float(inp["epsilon_kmeans"]), float(inp["epsilon_harmony"]), float(inp["block_size"]), False,
                     int(inp["random_state"]), rank, comm=True, init_centroids=inp["Y0"],
                     engine_options={"fused": fused, "relaxed": relaxed})
        Zc = ho.Z_corr                      # gathered over ranks
        q.put(dict(rank=rank, rounds=list(ho.kmeans_rounds), Z=Zc, obj=list(ho.objective_harmony),
                   lo=ho._lo, hi=ho._hi, fused=ho._engine.counter("fused"), tc5=ho._engine.counter("tc5")))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,fused", [(2, 1), (2, 0), (4, 1), (8, 1)])
@pytest.mark.parametrize("name", ["synth", "pbmc", "ircolitis"])
def test_sharded_runs_match_reference(name, world, fused):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if name == "ircolitis" and not fused:
        pytest.skip("the staged mode is covered by the two small cases")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, name, fused)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=900) for _ in procs], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    inp, gold = load_case(name)
    a = outs[0]
    for b in outs[1:]:
        assert b["fused"] == a["fused"] == fused and b["tc5"] == a["tc5"]
        assert b["rounds"] == a["rounds"]
        np.testing.assert_array_equal(a["Z"], b["Z"])          # every rank holds the same gathered result
    assert a["tc5"] == fused                                    # fused sharded runs are on the tensor-memory kernel
    assert a["rounds"] == list(gold["kmeans_rounds"])
    err = rel_max(a["Z"][gold["final_cells"]], gold["Zcorr_final"])
    print(f"\n[{name}] {world} GPUs, {'fused' if fused else 'staged'} mode: final Z_corr vs reference fp32 {err:.3e}")
    assert err < 1e-4
    np.testing.assert_allclose(a["obj"], gold["objective_harmony"], rtol=5e-5)


def test_two_gpus_relaxed_exchange_deviation():
    """Fused mode with one exchange per round ('relaxed', what north_star sketches): not exact --
    remote GPUs' block updates are one round stale -- so the deviation from the reference is
    measured and bounded here instead of being gated at 1e-4."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, "pbmc", 1, 1)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=600) for _ in procs], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    inp, gold = load_case("pbmc")
    err = rel_max(outs[0]["Z"][gold["final_cells"]], gold["Zcorr_final"])
    print(f"\n[pbmc] 2-GPU fused RELAXED mode: final Z_corr vs reference fp32 {err:.3e}, rounds {outs[0]['rounds']}")
    np.testing.assert_array_equal(outs[0]["Z"], outs[1]["Z"])
    assert err < 5e-2
