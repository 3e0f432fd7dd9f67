#!/usr/bin/env python
"""Timeline of the FUSED round kernel on 2 GPUs (rank 0 prints)."""
import os, sys, socket, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def worker(rank, world, port):
    import torch, torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bench
    from harmonypy_b200.harmony import Harmony, Comm
    from harmonypy_b200.synthetic import make_synthetic_arrays
    w = bench.WORKLOADS["syn1m"]
    N = w["per_gpu"] * world
    lo, hi = N * rank // world, N * (rank + 1) // world
    Z, codes = make_synthetic_arrays(N, w["d"], w["levels"], seed=0, lo=lo, hi=hi)
    Pr_b = bench.global_level_probs(w, N, lo, hi, codes, dist)
    comm = Comm(None)
    Y0 = bench.init_centroids(w, N) if rank == 0 else np.zeros((w["K"], w["d"]), np.float32)
    Y0 = comm.broadcast_array(Y0, 0)
    prob = bench.make_problem(w, Z, codes, Pr_b, N, lo)
    ho = Harmony(prob, 0.2, 10, 20, 1e-5, 1e-4, 0.05, False, 0, rank, perm_mode="device", comm=comm, run=False)
    eng = ho._engine
    ho.init_cluster(0, Y0)
    for _ in range(3):
        ho.kmeans_round()
    eng.set_option("trace", 1)
    torch.cuda.synchronize(); dist.barrier()
    ho.kmeans_round()
    G = eng.counter("grid"); nblk = eng.counter("nblk")
    buf = np.zeros((G + 1, 192), dtype=np.uint64)
    eng._ck(eng.lib.hmy_get(eng.h, 9, buf.ctypes.data_as(C.c_void_p), buf.nbytes), "trace")
    if rank == 0:
        t = buf[:G].astype(np.int64); ser = buf[G].astype(np.int64).reshape(32, 4)
        t0 = t[:, 0].min()
        st = lambda x: f"min {x.min()/1e3:7.1f} mean {x.mean()/1e3:7.1f} max {x.max()/1e3:7.1f} us"
        print("fused", eng.counter("fused"), "kernel span %.1f us" % ((t[:, 5 + 3 * (nblk - 1)].max() - t0) / 1e3))
        print("phase0", st(t[:, 1] - t[:, 0]), " barrier0+Told exchange", st(t[:, 2] - t[:, 1]))
        pr = np.array([t[:, 4 + 3 * b] - t[:, 3 + 3 * b] for b in range(nblk)])
        bw = np.array([t[:, 5 + 3 * b] - t[:, 4 + 3 * b] for b in range(nblk)])
        print("process", st(pr), " barrier wait", st(bw), " min-over-CTAs mean %.1f" % (bw.min(axis=1).mean() / 1e3))
        ok = ser[:, 0] > 0
        for name, a, b in (("fence", 0, 1), ("body (exchange)", 1, 2), ("reset+fence", 2, 3)):
            d = ser[ok, b] - ser[ok, a]
            print(f"serial {name:16s}", st(d), np.round(d / 1e3, 1))
    dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
