"""In-tree build of libharmony_b200.so (nvcc, sm_100a only).

    python -m harmonypy_b200.build [--force]

One object per (KPT, JPW) kernel instantiation plus the C-ABI translation unit, compiled
in parallel, linked into ``harmonypy_b200/libharmony_b200.so`` (git-ignored; it travels to
the GPU box with the working tree).  No torch, no CPU fallback: the Python host refuses to
run without this library.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libharmony_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
INSTANCES = [(k, j) for k in (1, 2, 4, 8) for j in (4, 8, 16)]
MMA_INSTANCES = [(4, 1), (8, 1), (14, 1), (16, 1), (8, 2), (14, 2), (16, 2)]
TC5_INSTANCES = [4, 7, 8]          # 16-cluster column chunks of the tensor-memory round kernel (option "tc5")


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _sources_digest():
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _compile(job):
    src, obj, defs = job
    cmd = [_nvcc(), *ARCH, *FLAGS, *defs, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(obj)}:\n{log[-4000:]}")
    return obj


def build(force=False, jobs=None, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    digest = _sources_digest()
    stamp = LIB + ".stamp"
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    work = [(os.path.join(CSRC, "hmy_api.cu"), os.path.join(OBJ, "hmy_api.o"), [])]
    for k, j in INSTANCES:
        work.append((os.path.join(CSRC, "hmy_inst.cu"), os.path.join(OBJ, f"hmy_inst_{k}_{j}.o"),
                     [f"-DHMY_KPT={k}", f"-DHMY_JPW={j}"]))
    for nt, wn in MMA_INSTANCES:
        work.append((os.path.join(CSRC, "hmy_inst_mma.cu"), os.path.join(OBJ, f"hmy_inst_mma_{nt}_{wn}.o"),
                     [f"-DHMY_NT={nt}", f"-DHMY_WN={wn}"]))
    work.append((os.path.join(CSRC, "hmy_lisi.cu"), os.path.join(OBJ, "hmy_lisi.o"), []))
    for nc in TC5_INSTANCES:
        work.append((os.path.join(CSRC, "hmy_inst_tc5.cu"), os.path.join(OBJ, f"hmy_inst_tc5_{nc}.o"),
                     [f"-DHMY_TC5_NC={nc}"]))
    jobs = jobs or min(len(work), os.cpu_count() or 4)
    if verbose:
        print(f"[harmonypy_b200.build] compiling {len(work)} objects for sm_100a with {jobs} jobs", flush=True)
    with ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(_compile, work))
    cmd = [_nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[harmonypy_b200.build] wrote {LIB}", flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
