"""CPU mirror of the engine's device-side block permutation (test infrastructure only).

``perm_mode="device"`` replaces the reference's host ``torch.randperm(N)`` per update_R (harmony.py:471) by a keyed
bijection of [0, N) evaluated on the GPU (k_assign_feistel, harmonypy_b200/csrc/hmy_round.cuh): cell id -> position,
a 4-round Feistel network on 2*hb bits with cycle walking; block = position // int(N * block_size) like
harmony.py:474-475, :483-484.  This module restates that bijection in NumPy so that a device-mode run can be replayed
through the oracle with the SAME block membership: ``device_perm(N, seed, r)`` is the array the reference would have
drawn (perm[position] = cell) in the r-th round since the seed was set.

There is no reference counterpart (the reference only has the host stream, which perm_mode="reference" reproduces
bit for bit); tests/test_device_perm_mirror.py checks the bijection properties on the CPU, the GPU test compares the
engine's device-mode run with the oracle fed from this mirror.
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint64(0xFFFFFFFF)


def _mix(x, key):
    """hmy_mix: 32-bit avalanche (all arithmetic modulo 2^32, carried in uint64 lanes)."""
    x = (x ^ key) & _U32
    x = (x * np.uint64(0x9E3779B1)) & _U32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x85EBCA77)) & _U32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE3D)) & _U32
    x ^= x >> np.uint64(16)
    return x


def engine_seed(seed_option):
    """What hmy_set_option("seed", v) stores; v = 0 is also the library default."""
    return (int(seed_option) * 0x9E3779B97F4A7C15 + 0x243F6A8885A308D3) & 0xFFFFFFFFFFFFFFFF


def device_positions(n_global, seed_option, round_index, ids=None):
    """Position of every cell id in the pseudo-random order of round ``round_index`` (k_assign_feistel)."""
    n_global = int(n_global)
    hb = 1
    while (1 << (2 * hb)) < n_global:
        hb += 1
    mask = np.uint64((1 << hb) - 1)
    seed = engine_seed(seed_option)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64(((seed >> 32) ^ ((int(round_index) * 0x632BE5AB) & 0xFFFFFFFF)) & 0xFFFFFFFF)
    x = np.arange(n_global, dtype=np.uint64) if ids is None else np.asarray(ids, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    hbu = np.uint64(hb)
    while todo.any():
        v = x[todo]
        l, r = v >> hbu, v & mask
        for i in range(4):
            key = (k0 + np.uint64((0x9E3779B9 * (i + 1)) & 0xFFFFFFFF) + k1) & _U32
            f = _mix(r & _U32, key) & mask
            l, r = r, l ^ f
        v = (l << hbu) | r
        x[todo] = v
        todo[todo] = v >= np.uint64(n_global)
    return x.astype(np.int64)


def device_perm(n_global, seed_option, round_index):
    """perm with perm[position] = cell: drop-in for the torch.randperm(N) of harmony.py:471."""
    pos = device_positions(n_global, seed_option, round_index)
    perm = np.empty(int(n_global), dtype=np.int64)
    perm[pos] = np.arange(int(n_global), dtype=np.int64)
    return perm
