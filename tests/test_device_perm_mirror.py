"""CPU checks of the NumPy mirror of the device-side block permutation (oracle/device_perm.py): it is a bijection
for every size, it depends on seed and round, and the blocks it induces have the reference's sizes."""
import numpy as np
import pytest

from oracle.device_perm import device_perm, device_positions, engine_seed
from oracle.harmony_oracle import block_bounds


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 17, 255, 256, 257, 1000, 4097, 65536, 100003])
def test_bijection_for_every_size(n):
    for seed, rnd in ((0, 0), (0, 1), (7, 3)):
        pos = device_positions(n, seed, rnd)
        assert pos.min() == 0 and pos.max() == n - 1 and len(np.unique(pos)) == n
        perm = device_perm(n, seed, rnd)
        assert (pos[perm] == np.arange(n)).all()


def test_depends_on_seed_and_round_and_mixes_well():
    n = 50000
    a, b, c = device_positions(n, 0, 0), device_positions(n, 0, 1), device_positions(n, 1, 0)
    assert (a != b).mean() > 0.99 and (a != c).mean() > 0.99
    # neighbouring cells land in unrelated blocks: block ids of consecutive cells are almost uncorrelated
    blk = a // (n // 20)
    r = np.corrcoef(blk[:-1], blk[1:])[0, 1]
    assert abs(r) < 0.02
    assert engine_seed(0) == 0x243F6A8885A308D3


def test_blocks_have_the_reference_sizes():
    n, bs = 12345, 0.05
    bounds = block_bounds(n, bs)                      # harmony.py:474-475, :483-484
    pos = device_positions(n, 3, 2)
    cpb = int(n * bs)
    blk = np.minimum(pos // cpb, len(bounds) - 1)
    sizes = np.bincount(blk, minlength=len(bounds))
    assert list(sizes) == [hi - lo for lo, hi in bounds]
