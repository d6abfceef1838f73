"""CPU: the work split of the refinement's fused sweep (csrc/refine.cu sweep_range, the function the kernel calls, through
mc3d_refine_sweep_range): block-owned item ranges that tile the shard, start at multiples of four items (pairs of elements stay
8-byte aligned), hold two 2-frame edges plus interior, and give the blocks next to a neighbour rank one trip less."""
import ctypes

import numpy as np
import pytest


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    import mc3d_b200
    return mc3d_b200.lib()


def ranges(lib, n_items, grid, joints, rank, world, elem):
    out = []
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    for b in range(grid):
        assert lib.mc3d_refine_sweep_range(n_items, grid, joints, rank, world, elem, b, ctypes.byref(lo), ctypes.byref(hi)) == 0
        out.append((lo.value, hi.value))
    return out


@pytest.mark.parametrize('elem', [4, 8])
@pytest.mark.parametrize('joints', [1, 17, 133])
def test_ranges_tile_the_shard(lib, elem, joints):
    rng = np.random.default_rng(joints * elem)
    trip = 256 * (8 // elem)
    for _ in range(60):
        grid = int(rng.integers(1, 445))
        per = int(rng.integers(4 * joints + 40, 6000))
        n_items = grid * per + int(rng.integers(0, grid))
        world = int(rng.integers(1, 9))
        rank = int(rng.integers(0, world))
        r = ranges(lib, n_items, grid, joints, rank, world, elem)
        assert r[0][0] == 0 and r[-1][1] == n_items
        for (a0, a1), (b0, _) in zip(r[:-1], r[1:]):
            assert a1 == b0                                        # contiguous
        for lo, hi in r:
            assert lo % 4 == 0                                     # aligned starts
            assert hi - lo >= 4 * joints + 28                      # two edges of 2 J items and some interior
        mids = [hi - lo for lo, hi in r[(1 if rank > 0 else 0):(len(r) - 1 if rank < world - 1 else len(r))]]
        short = max(per - trip, min(per, 480)) & ~3
        if grid >= 4 and world > 1 and short >= 4 * joints + 32:   # (never below two edges and some interior)
            if rank > 0:                                           # one trip less (or one trip at most) next to a neighbour rank
                assert r[0][1] - r[0][0] == short and short <= min(mids)
            if rank < world - 1:
                assert short <= r[-1][1] - r[-1][0] <= short + 3
        if len(mids) > 1:
            assert max(mids[:-1]) - min(mids[:-1]) <= 8            # even middle ranges


def test_single_rank_is_the_even_split(lib):
    n, G = 1_700_000, 296
    r = ranges(lib, n, G, 17, 0, 1, 4)
    assert [lo for lo, _ in r] == [(n * b // G) & ~3 for b in range(G)]


def test_bad_arguments(lib):
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    assert lib.mc3d_refine_sweep_range(100, 4, 17, 0, 1, 4, 4, ctypes.byref(lo), ctypes.byref(hi)) != 0       # block >= grid
    assert lib.mc3d_refine_sweep_range(100, 4, 17, 2, 2, 4, 0, ctypes.byref(lo), ctypes.byref(hi)) != 0       # rank >= world
    assert lib.mc3d_refine_sweep_range(100, 4, 17, 0, 1, 2, 0, ctypes.byref(lo), ctypes.byref(hi)) != 0       # element size
