"""CPU tests of the multi-rank host logic over gloo (world_size 2 and 4): the axis-sweep
FillBoundary / SumBoundary schedule of parallel.HaloExchanger against the oracle's brute-force
multi-box semantics, and the neighbour bookkeeping of the brick decomposition."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, n_cell, ng, src_ng, seed, out_dir):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from helpers import NumpyHaloOps
    from oracle import oracle
    from warpx_b200 import abi, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = parallel.brick_grid(world)
    geom = abi.make_geom(n_cell, (0, 0, 0), (1, 1, 1))
    ok = True
    for comp in (0, 4, 8):          # Ex (0,1,1), By (0,1,0), jz (1,1,0)
        stag = abi.YEE_STAG[comp]
        rng = np.random.default_rng(seed + comp)
        # the same global periodic field on every rank (unique values per location)
        glob = rng.standard_normal(n_cell[::-1])
        boxes = []
        for r in range(world):
            dec = parallel.Decomposition(tuple(n_cell), nb, r)
            hf = oracle.HostFab(dec.box_lo, dec.box_hi, ng, stag)
            d = hf.desc
            kk, jj, ii = np.meshgrid(*[np.arange(d.lo[ax], d.hi[ax] + 1) for ax in (2, 1, 0)], indexing="ij")
            hf.a[...] = glob[kk % n_cell[2], jj % n_cell[1], ii % n_cell[0]]
            # spoil the guards so that the exchange has to restore them
            valid = np.zeros(hf.a.shape, dtype=bool)
            valid[d.valid_slices()] = True
            hf.a[~valid] = rng.standard_normal(int((~valid).sum())) if r == rank else 7.0
            boxes.append(hf)
        mine = boxes[rank]
        dec = parallel.Decomposition(tuple(n_cell), nb, rank)
        halo = parallel.HaloExchanger(dec, NumpyHaloOps(torch, n_cell), dist)
        # ---- FillBoundary(ng') ----
        fab = mine.desc
        fab.host = mine.a
        halo.fill_boundary([fab], (ng[0] - 1, ng[1], ng[2] - 1))
        d = mine.desc
        kk, jj, ii = np.meshgrid(*[np.arange(d.lo[ax], d.hi[ax] + 1) for ax in (2, 1, 0)], indexing="ij")
        expect = glob[kk % n_cell[2], jj % n_cell[1], ii % n_cell[0]]
        filled = np.zeros(mine.a.shape, dtype=bool)
        filled[tuple(slice(ng[ax] - g, mine.a.shape[2 - ax] - (ng[ax] - g))
                     for ax, g in ((2, ng[2] - 1), (1, ng[1]), (0, ng[0] - 1)))] = True
        ok &= bool(np.array_equal(mine.a[filled], expect[filled]))
        # ---- SumBoundary(src_ng, dst = all) against the oracle on the full set of boxes ----
        rng2 = np.random.default_rng(1000 + seed + comp)
        allb = []
        for r in range(world):
            decr = parallel.Decomposition(tuple(n_cell), nb, r)
            hf = oracle.HostFab(decr.box_lo, decr.box_hi, ng, stag)
            hf.a[...] = rng2.standard_normal(hf.a.shape)
            allb.append(hf)
        mine2 = oracle.HostFab(dec.box_lo, dec.box_hi, ng, stag, data=allb[rank].a.copy())
        fab2 = mine2.desc
        fab2.host = mine2.a
        halo.sum_boundary([fab2], src_ng, ng)
        oracle.lib().orc_sum_boundary(oracle.fab_array(allb), world, abi.int3(src_ng), abi.int3(ng), C.byref(geom))
        err = np.max(np.abs(mine2.a - allb[rank].a)) / np.max(np.abs(allb[rank].a))
        ok &= bool(err < 1e-14)
    open(os.path.join(out_dir, "rank%d.%s" % (rank, "ok" if ok else "fail")), "w").close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_schedule_over_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    n_cell = (16, 12, 8)
    port = 29600 + world
    mp.spawn(_worker, args=(world, port, n_cell, (3, 3, 3), (2, 3, 1), 5, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank%d.ok" % r for r in range(world)]


def test_brick_decomposition_neighbours():
    from warpx_b200 import parallel
    assert parallel.brick_grid(1) == (1, 1, 1) and parallel.brick_grid(2) == (2, 1, 1)
    assert parallel.brick_grid(4) == (2, 2, 1) and parallel.brick_grid(8) == (2, 2, 2)
    dec = parallel.Decomposition((64, 64, 64), (2, 2, 2), 5)      # coord (1, 0, 1)
    assert dec.coord == (1, 0, 1) and dec.box_lo == (32, 0, 32) and dec.box_hi == (63, 31, 63)
    assert dec.neighbour(0, 1) == dec.neighbour(0, 0) == dec.rank_of((0, 0, 1))   # two bricks: same rank
    assert dec.neighbour(1, 1) == dec.rank_of((1, 1, 1))
    import numpy as np
    cell = np.array([31, 32, 63, 0, 40])
    down, up = parallel.particle_destinations(cell, dec, 0)
    assert list(up) == [True, False, False, True, False] and not down.any()
    dec4 = parallel.Decomposition((64, 8, 8), (4, 1, 1), 0)
    down, up = parallel.particle_destinations(np.array([0, 15, 16, 63, 48]), dec4, 0)
    assert list(up) == [False, False, True, False, False] and list(down) == [False, False, False, True, True]
