"""Host-side logic of the row-band multi-GPU path, on CPU: band partition, halo plan, and a
real world_size-2 exchange over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_harness as ph

bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])


def test_band_rows_partition_the_screen():
    for h, world in [(1080, 1), (1080, 2), (1080, 8), (4320, 8), (1081, 4), (7, 3)]:
        rows = [bands.band_rows(h, world, r) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == h
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in rows]
        assert max(sizes) - min(sizes) <= 1


def test_halo_covers_the_neighbour_reach():
    # floor(cos*r)/round(cos*r) with r <= spatialRadius = 30 reaches at most 30 rows; +1 slack
    assert bands.halo_rows_for(30.0) == 31
    assert bands.halo_rows_for(30.5) == 32


def test_halo_plan_is_symmetric():
    h, world, halo = 4320, 8, 31
    plans = [bands.halo_plan(h, world, r, halo) for r in range(world)]
    assert len(plans[0]) == 1 and len(plans[-1]) == 1 and all(len(p) == 2 for p in plans[1:-1])
    for r, plan in enumerate(plans):
        b, e = bands.band_rows(h, world, r)
        for peer, send, recv in plan:
            assert b <= send[0] < send[1] <= e                      # sends only rows it owns
            assert recv[1] <= b or recv[0] >= e                     # receives only rows outside its band
            back = [x for x in plans[peer] if x[0] == r][0]
            assert back[1] == recv and back[2] == send              # the peer's message mirrors it
    with pytest.raises(ValueError):
        bands.halo_plan(100, 8, 3, 31)                              # bands thinner than the halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, halo, out_dir, bounds=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = bands.band_rows(h, world, rank, bounds)
    a0, a1 = max(0, b - halo), min(h, e + halo)
    # every row carries its global row index and the owner's rank; halos start out as garbage
    buf = torch.full((a1 - a0, w), -1.0)
    for y in range(b, e):
        buf[y - a0] = y * 1000.0 + rank
    plan = bands.halo_plan(h, world, rank, halo, bounds)
    bands.exchange_halo(buf, a0, plan, dist)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), buf.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bounds", [(2, None), (3, None), (3, [0, 17, 83, 120]), (2, [0, 61, 80])])
def test_exchange_halo_over_gloo(world, bounds, tmp_path):
    """Equal bands and bands of unequal height (bands.balanced_bounds cuts them like that): after one exchange every
    allocated row holds what its owner wrote."""
    h, w, halo = 40 * world, 8, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, h, w, halo, str(tmp_path), bounds), nprocs=world, join=True)
    for rank in range(world):
        b, e = bands.band_rows(h, world, rank, bounds)
        a0, a1 = max(0, b - halo), min(h, e + halo)
        got = np.load(tmp_path / f"rank{rank}.npy")
        for y in range(a0, a1):
            owner = [r for r in range(world) if bands.band_rows(h, world, r, bounds)[0] <= y < bands.band_rows(h, world, r, bounds)[1]][0]
            assert (got[y - a0] == y * 1000.0 + owner).all(), (rank, y)


def test_temporal_row_reach_bounds_the_reprojection():
    """bands.temporal_row_reach: the halo a moving camera needs (restirOmni.glsl:163-171 on the host, per band)."""
    import torch

    bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    w, h = 8, 40
    alloc_begin, row_begin, row_end = 5, 10, 30
    rows = 30
    # clip = (x, y, 0, 1): column-major prev_pv with px = x, py = y, pw = 1
    pv = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]
    world = torch.zeros((rows, w, 4), dtype=torch.float32)
    normal = torch.zeros((rows, w, 4), dtype=torch.int16)
    for r in range(rows):
        g = r + alloc_begin                       # global row; send it to row g + 3.25
        world[r, :, 1] = (g + 3.25) / h * 2.0 - 1.0
        world[r, :, 0] = 0.0
    normal[..., 2] = 32767
    assert bands.temporal_row_reach(world, normal, pv, w, h, alloc_begin, row_begin, row_end, torch) == 3 + 1
    # a far jump on a background pixel (normal == 0) does not count; on a surface pixel it does
    world[12, 3, 1] = (2 + 0.5) / h * 2.0 - 1.0   # global row 17 -> row 2
    normal[12, 3, :] = 0
    assert bands.temporal_row_reach(world, normal, pv, w, h, alloc_begin, row_begin, row_end, torch) == 4
    normal[12, 3, 2] = 32767
    assert bands.temporal_row_reach(world, normal, pv, w, h, alloc_begin, row_begin, row_end, torch) == 15 + 1
    # pixels reprojected off-screen are not looked up
    world[12, 3, 1] = 5.0
    assert bands.temporal_row_reach(world, normal, pv, w, h, alloc_begin, row_begin, row_end, torch) == 4


def test_balanced_bounds_equalise_cost():
    """bands.balanced_bounds: boundaries of equal cost from per-band times, minimum band height respected."""
    bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    h, world = 4320, 8
    equal = [bands.band_rows(h, world, r)[0] for r in range(world)] + [h]
    # uniform cost: nothing moves (up to a row)
    same = bands.balanced_bounds(h, world, equal, [1.0] * world, 137)
    assert all(abs(a - b) <= 1 for a, b in zip(same, equal))
    # the top band is three times cheaper per row, the bottom one twice as expensive: the cuts move accordingly
    secs = [1.0 / 3.0, 1, 1, 1, 1, 1, 1, 2.0]
    new = bands.balanced_bounds(h, world, equal, secs, 137)
    assert new[0] == 0 and new[-1] == h and all(b - a >= 137 for a, b in zip(new, new[1:]))
    dens = []
    for r in range(world):
        dens += [secs[r] / (equal[r + 1] - equal[r])] * (equal[r + 1] - equal[r])
    cost = [sum(dens[new[r]:new[r + 1]]) for r in range(world)]
    assert max(cost) / (sum(cost) / world) < 1.01
    assert new[1] > equal[1] and new[-2] > equal[-2]
    # band_rows / halo_plan follow explicit bounds
    assert bands.band_rows(h, world, 3, new) == (new[3], new[4])
    plan = bands.halo_plan(h, world, 3, 137, new)
    assert plan[0][1] == (new[3], new[3] + 137) and plan[1][2] == (new[4], new[4] + 137)
    # extreme skew: the minimum height wins
    skew = bands.balanced_bounds(h, world, equal, [100.0] + [0.001] * 7, 137)
    assert all(b - a >= 137 for a, b in zip(skew, skew[1:]))


def test_cpp_balanced_bounds_equal_the_python_ones():
    """restir_band_balanced_bounds (host C++, for C++ drivers of restir::BandSet) against bands.balanced_bounds."""
    import parity_harness as ph

    bands = __import__("restir_vulkan_b200.bands", fromlist=["bands"])
    rng = np.random.default_rng(4)
    for h, world, min_rows in [(4320, 8, 137), (1080, 2, 31), (2160, 3, 54), (8640, 8, 139), (64, 4, 16)]:
        equal = [bands.band_rows(h, world, r)[0] for r in range(world)] + [h]
        for _ in range(20):
            secs = list(rng.uniform(0.2, 3.0, world))
            want = bands.balanced_bounds(h, world, equal, secs, min_rows)
            assert ph.capi.band_balanced_bounds(h, equal, secs, min_rows) == want
            again = bands.balanced_bounds(h, world, want, secs[::-1], min_rows)      # from unequal bands too
            assert ph.capi.band_balanced_bounds(h, want, secs[::-1], min_rows) == again
    with pytest.raises(ph.capi.RestirError):
        ph.capi.band_balanced_bounds(100, [0, 50, 100], [1.0, 1.0], 60)              # two bands of 60 rows do not fit
