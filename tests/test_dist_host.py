"""Host-side logic of the z-slab decomposition, on CPU: the slab plan (wl_dist_plan, no GPU touched)
and the halo-exchange protocol (same send/recv order as wl_dist.cu) run by two gloo ranks on numpy
planes.  The CUDA path itself is checked on ≥2 GPUs by tests/test_multi_gpu.py."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waterlily_b200 import _cabi, lib


def test_plan_covers_domain_and_agglomerates():
    L = lib.load()
    for R in (1, 2, 4, 8):
        plans = [_cabi.dist_plan(L, (512, 512, 512), (1, 2, 3), r, R) for r in range(R)]
        nl = len(plans[0])
        assert nl == 9                                       # 512³ → 2³
        for l in range(nl):
            nzg = 512 >> l
            d = plans[0][l]["distributed"]
            assert all(p[l]["distributed"] == d for p in plans)
            if d:
                assert R > 1 and nzg // R >= 16 and nzg ** 3 > 1 << 21   # enough planes per rank, and more than 128³ cells in all
                assert [p[l]["zoff"] for p in plans] == [r * (nzg // R) for r in range(R)]   # disjoint cover
                assert sum(p[l]["nz_local"] for p in plans) == nzg
                assert all(p[l]["zlo_halo"] and p[l]["zhi_halo"] for p in plans)               # periodic ring
            else:
                assert all(p[l]["nz_local"] == nzg and p[l]["zoff"] == 0 for p in plans)
        if R == 8:
            assert [p["distributed"] for p in plans[0]] == [1, 1, 0, 0, 0, 0, 0, 0, 0]       # 64, 32 planes/rank; the 128³ level and below are replicated
    # the plane criterion alone (WL_DIST_REPL_CELLS=0): 64, 32, 16 planes/rank are distributed, 8 and fewer replicated
    os.environ["WL_DIST_REPL_CELLS"] = "0"
    try:
        assert [p["distributed"] for p in _cabi.dist_plan(L, (512, 512, 512), (1, 2, 3), 0, 8)] == [1, 1, 1, 0, 0, 0, 0, 0, 0]
    finally:
        del os.environ["WL_DIST_REPL_CELLS"]
    # non-periodic z: end ranks keep a physical boundary
    plans = [_cabi.dist_plan(L, (256, 128, 128), (), r, 4) for r in range(4)]
    assert [p[0]["zlo_halo"] for p in plans] == [0, 1, 1, 1] and [p[0]["zhi_halo"] for p in plans] == [1, 1, 1, 0]


def test_plan_rejects_bad_splits():
    L = lib.load()
    with pytest.raises(_cabi.WLError, match="even number of z-planes per rank"):
        _cabi.dist_plan(L, (64, 64, 24), (), 0, 8)          # 3 planes per rank
    with pytest.raises(_cabi.WLError, match="needs D = 3"):
        _cabi.dist_plan(L, (64, 64), (), 0, 2)


def _halo_worker(rank, world, port, periodic, depth, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = lib.load()
    NZ, plane = 16, 6
    p = _cabi.dist_plan(L, (8, 8, NZ), (1, 2, 3) if periodic else (), rank, world)[0]
    nz, zoff = p["nz_local"], p["zoff"]
    n = nz + 2
    # local array with `depth-1` spare planes each side: index k ↔ row k + depth - 1
    a = np.full((n + 2 * (depth - 1), plane), np.nan, dtype=np.float32)
    row = lambda k: k + depth - 1
    for k in range(1, n - 1):
        a[row(k)] = 100 * (zoff + k) + np.arange(plane)           # global plane id in the value
    up, dn = (rank + 1) % world, (rank - 1) % world
    t = torch.from_numpy(a)
    # same order as dist_halo: send_up, recv_dn, send_dn, recv_up
    ops = []
    if p["zhi_halo"]:
        ops.append(dist.P2POp(dist.isend, t[row(n - 1 - depth):row(n - 1)].clone(), up))
    if p["zlo_halo"]:
        lo = torch.empty(depth, plane)
        ops.append(dist.P2POp(dist.irecv, lo, dn))
    if p["zlo_halo"]:
        ops.append(dist.P2POp(dist.isend, t[row(1):row(1 + depth)].clone(), dn))
    if p["zhi_halo"]:
        hi = torch.empty(depth, plane)
        ops.append(dist.P2POp(dist.irecv, hi, up))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    ok = True
    if p["zlo_halo"]:
        for d in range(depth):                                         # ghost plane 1-depth+d ← global plane below
            gk = (zoff + 1 - depth + d - 1) % NZ + 1
            ok &= bool(np.array_equal(lo[d].numpy(), 100 * gk + np.arange(plane, dtype=np.float32)))
    if p["zhi_halo"]:
        for d in range(depth):
            gk = (zoff + nz + d) % NZ + 1
            ok &= bool(np.array_equal(hi[d].numpy(), 100 * gk + np.arange(plane, dtype=np.float32)))
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("periodic,depth", [(True, 1), (True, 2), (False, 2)])
def test_halo_protocol_two_ranks_gloo(periodic, depth):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + depth + (10 if periodic else 0)
    mp.spawn(_halo_worker, args=(world, port, periodic, depth, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
