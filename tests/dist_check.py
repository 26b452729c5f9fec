"""Run under torchrun with N ranks (one per GPU): the z-slab path against the single-GPU path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank steps its slab of a distributed Simulation; rank 0 also steps the same problem on one GPU
(the parity-tested single-GPU path).  Fields are assembled on rank 0 with a gloo reduce and must
agree to reduction-order noise; V-cycle counts and Δt must be identical.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from waterlily_b200 import AutoBody, Simulation, _cabi, lib, sim_step  # noqa: E402


def tgv(N, **kw):
    kappa = 2 * np.pi / N

    def ulam(i, x):
        X, Y, Z = x[..., 0] * kappa, x[..., 1] * kappa, x[..., 2] * kappa
        return -np.sin(X) * np.cos(Y) * np.cos(Z) if i == 1 else (np.cos(X) * np.sin(Y) * np.cos(Z) if i == 2 else np.zeros((1, 1, 1)))
    return Simulation((N, N, N), (0, 0, 0), N, U=1, nu=1 / (kappa * 1600), ulam=ulam, perdir=(1, 2, 3), **kw)


def sphere(n, m, exitBC, **kw):
    radius, center = m / 8, m / 2 - 1
    sdf = lambda x, t: torch.sqrt(((x - center) ** 2).sum(-1)) - radius
    return Simulation((n, m, m), (1, 0, 0), 2 * radius, nu=2 * radius / 3700, body=AutoBody(sdf), exitBC=exitBC, **kw)


def zwall(N, **kw):
    # non-periodic in z only through the walls, periodic in x: physical z boundaries on the end ranks
    return Simulation((N, N // 2, N), (0.5, 0, 0.25), N, nu=0.01, perdir=(1,), **kw)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    ident = [_cabi.nccl_unique_id(lib.load()) if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    dd = (rank, world, ident[0])
    cases = [("tgv64", lambda **kw: tgv(64, **kw), 4), ("sphere_exit", lambda **kw: sphere(64, 32, True, **kw), 4),
             ("sphere", lambda **kw: sphere(64, 32, False, **kw), 3), ("zwall", lambda **kw: zwall(32, **kw), 3)]
    ok = True
    for name, make, steps in cases:
        sd = make(device=dev, dist=dd)
        for _ in range(steps):
            sim_step(sd, remeasure=False)
        got = {}
        for f in ("u", "p", "f"):
            t = torch.from_numpy(np.ascontiguousarray(getattr(sd.flow, f).numpy()))
            dist.reduce(t, dst=0)  # disjoint planes: the sum assembles the global field
            got[f] = t.numpy()
        info = sd._eng.dist_info()
        if rank == 0:
            s1 = make(device=dev)
            for _ in range(steps):
                sim_step(s1, remeasure=False)
            for f in ("u", "p", "f"):
                ref = getattr(s1.flow, f).numpy().astype(np.float64)
                a = got[f].astype(np.float64)
                if f == "f":  # ghost layers of f are scratch that BC! never fills: compare the interior
                    sl = tuple(slice(1, -1) for _ in range(3))
                    a, ref = a[sl], ref[sl]
                err = np.sqrt(((a - ref) ** 2).sum() / max((ref ** 2).sum(), 1e-300))
                good = err < 2e-6
                ok &= good
                print(f"[{name}] {f}: rel L2 vs 1 GPU = {err:.2e} {'ok' if good else 'FAIL'}")
            same_n = sd.pois.n == s1.pois.n
            same_dt = np.allclose(np.array(sd.flow.dt, dtype=np.float64), np.array(s1.flow.dt, dtype=np.float64), rtol=1e-6)
            ok &= same_n and same_dt
            print(f"[{name}] V-cycles {sd.pois.n} vs {s1.pois.n} {'ok' if same_n else 'FAIL'}; dt {'ok' if same_dt else 'FAIL'}; "
                  f"levels {sd._eng.nlevels()}, {info['nz_local']} planes/rank, {info['collectives']} NCCL calls")
            s1.close()
        dist.barrier()   # closing a slab handle is collective (peer mappings): all ranks arrive together
        sd.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
