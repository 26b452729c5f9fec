"""Measures host-side launch cost on the GPU box: many tiny launches through the C ABI."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterlily_b200 import _cabi, lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
e = _cabi.Engine(lib.load(), (N, N, N), nu=0.01, perdir=(1, 2, 3))
L = e.nlevels() - 1
e.pois_jacobi(L, 10)
for it in (100, 1000):
    l0 = e.launch_count(); t0 = time.perf_counter()
    e.pois_jacobi(L, it)
    dt = time.perf_counter() - t0; n = e.launch_count() - l0
    print(f"coarsest-level Jacobi x{it}: {n} launches in {dt*1e3:.2f} ms -> {dt/n*1e6:.2f} us/launch")
rng = np.random.default_rng(0)
e.upload("u", rng.random(e.shape(_cabi.U), dtype=np.float32) - 0.5)
U = (0.0, 0.0, 0.0)
for prof in (False, True):
    e.profile(prof)
    e.mom_step(U)
    l0 = e.launch_count(); t0 = time.perf_counter()
    for _ in range(5): e.mom_step(U)
    dt = time.perf_counter() - t0; n = e.launch_count() - l0
    print(f"mom_step N={N} profiling={prof}: {dt/5*1e3:.2f} ms/step, {n/5:.0f} launches/step -> {dt/n*1e6:.2f} us/launch")
