"""Micro-benchmarks of individual C-ABI entry points on one GPU (development aid; not bench.py).
Times come from the library's own CUDA-event profiling classes."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterlily_b200 import _cabi, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--what", default="conv_diff")
    args = ap.parse_args()
    N = args.size
    e = _cabi.Engine(lib.load(), (N, N, N), nu=0.01, perdir=(1, 2, 3))
    rng = np.random.default_rng(0)
    for f in ("u", "u0", "V"):
        a = rng.random(e.shape(_cabi.FIELD_IDS[f]), dtype=np.float32) - 0.5
        e.upload(f, a)
        del a
    names = e.profile_classes()
    if args.what == "all":
        # one pass through every fine-level kernel class (for ncu captures)
        e.upload("p", rng.random(e.shape(_cabi.P), dtype=np.float32) - 0.5)
        e.upload("sigma", rng.random(e.shape(_cabi.SIGMA), dtype=np.float32) - 0.5)
        e.profile(True)
        for _ in range(args.reps):
            e.conv_diff("u")
            e.bdim()
            e.pois_residual(0)
            e.pois_jacobi(0, 1)
            e.pois_pcg(0, 2)
            e.project(1.0)
            e.cfl()
        for i, name in enumerate(names):
            ms, n, b = e.profile_read(i)
            if n:
                print(f"{name:18s} {ms / n:8.3f} ms/launch  {b / n / (ms / n * 1e-3) / 1e9 if ms else 0:7.0f} GB/s  ({n} launches)")
    if args.what == "conv_diff":
        e.conv_diff("u")
        e.profile(True)
        for _ in range(args.reps):
            e.conv_diff("u")
        ms, n, b = e.profile_read(names.index("conv_diff+bdim1"))
        print(f"conv_diff N={N} tma={os.environ.get('WL_CD_TMA','1')} cells/thread={os.environ.get('WL_CD_CELLS','2')} "
              f"zchunk={os.environ.get('WL_CD_TMA_ZCHUNK','64')}: "
              f"{ms / n:.3f} ms/launch, {b / n / (ms / n * 1e-3) / 1e9:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
