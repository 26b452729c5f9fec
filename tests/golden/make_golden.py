"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/wl_oracle.cpp).

The reference (WaterLily.jl) is Julia and cannot run in this environment, so these are NOT reference
outputs: they are a frozen copy of what the oracle — pinned on the reference's own known-answer tests
(tests/test_oracle_pins.py) — produced when this file was last run.  They guard the oracle (and, on a
GPU box, the CUDA path) against accidental changes.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.oracle import OracleSimulation  # noqa: E402
from waterlily_b200 import AutoBody, sim_step  # noqa: E402


def circle(cls, n=96, m=64, Re=100):
    radius, center = m / 8, m / 2 - 1
    sdf = lambda x, t: torch.sqrt(((x - center) ** 2).sum(-1)) - radius
    return cls((n, m), (1, 0), 2 * radius, nu=2 * radius / Re, body=AutoBody(sdf))


def tgv(cls, N=32, Re=1600.0):
    kappa = 2 * np.pi / N

    def ulam(i, x):
        X, Y, Z = x[..., 0] * kappa, x[..., 1] * kappa, x[..., 2] * kappa
        return -np.sin(X) * np.cos(Y) * np.cos(Z) if i == 1 else (np.cos(X) * np.sin(Y) * np.cos(Z) if i == 2 else np.zeros_like(X))
    return cls((N, N, N), (0, 0, 0), N, U=1, nu=1 / (kappa * Re), ulam=ulam, perdir=(1, 2, 3))


CASES = {"circle_96x64_re100_10steps": (circle, 10, 4), "tgv_32_re1600_4steps": (tgv, 4, 4)}


def run(make, steps, stride, cls=OracleSimulation):
    sim = make(cls)
    for _ in range(steps):
        sim_step(sim, remeasure=False)
    sub = tuple(slice(None, None, stride) for _ in sim._eng.N)
    return {"u": sim.flow.u.numpy()[sub], "p": sim.flow.p.numpy()[sub], "dt": np.asarray(sim.flow.dt, dtype=np.float64),
            "n": np.asarray(sim.pois.n, dtype=np.int64), "stride": np.int64(stride)}


if __name__ == "__main__":
    for name, (make, steps, stride) in CASES.items():
        out = run(make, steps, stride)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()})
