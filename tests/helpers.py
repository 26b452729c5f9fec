"""Shared helpers for the oracle-pin and parity tests (ports of helpers in the reference's
test/maintests.jl; line numbers refer to that file)."""
import numpy as np

from waterlily_b200 import _cabi


def L2(a):
    """WaterLily.L₂(a): sum of squares over inside(a) (src/util.jl:68)."""
    a = np.asarray(a, dtype=np.float64)
    inside = tuple(slice(1, n - 1) for n in a.shape)
    return float((a[inside] ** 2).sum())


def poisson_setup(make_engine, kind, Ng):
    """Poisson_setup(poisson, N) — maintests.jl:68-79.  ``Ng`` is the padded size, as there."""
    D = len(Ng)
    e = make_engine(tuple(n - 2 for n in Ng), pois_kind=kind)
    c = np.ones(tuple(Ng) + (D,), dtype=np.float32, order="F")
    e.upload("mu0", c)
    e.bc("mu0", (0.0,) * D)
    e.pois_update()
    soln = np.zeros(Ng, dtype=np.float32, order="F")
    soln += (np.arange(1, Ng[0] + 1, dtype=np.float32) - 2).reshape((-1,) + (1,) * (D - 1))
    e.upload("p", soln)
    e.pois_mult()                      # z = mult!(pois, soln)
    e.upload("p", np.zeros(Ng, dtype=np.float32, order="F"))
    n = e.pois_solve()
    x = e.download("p")
    x = x - x[(1,) * D]
    return L2(x - soln) / L2(soln), n, e


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.sqrt((b ** 2).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / (den if den > 0 else 1.0))
