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


def check_accelerating_circle(cls):
    """"Circle in accelerating flow", maintests.jl:304-316: `make_accel_circle(radius=32, H=16)` — a 1024² grid, u_BC = (t, 0),
    a circle of radius 32 in the centre.  After one step the pressure force is the added-mass force −π L² (inertial coefficient
    1 of a circle), the flow past the shoulder is ≈ 2U, and the solver needs at most two V-cycles per projection afterwards."""
    import numpy as np
    import torch
    from waterlily_b200 import AutoBody, pressure_force, sim_step
    radius, H = 32, 16

    def u_bc(i, t):
        return t if i == 1 else 0 * t
    sim = cls((2 * H * radius, 2 * H * radius), u_bc, radius, U=1,
              body=AutoBody(lambda x, t: torch.sqrt(((x - H * radius) ** 2).sum(-1)) - radius))
    sim_step(sim)
    force = np.asarray(pressure_force(sim)) / (np.pi * sim.L ** 2)
    assert np.allclose(force, [-1.0, 0.0], atol=0.04), force                                    # :310
    u = sim.flow.u.numpy()
    assert u.max() / u[1, 1, 0] > 1.91                                                          # :311  (≈ 2U)
    for _ in range(3):
        sim_step(sim)
    assert all(n <= 2 for n in sim.pois.n), sim.pois.n                                          # :313
    return sim
