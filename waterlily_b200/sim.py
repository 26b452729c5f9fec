"""Host-side mirror of the reference's user API for the ``mom_step!`` path.

Same names, argument meaning and error behaviour as the reference (marinlauber/WaterLily v1.3.0):

=====================  ====================================  =================================
here                   reference                             file:line
=====================  ====================================  =================================
``Simulation``         ``Simulation(dims,u_BC,L;...)``       src/WaterLily.jl:59-79
``sim_step``           ``sim_step!(sim[,t_end];...)``        src/WaterLily.jl:98-109
``sim_time``           ``sim_time(sim)``                     src/WaterLily.jl:89
``measure``            ``measure!(sim,t)``                   src/WaterLily.jl:116-119
``Flow``               ``Flow``                              src/Flow.jl:92-122
``mom_step``           ``mom_step!(a::Flow,b)``              src/Flow.jl:153-169
``MultiLevelPoisson``  ``MultiLevelPoisson`` / ``solver!``   src/MultiLevelPoisson.jl:44-99
=====================  ====================================  =================================

Julia is not available in this environment, so the host language above the C ABI is Python
(``julia/WaterLilyB200.jl`` shows the ``ccall`` binding a Julia maintainer would load instead).
All field storage and every per-step kernel live behind the handle in ``libwaterlily_b200.so``;
this module only sequences calls and evaluates the host-side closures (``u_BC(i,t)``, ``g(i,t)``,
``uλ``, the body ``sdf``) exactly where the reference evaluates them on the host.

``perdir`` uses the reference's 1-based direction numbers.
"""
from __future__ import annotations

import math

import numpy as np

from . import _cabi
from ._cabi import Engine
from .body import AbstractBody, LocLazy, NoBody, measure_fields


class DeviceField:
    """A field that lives behind the handle — the ``B200Array`` of INTEGRATION.md.

    ``np.asarray(field)`` / ``field.numpy()`` downloads (``Array(::B200Array)``), ``field.set(a)``
    uploads (``copyto!``)."""

    def __init__(self, eng: Engine, field: int):
        self._eng, self._field = eng, field

    @property
    def shape(self):
        return self._eng.shape(self._field)

    def numpy(self):
        return self._eng.download(self._field)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def set(self, arr):
        self._eng.upload(self._field, arr)

    def __getitem__(self, idx):
        return self.numpy()[idx]


class LevelField(DeviceField):
    def __init__(self, eng: Engine, level: int, lfield: int):
        self._eng, self._level, self._lfield = eng, level, lfield

    @property
    def shape(self):
        return self._eng.level_shape(self._level, self._lfield)

    def numpy(self):
        return self._eng.level_download(self._level, self._lfield)

    def set(self, arr):
        self._eng.level_upload(self._level, self._lfield, arr)


def BCTuple(U, dt, N, t=None):
    """Reference src/Flow.jl:79-80."""
    if callable(U):
        t = float(np.sum(np.asarray(dt, dtype=np.float64))) if t is None else t
        return tuple(U(i, t) for i in range(1, N + 1))
    return tuple(U)


def _ddt(fun, i, t):
    """d/dt of a host closure ``U(i,t)`` — the reference uses ForwardDiff.derivative (src/Flow.jl:71-72).

    In order: (1) a user-supplied exact derivative ``fun.ddt(i,t)``; (2) forward-mode AD through the
    closure when it is written with torch ops (``torch.func.jvp``, exact like ForwardDiff); (3) a
    4th-order finite difference in Float64 — central, or one-sided forward when the two disagree,
    i.e. at a kink such as a ramp that starts at ``t`` (ForwardDiff differentiates the branch taken
    at ``t``, which for ``t ≥ t_kink`` conditions is the forward one)."""
    exact = getattr(fun, "ddt", None)
    if callable(exact):
        return float(exact(i, t))
    try:
        import torch
        tt = torch.tensor(float(t), dtype=torch.float64)
        y, dy = torch.func.jvp(lambda s: torch.as_tensor(fun(i, s), dtype=torch.float64), (tt,), (torch.ones_like(tt),))
        if torch.is_tensor(dy) and bool(torch.isfinite(dy)):
            # a closure that ignores the dual part (pure-Python arithmetic on float(s)) raises above or returns 0 with y ≠ const;
            # accept the AD result only if it agrees with a coarse difference, else fall through
            h = 1e-3 * max(1.0, abs(t))
            fd = (float(fun(i, t + h)) - float(fun(i, t))) / h
            if abs(float(dy) - fd) <= 1e-2 * max(1.0, abs(fd)):
                return float(dy)
    except Exception:
        pass
    h = 1e-3 * max(1.0, abs(t))
    f = lambda s: float(fun(i, s))
    central = (f(t - 2 * h) - 8 * f(t - h) + 8 * f(t + h) - f(t + 2 * h)) / (12 * h)
    forward = (-25 * f(t) + 48 * f(t + h) - 36 * f(t + 2 * h) + 16 * f(t + 3 * h) - 3 * f(t + 4 * h)) / (12 * h)
    return forward if abs(central - forward) > 1e-6 * max(1.0, abs(forward)) else central


class Flow:
    """Reference src/Flow.jl:92-122.  Fields are :class:`DeviceField` views of the handle."""

    def __init__(self, eng: Engine, N, U, dt=0.25, nu=0.0, g=None, ulam=None, perdir=(), exitBC=False, T=np.float32):
        if T is not np.float32 and np.dtype(T) != np.float32:
            raise AssertionError("waterlily_b200 implements the reference's default T=Float32 only")
        self._eng = eng
        self.N, self.D = tuple(N), len(N)
        self.u, self.u0, self.f = DeviceField(eng, _cabi.U), DeviceField(eng, _cabi.U0), DeviceField(eng, _cabi.F)
        self.p, self.sigma = DeviceField(eng, _cabi.P), DeviceField(eng, _cabi.SIGMA)
        self.V, self.mu0, self.mu1 = DeviceField(eng, _cabi.V), DeviceField(eng, _cabi.MU0), DeviceField(eng, _cabi.MU1)
        self.U, self.nu, self.g = U, np.float32(nu), g
        self.dt = [np.float32(dt)]  # Δt history, host memory like the reference
        self.exitBC, self.perdir = bool(exitBC), tuple(perdir)
        # u = apply!(uλ,u); BC!; exitBC!(u,u,U,0); u⁰ = copy(u)   (src/Flow.jl:114-116)
        Ng = eng.Ng
        u = np.zeros(Ng + (self.D,), dtype=np.float32, order="F")
        if ulam is not None:
            for i in range(1, self.D + 1):
                u[..., i - 1] = np.asarray(ulam(i, LocLazy(i, Ng)), dtype=np.float32)
        U0 = BCTuple(U, [0.0], self.D)
        eng.upload(_cabi.U, u)
        eng.bc(_cabi.U, U0, exitBC, None)
        eng.upload(_cabi.U0, eng.download(_cabi.U))
        eng.exit_bc(U0, 0.0)
        eng.upload(_cabi.U0, eng.download(_cabi.U))

    def time(self):
        """Reference src/Flow.jl:129."""
        return float(np.sum(np.asarray(self.dt[:-1], dtype=np.float64)))


class MultiLevelPoisson:
    """Reference src/MultiLevelPoisson.jl:44-60: view of the solver state owned by the handle."""

    def __init__(self, eng: Engine):
        self._eng = eng
        self.n = []  # V-cycles per solve, two entries per mom_step! (:98)

    @property
    def levels(self):
        eng = self._eng
        return [type("PoissonView", (), {name: LevelField(eng, l, fid) for name, fid in _cabi.LEVEL_IDS.items()})()
                for l in range(eng.nlevels())]

    def update(self):
        """``update!(ml)`` (:62-68)."""
        self._eng.pois_update()

    def solver(self, tol=1e-4, itmx=32):
        """``solver!(ml;tol,itmx)`` (:87-99)."""
        n = self._eng.pois_solve(tol, itmx)
        self.n.append(n)
        return n


class _SimulationBase:
    """Everything of ``Simulation`` that is host logic; subclasses bind an ABI library."""

    _cabi_lib = None  # set by subclasses

    def __init__(self, dims, u_BC, L, *, dt=0.25, nu=0.0, g=None, U=None, eps=1, perdir=(), ulam=None, exitBC=False,
                 body: AbstractBody | None = None, T=np.float32, device=0, dist=None):
        N = len(dims)
        # the reference's constructor assertions (src/WaterLily.jl:70-72)
        if callable(u_BC) and callable(ulam):
            raise AssertionError("`u_BC` and `uλ` cannot be both specified as Function")
        if U is None and callable(u_BC):
            raise AssertionError("`U` must be specified if `u_BC` is a Function")
        if ulam is None:
            if callable(u_BC):
                ulam = lambda i, x: np.full(x.shape[:-1], u_BC(i, 0.0), dtype=np.float32)
            else:
                ulam = lambda i, x: np.full(x.shape[:-1], u_BC[i - 1], dtype=np.float32)
        if U is None:
            U = math.sqrt(sum(float(v) ** 2 for v in u_BC))
        self.U, self.L, self.eps = U, L, eps
        self.body = body if body is not None else NoBody()
        self._eng = Engine(self._lib(), dims, nu=nu, dt0=dt, perdir=perdir, exitBC=exitBC, device=device, dist=dist)
        self.flow = Flow(self._eng, dims, u_BC, dt=dt, nu=nu, g=g, ulam=ulam, perdir=perdir, exitBC=exitBC, T=T)
        self.pois = MultiLevelPoisson(self._eng)
        measure_flow(self.flow, self.body, t=0.0, eps=eps)
        self.pois.update()

    @classmethod
    def _lib(cls):
        raise NotImplementedError

    def close(self):
        self._eng.close()


def measure_flow(a: Flow, body: AbstractBody, t=0.0, eps=1):
    """``measure!(a::Flow,body;t,ϵ)`` — reference src/Body.jl:31-53 (NoBody: :76, a no-op)."""
    if type(body) is NoBody:
        return
    eng = a._eng
    mu0, mu1, V, d = measure_fields(body, eng.Ng, t=t, eps=eps)
    sg = eng.download(_cabi.SIGMA)
    inside = tuple(slice(1, n - 1) for n in eng.Ng)
    sg[inside] = d[inside]  # the reference uses a.σ as scratch for d (:34,50)
    eng.upload(_cabi.SIGMA, sg)
    eng.upload(_cabi.MU0, mu0)
    eng.upload(_cabi.MU1, mu1)
    eng.upload(_cabi.V, V)
    zero = (0.0,) * a.D
    eng.bc(_cabi.MU0, zero, False, None)
    eng.bc(_cabi.V, zero, a.exitBC, None)


def time(sim):
    return sim.flow.time()


def sim_time(sim):
    """Reference src/WaterLily.jl:89."""
    return sim.flow.time() * sim.U / sim.L


def measure(sim, t=None):
    """``measure!(sim,t=sum(Δt))`` — reference src/WaterLily.jl:116-119."""
    t = float(np.sum(np.asarray(sim.flow.dt, dtype=np.float64))) if t is None else t
    measure_flow(sim.flow, sim.body, t=t, eps=sim.eps)
    sim.pois.update()


def mom_step(a: Flow, b: MultiLevelPoisson):
    """``mom_step!(a,b)`` — reference src/Flow.jl:153-169.

    The host evaluates exactly what the reference evaluates on the host — ``BCTuple(a.U,a.Δt,N)``
    (:154) and the uniform accelerations of the two ``accelerate!`` calls (:158,165) — and makes
    one C-ABI call; both pushes (``a.Δt`` and two entries of ``b.n``) happen here."""
    dts = np.asarray(a.dt, dtype=np.float64)
    t0, t1 = float(dts[:-1].sum()), float(dts.sum())
    U = BCTuple(a.U, a.dt, a.D)
    gp = gc = None
    if a.g is not None or callable(a.U):
        def acc(i, t):
            v = float(a.g(i, t)) if a.g is not None else 0.0
            if callable(a.U):
                v += _ddt(a.U, i, t)
            return v
        gp = [acc(i, t0) for i in range(1, a.D + 1)]
        gc = [acc(i, t1) for i in range(1, a.D + 1)]
    dt_next, n1, n2 = a._eng.mom_step(U, gp, gc)
    a.dt.append(np.float32(dt_next))
    b.n.extend([n1, n2])


def sim_step(sim, t_end=None, *, remeasure=True, max_steps=None, verbose=False):
    """``sim_step!(sim,t_end;remeasure,max_steps,verbose)`` / ``sim_step!(sim;remeasure)`` —
    reference src/WaterLily.jl:98-109."""
    if t_end is None:
        if remeasure:
            measure(sim)
        mom_step(sim.flow, sim.pois)
        return
    steps0 = len(sim.flow.dt)
    while sim_time(sim) < t_end and (max_steps is None or len(sim.flow.dt) - steps0 < max_steps):
        sim_step(sim, remeasure=remeasure)
        if verbose:
            print(f"tU/L={sim_time(sim):.4f}, Δt={float(sim.flow.dt[-1]):.3f}")


# ---- Metrics: surface-integrated loads (reference src/Metrics.jl:94-141) -------------------------
def _nds(sim, t=None):
    from .body import nds_field
    t = sim.flow.time() if t is None else t
    return nds_field(sim.body, sim._eng.Ng, t=t)


def pressure_force(sim):
    """``pressure_force(sim)`` — reference src/Metrics.jl:94-100 (flow.f is overwritten, as there)."""
    return sim._eng.pressure_force(_nds(sim))


def viscous_force(sim):
    """``viscous_force(sim)`` — reference src/Metrics.jl:112-118."""
    return sim._eng.viscous_force(_nds(sim))


def total_force(sim):
    """``total_force(sim)`` — reference src/Metrics.jl:125."""
    nds = _nds(sim)
    return sim._eng.pressure_force(nds) + sim._eng.viscous_force(None)


def pressure_moment(x0, sim):
    """``pressure_moment(x₀,sim)`` — reference src/Metrics.jl:133-141."""
    return sim._eng.pressure_moment(x0, _nds(sim))


# ---- Metrics: derived fields (reference src/Metrics.jl:14-77), `@inside sim.flow.σ[I] = f(I,sim.flow.u)` -------
def ke_field(sim, U=None):
    """``@inside σ[I] = ke(I,u,U)`` — reference src/Metrics.jl:20-22."""
    return sim._eng.derived("ke", U=U)


def curl_field(sim, i):
    """``@inside σ[I] = curl(i,I,u)`` — reference src/Metrics.jl:55 (README.md:97 uses i=3 in 2-D)."""
    return sim._eng.derived("curl", comp=i)


def omega_mag_field(sim):
    """``@inside σ[I] = ω_mag(I,u)`` — reference src/Metrics.jl:67."""
    return sim._eng.derived("omega_mag")


def lambda2_field(sim):
    """``@inside σ[I] = λ₂(I,u)`` — reference src/Metrics.jl:40-44."""
    return sim._eng.derived("lambda2")


def omega_theta_field(sim, z, center):
    """``@inside σ[I] = ω_θ(I,z,center,u)`` — reference src/Metrics.jl:74-78."""
    return sim._eng.derived("omega_theta", axis=z, center=center)
