"""Parity of the CUDA path (through the C ABI) against the CPU oracle, function by function and for
whole mom_step! runs, on identical seeded inputs.  Tolerances are the north star's: per-field
relative L2 ≤ 1e-5 after one step, ≤ 1e-4 after many, V-cycle counts within ±1."""
import numpy as np
import pytest
import torch

from oracle import oracle
from oracle.oracle import OracleSimulation
from waterlily_b200 import _cabi, lib, sim_step, AutoBody, Simulation
from helpers import L2, poisson_setup, rel_l2

pytestmark = pytest.mark.gpu

FIELDS_V = ("u", "u0", "f", "V", "mu0")


def cuda_engine(dims, **kw):
    return _cabi.Engine(lib.load(), dims, **kw)


def pair(dims, perdir=(), exitBC=False, nu=0.01, dt=0.25, kind=_cabi.POIS_SINGLE, seed=0, random_fields=("u", "u0", "f", "p", "sigma", "V")):
    eo = oracle.engine(dims, nu=nu, dt0=dt, perdir=perdir, exitBC=exitBC, pois_kind=kind)
    ec = cuda_engine(dims, nu=nu, dt0=dt, perdir=perdir, exitBC=exitBC, pois_kind=kind)
    rng = np.random.default_rng(seed)
    for name in random_fields:
        a = (rng.random(eo.shape(_cabi.FIELD_IDS[name])) * 2 - 1).astype(np.float32)
        eo.upload(name, a)
        ec.upload(name, a)
    return eo, ec


DIMS = [((8, 12), ()), ((8, 12), (1,)), ((8, 12), (2,)), ((8, 12), (1, 2)),
        ((8, 6, 10), ()), ((8, 6, 10), (3,)), ((8, 6, 10), (1, 2, 3)), ((8, 6, 10), (2,))]


@pytest.mark.parametrize("dims,perdir", DIMS)
@pytest.mark.parametrize("saveexit", [False, True])
def test_bc_bit_exact(dims, perdir, saveexit):
    eo, ec = pair(dims, perdir)
    A = (1.0, 0.5, -0.25)[:len(dims)]
    for e in (eo, ec):
        e.bc("u", A, saveexit, None)
    assert np.array_equal(eo.download("u"), ec.download("u"))
    for e in (eo, ec):
        e.per_bc("sigma", None)
    assert np.array_equal(eo.download("sigma"), ec.download("sigma"))


@pytest.mark.parametrize("dims", [(8, 12), (8, 6, 10)])
def test_exit_bc(dims):
    eo, ec = pair(dims, exitBC=True)
    U = (1.0, 0.0, 0.0)[:len(dims)]
    for e in (eo, ec):
        e.exit_bc(U, 0.3)
    assert np.allclose(eo.download("u"), ec.download("u"), rtol=0, atol=2e-7)


@pytest.mark.parametrize("dims,perdir", DIMS + [((32, 32, 32), (1, 2, 3)), ((48, 24, 24), ())])
def test_conv_diff(dims, perdir):
    eo, ec = pair(dims, perdir, nu=0.05)
    for e in (eo, ec):
        e.conv_diff("u")
    fo, fc = eo.download("f"), ec.download("f")
    assert rel_l2(fc, fo) < 2e-7
    assert np.abs(fc - fo).max() < 5e-6
    # σ ≡ Φ: the ghost layer must hold the same stale fluxes; the interior is scratch
    so, sc = eo.download("sigma"), ec.download("sigma")
    shell = np.ones(so.shape, bool)
    shell[tuple(slice(1, -1) for _ in dims)] = False
    assert np.abs(so[shell] - sc[shell]).max() < 5e-6


@pytest.mark.parametrize("dims,perdir", [((8, 12), ()), ((8, 6, 10), (3,))])
def test_bdim_accelerate_scale(dims, perdir):
    D = len(dims)
    eo, ec = pair(dims, perdir, random_fields=("u", "u0", "f", "V"))
    rng = np.random.default_rng(5)
    mu0 = rng.random(eo.shape(_cabi.MU0)).astype(np.float32)
    mu1 = (rng.random(eo.shape(_cabi.MU1)) - 0.5).astype(np.float32)
    for e in (eo, ec):
        e.upload("mu0", mu0)
        e.upload("mu1", mu1)
        e.accelerate((0.3, -0.2, 0.1)[:D])
        e.bdim()
        e.scale_u(0.5)
    assert rel_l2(ec.download("f"), eo.download("f")) < 2e-7
    assert rel_l2(ec.download("u"), eo.download("u")) < 2e-7


def _random_pois(dims, perdir, kind, seed=3):
    eo, ec = pair(dims, perdir, kind=kind, seed=seed, random_fields=("p", "sigma"))
    rng = np.random.default_rng(seed + 1)
    L = rng.random(eo.shape(_cabi.MU0)).astype(np.float32)
    # solid patch: L=0 makes D=0 → iD=0 cells
    L[tuple(slice(2, 4) for _ in dims)] = 0
    for e in (eo, ec):
        e.upload("mu0", L)
        e.bc("mu0", (0.0,) * len(dims), False, None)
        e.pois_update()
    return eo, ec


@pytest.mark.parametrize("dims,perdir", [((16, 32), ()), ((16, 32), (1,)), ((16, 8, 24), ()), ((16, 8, 24), (1, 2, 3))])
def test_poisson_pieces(dims, perdir):
    eo, ec = _random_pois(dims, perdir, _cabi.POIS_MULTILEVEL)
    assert eo.nlevels() == ec.nlevels() >= 3
    for l in range(eo.nlevels()):
        assert eo.level_dims(l) == ec.level_dims(l)
        for f in ("L", "D", "iD"):
            assert np.allclose(ec.level_download(l, f), eo.level_download(l, f), rtol=1e-6, atol=1e-7), (l, f)
    r2o, r2c = eo.pois_residual(0), ec.pois_residual(0)
    assert abs(r2c - r2o) <= 1e-5 * abs(r2o)
    assert rel_l2(ec.level_download(0, "r"), eo.level_download(0, "r")) < 1e-6
    for e in (eo, ec):
        e.pois_jacobi(0, 2)
    for f in ("r", "x", "eps"):
        assert rel_l2(ec.level_download(0, f), eo.level_download(0, f)) < 1e-6, f
    for e in (eo, ec):
        e.pois_pcg(0, 6)
    for f in ("r", "x", "eps"):
        assert rel_l2(ec.level_download(0, f), eo.level_download(0, f)) < 1e-5, f
    for e in (eo, ec):
        e.pois_vcycle(0)
    for l in range(eo.nlevels()):
        for f in ("r", "x"):
            a, b = ec.level_download(l, f), eo.level_download(l, f)
            # coarse residuals are tiny after pcg!: measure the error against the level's source scale
            scale = max(np.abs(eo.level_download(l, "x")).max(), np.abs(b).max(), 1e-30) if f == "r" else np.abs(b).max()
            assert np.abs(a - b).max() <= 2e-5 * scale, (l, f, np.abs(a - b).max(), scale)
    for e in (eo, ec):
        e.pois_mult()
    assert rel_l2(ec.download("sigma"), eo.download("sigma")) < 1e-6


@pytest.mark.parametrize("kind,Ng,nmax", [(_cabi.POIS_SINGLE, (5, 5), 10**6), (_cabi.POIS_SINGLE, (66, 66), 310), (_cabi.POIS_SINGLE, (18, 18, 18), 35),
                                          (_cabi.POIS_MULTILEVEL, (66, 66), 4), (_cabi.POIS_MULTILEVEL, (18, 18, 18), 4)])
def test_reference_poisson_pins_on_cuda(kind, Ng, nmax):
    # test/maintests.jl:81-94,109-116 through the CUDA library
    err, n, _ = poisson_setup(cuda_engine, kind, Ng)
    erro, no, _ = poisson_setup(oracle.engine, kind, Ng)
    assert err < (1e-5 if Ng == (5, 5) else 1e-6) and n < nmax
    assert abs(n - no) <= 1


def test_reference_multilevel_pins_on_cuda():
    # test/maintests.jl:101-107
    err, n, e = poisson_setup(cuda_engine, _cabi.POIS_MULTILEVEL, (10, 10))
    D3 = np.array([[0, 0, 0, 0], [0, -2, -2, 0], [0, -2, -2, 0], [0, 0, 0, 0]], dtype=np.float32)
    assert (e.level_download(2, "D") == D3).all() and err < 1e-5
    L = e.level_download(0, "L")
    L[4:6, :, 0] = 0
    e.level_upload(0, "L", L)
    e.pois_update()
    assert (e.level_download(2, "D") == D3 / 2).all()


def compare_sims(so, sc, tol, fields=("u", "p", "f")):
    out = {}
    for f in fields:
        a, b = getattr(sc.flow, f).numpy(), getattr(so.flow, f).numpy()
        out[f] = rel_l2(a, b)
        assert out[f] <= tol, (f, out[f])
    return out


def run_pair(make, steps, tol1=1e-5, tolN=1e-4):
    so, sc = make(OracleSimulation), make(Simulation)
    compare_sims(so, sc, 1e-7, fields=("u", "mu0"))
    for k in range(steps):
        sim_step(so, remeasure=False)
        sim_step(sc, remeasure=False)
        compare_sims(so, sc, tol1 if k == 0 else tolN)
        assert abs(float(so.flow.dt[-1]) - float(sc.flow.dt[-1])) <= tolN * float(so.flow.dt[-1])
    assert len(so.pois.n) == len(sc.pois.n)
    assert all(abs(a - b) <= 1 for a, b in zip(so.pois.n, sc.pois.n)), (so.pois.n, sc.pois.n)
    return so, sc


def tgv3d(cls, N, Re=1600.0):
    # BASELINE.json config "3D Taylor-Green vortex N³ periodic" 
    kappa = 2 * np.pi / N
    nu = 1 / (kappa * Re)

    def ulam(i, x):
        X, Y, Z = x[..., 0] * kappa, x[..., 1] * kappa, x[..., 2] * kappa
        if i == 1:
            return -np.sin(X) * np.cos(Y) * np.cos(Z)
        if i == 2:
            return np.cos(X) * np.sin(Y) * np.cos(Z)
        return np.zeros_like(X)
    return cls((N, N, N), (0, 0, 0), N, U=1, nu=nu, ulam=ulam, perdir=(1, 2, 3))


def circle2d(cls, n=3 * 2 ** 5, m=2 ** 6, Re=100):
    # README example
    radius, center = m / 8, m / 2 - 1
    sdf = lambda x, t: torch.sqrt(((x - center) ** 2).sum(-1)) - radius
    return cls((n, m), (1, 0), 2 * radius, nu=2 * radius / Re, body=AutoBody(sdf))


def sphere3d(cls, n=64, m=32, exitBC=True):
    # SURVEY.md §8d C3 at reduced size
    radius, center = m / 8, m / 2 - 1
    sdf = lambda x, t: torch.sqrt(((x - center) ** 2).sum(-1)) - radius
    return cls((n, m, m), (1, 0, 0), 2 * radius, nu=2 * radius / 3700, body=AutoBody(sdf), exitBC=exitBC)


def test_mom_step_circle_2d():
    run_pair(circle2d, 20)


def test_mom_step_tgv_32():
    run_pair(lambda cls: tgv3d(cls, 32), 6)


def test_mom_step_tgv_64():
    run_pair(lambda cls: tgv3d(cls, 64), 3)


@pytest.mark.parametrize("exitBC", [True, False])
def test_mom_step_sphere(exitBC):
    run_pair(lambda cls: sphere3d(cls, exitBC=exitBC), 5)


def test_mom_step_100_circle():
    # north star: ≤1e-4 after 100 steps
    so, sc = circle2d(OracleSimulation), circle2d(Simulation)
    for _ in range(100):
        sim_step(so, remeasure=False)
        sim_step(sc, remeasure=False)
    compare_sims(so, sc, 1e-4, fields=("u", "p"))
    assert sum(abs(a - b) for a, b in zip(so.pois.n, sc.pois.n)) <= 2


def test_accelerating_flow_and_g():
    # test/maintests.jl:280-302 on the CUDA path: g(i,t) and periodic x
    N, jerk = 8, 4
    make = lambda cls: cls((N, N), (np.sqrt(N), 0.0), N, nu=0.001, g=lambda i, t: t * jerk if i == 1 else 0.0, dt=0.001, perdir=(1,))
    so, sc = make(OracleSimulation), make(Simulation)
    sim_step(so, 1.0)
    sim_step(sc, 1.0)
    u = sc.flow.u.numpy()
    uFinal = np.sqrt(N) + 0.5 * jerk * sc.flow.time() ** 2
    assert L2(u[:, :, 0] - uFinal) < 1e-4 and L2(u[:, :, 1]) < 1e-4
    assert rel_l2(u, so.flow.u.numpy()) < 1e-5


def test_accelerating_circle_added_mass_on_cuda():
    # test/maintests.jl:304-316 on the CUDA path (1024² grid, time-dependent u_BC, pressure_force), and against the oracle
    from helpers import check_accelerating_circle
    sc = check_accelerating_circle(Simulation)
    so = check_accelerating_circle(OracleSimulation)
    assert all(abs(a - b) <= 1 for a, b in zip(sc.pois.n, so.pois.n)), (sc.pois.n, so.pois.n)
    for f in ("u", "p"):
        assert rel_l2(getattr(sc.flow, f).numpy(), getattr(so.flow, f).numpy()) < 1e-4, f


def test_moving_body_pins_on_cuda():
    # test/maintests.jl:398-400: accelerating circle, pois.n == [2,1]
    radius = 8
    e1 = torch.tensor([1.0, 0.0])
    circle = lambda x, t: torch.sqrt(((x - 2 * radius) ** 2).sum(-1)) - radius
    accel = lambda x, t: x - e1 * (2 * t ** 2)
    for exitBC in (True, False):
        sim = Simulation((4 * radius, 4 * radius), (0, 0), radius, U=1, body=AutoBody(circle, accel), nu=radius / 250, exitBC=exitBC)
        sim_step(sim)
        assert sim.pois.n == [2, 1]
        assert sim.flow.u.numpy().max() > sim.flow.V.numpy().max() > 0


def test_full_size_properties_tgv128():
    # BASELINE config C2 at full size: size-independent properties (the oracle is too slow to step 128³ in a test)
    sim = tgv3d(Simulation, 128)
    for _ in range(3):
        sim_step(sim, remeasure=False)
    u = sim.flow.u.numpy().astype(np.float64)
    div = (u[2:, 1:-1, 1:-1, 0] - u[1:-1, 1:-1, 1:-1, 0]) + (u[1:-1, 2:, 1:-1, 1] - u[1:-1, 1:-1, 1:-1, 1]) + (u[1:-1, 1:-1, 2:, 2] - u[1:-1, 1:-1, 1:-1, 2])
    assert (div ** 2).sum() < 5e-4          # solver tolerance: Σr² < 1e-4 per projection (scaled by dt)
    assert np.array_equal(u[0], u[-2]) and np.array_equal(u[-1], u[1])   # periodic ghosts
    ke = 0.5 * (u[1:-1, 1:-1, 1:-1] ** 2).sum() / 128 ** 3
    assert 0.115 < ke < 0.126              # TGV kinetic energy 1/8 at t=0, decaying slowly
    assert all(1 <= n <= 32 for n in sim.pois.n)
    assert 0.5 < float(sim.flow.dt[-1]) < 1.1


@pytest.mark.parametrize("solve_cells", ["0", "1000000000"])   # coarse tail only / the whole solver! in one kernel
@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 64), 3), (lambda cls: sphere3d(cls, exitBC=True), 4), (circle2d, 10)])
def test_fused_vcycle_matches_per_level_kernels(make, steps, solve_cells, monkeypatch):
    # wl_vcycle.cu (one persistent kernel for the coarse levels, or for the whole solver!) against the
    # per-function kernels of wl_poisson.cu: same arithmetic, different reduction tree → agreement to
    # reduction-order noise
    monkeypatch.setenv("WL_FUSED_VCYCLE", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_FUSED_VCYCLE", "1")
    monkeypatch.setenv("WL_FUSED_SOLVE_CELLS", solve_cells)
    monkeypatch.setenv("WL_FUSED_MARCH_CELLS", "0")   # z-marching / staged tile phases inside the persistent kernel even on these small grids
    s1 = make(Simulation)
    l0, l1 = s0._eng.launch_count(), s1._eng.launch_count()
    for _ in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
    assert s0.pois.n == s1.pois.n
    for f in ("u", "p"):
        assert rel_l2(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()) < 2e-6, f
    assert s1._eng.launch_count() - l1 < s0._eng.launch_count() - l0   # the fused path really ran


def test_mom_step_tgv_128_full_config():
    # BASELINE.json configs[1] at its full size: 3-D TGV 128³ periodic, two steps
    run_pair(lambda cls: tgv3d(cls, 128), 2)


def test_mom_step_sphere_full_config():
    # BASELINE.json configs[2] at its full size: sphere 2^8 x 2^7 x 2^7, Re=3700, exitBC
    run_pair(lambda cls: sphere3d(cls, n=256, m=128, exitBC=True), 2)


def test_metrics_reference_pins_on_cuda():
    # maintests.jl:341-368 through the CUDA library, and the same numbers as the oracle
    from test_oracle_pins import check_force_routines
    fc = check_force_routines(cuda_engine)
    fo = check_force_routines(oracle.engine)
    assert np.allclose(fc, fo, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("make", [lambda cls: sphere3d(cls, exitBC=True), circle2d])
def test_metrics_on_a_running_flow(make):
    # pressure_force / viscous_force / pressure_moment (src/Metrics.jl:94-141) of a developing wake
    from waterlily_b200 import pressure_force, viscous_force, total_force, pressure_moment
    so, sc = make(OracleSimulation), make(Simulation)
    for _ in range(5):
        sim_step(so, remeasure=False)
        sim_step(sc, remeasure=False)
    x0 = tuple(float(n) / 2 for n in sc._eng.N)
    for fn in (pressure_force, viscous_force, total_force, lambda s: pressure_moment(x0, s)):
        a, b = fn(sc), fn(so)
        scale = max(np.abs(b).max(), 1e-12)
        assert np.abs(a - b).max() <= 2e-4 * scale, (a, b)
    assert np.abs(pressure_force(so)).max() > 1e-3   # the comparison is not vacuous


def test_derived_fields_on_cuda():
    # Metrics.jl diagnostics (maintests.jl:320-337) through the CUDA library, and against the oracle on a developed flow
    from test_oracle_pins import check_derived_fields
    fc, fo = check_derived_fields(cuda_engine), check_derived_fields(oracle.engine)
    for k in fc:
        assert np.allclose(fc[k][1:-1, 1:-1, 1:-1], fo[k][1:-1, 1:-1, 1:-1], rtol=1e-5, atol=1e-5), k
    so, sc = tgv3d(OracleSimulation, 32), tgv3d(Simulation, 32)
    for _ in range(3):
        sim_step(so, remeasure=False)
        sim_step(sc, remeasure=False)
    inside = (slice(1, -1),) * 3
    for kind, kw in (("ke", {}), ("ke", {"U": (0.1, -0.2, 0.3)}), ("curl", {"comp": 1}), ("curl", {"comp": 3}), ("omega_mag", {}),
                     ("lambda2", {}), ("omega_theta", {"axis": (0, 0, 1), "center": (16, 16, 16)})):
        a, b = sc._eng.derived(kind, **kw)[inside], so._eng.derived(kind, **kw)[inside]
        assert np.abs(a - b).max() <= 1e-4 * max(np.abs(b).max(), 1e-6), (kind, np.abs(a - b).max(), np.abs(b).max())


@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 32), 3), (lambda cls: sphere3d(cls, exitBC=True), 3), (circle2d, 6)])
@pytest.mark.parametrize("solve_cells", ["0", "1000000000"])
def test_uniform_coefficient_packs_are_bit_identical(make, steps, solve_cells, monkeypatch):
    # Lvl::umask (wl_common.cuh): packs whose face coefficients all equal the level constant use constants for
    # L, D, iD instead of loading them — same operands, same operations: every field must come out bit for bit
    # the same as with the fast path switched off (WL_UNIFORM=0), in the per-level kernels and in the fused solver
    monkeypatch.setenv("WL_FUSED_SOLVE_CELLS", solve_cells)
    monkeypatch.setenv("WL_UNIFORM", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_UNIFORM", "1")
    s1 = make(Simulation)
    for _ in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
    assert s0.pois.n == s1.pois.n and s0.flow.dt == s1.flow.dt
    for f in ("u", "p", "f", "sigma"):
        assert np.array_equal(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()), f


def test_uniform_mask_goes_stale_with_L():
    # changing L (≡ μ₀) without update! must fall back to loading L, D, iD (the reference would use the new L with
    # the old D); update! re-validates the mask
    eo, ec = pair((16, 16, 16), perdir=(1, 2, 3), kind=_cabi.POIS_MULTILEVEL, random_fields=("p", "sigma"))
    rng = np.random.default_rng(3)
    L = np.ones(eo.shape(_cabi.FIELD_IDS["mu0"]), dtype=np.float32, order="F")
    L[5:9, 4:7, 3:11, :] = rng.random((4, 3, 8, 3)).astype(np.float32)
    for e in (eo, ec):
        e.level_upload(0, "L", L)          # no update!: D, iD still describe L = 1
        e.pois_residual(0)
    assert rel_l2(ec.level_download(0, "r"), eo.level_download(0, "r")) < 1e-6
    for e in (eo, ec):
        e.pois_update()
        e.pois_residual(0)
    assert rel_l2(ec.level_download(0, "r"), eo.level_download(0, "r")) < 1e-6
    assert np.array_equal(ec.level_download(0, "D"), eo.level_download(0, "D"))


def _head_case(cls_engine, dims, perdir, mu0, x0, z):
    e = cls_engine(dims, perdir=perdir, pois_kind=_cabi.POIS_MULTILEVEL)
    e.upload("mu0", np.asfortranarray(mu0))
    e.bc("mu0", (0.0, 0.0, 0.0))
    e.pois_update()
    e.upload("p", np.asfortranarray(x0))
    e.per_bc("p", None)
    e.upload("sigma", np.asfortranarray(z))
    return e


def _head_fields(dims, uniform, mean, seed=3):
    rng = np.random.default_rng(seed)
    Ng = tuple(n + 2 for n in dims)
    mu0 = np.ones(Ng + (3,), dtype=np.float32) if uniform else (0.2 + 0.8 * rng.random(Ng + (3,))).astype(np.float32)
    x0 = (rng.random(Ng) * 2 - 1).astype(np.float32)
    z = (rng.random(Ng) * 2 - 1).astype(np.float64)
    z[1:-1, 1:-1, 1:-1] -= z[1:-1, 1:-1, 1:-1].mean() - mean
    return mu0, x0, z.astype(np.float32)


@pytest.mark.parametrize("dims", [(128, 16, 16), (256, 24, 8)])
@pytest.mark.parametrize("perdir", [(), (1, 2, 3), (2,), (1, 3)])
@pytest.mark.parametrize("uniform", [False, True])
def test_fused_solver_head_is_bit_identical(dims, perdir, uniform, monkeypatch):
    # wl_resjac.cu: residual! + Jacobi! + restrict! + fill!(coarse.x,0) as one pass (several tiles in x, y and z) against
    # the four separate kernels: same operations in the same order per cell → the same bits in r, x and the coarse r
    # (right-hand side with zero mean: residual!'s mean subtraction does not fire)
    mu0, x0, z = _head_fields(dims, uniform, 0.0)
    out = []
    for which in ("0", "1"):
        monkeypatch.setenv("WL_RESJAC", which)
        e = _head_case(cuda_engine, dims, perdir, mu0, x0, z)
        e.pois_head()
        out.append((e.level_download(0, "r"), e.download("p"), e.level_download(1, "r"), e.level_download(1, "x")))
    inner = (slice(1, -1),) * 3
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a[inner], b[inner])
    eo = _head_case(oracle.engine, dims, perdir, mu0, x0, z)
    eo.pois_head()
    assert rel_l2(out[1][0][inner], eo.level_download(0, "r")[inner]) < 2e-6
    assert rel_l2(out[1][1][inner], eo.download("p")[inner]) < 2e-6
    assert rel_l2(out[1][2][inner], eo.level_download(1, "r")[inner]) < 2e-6


@pytest.mark.parametrize("perdir", [(), (1, 2, 3), (2,)])
@pytest.mark.parametrize("uniform", [False, True])
def test_fused_solver_head_mean_subtraction(perdir, uniform, monkeypatch):
    # right-hand side with a (huge) mean: the fused pass runs without residual!'s subtraction and is corrected afterwards
    # by the exact linear identity (k_resjac_fix) — equal to the separate kernels and the oracle up to rounding
    dims = (128, 16, 16)
    mu0, x0, z = _head_fields(dims, uniform, 0.1)
    inner = (slice(1, -1),) * 3
    got = []
    for which in ("oracle", "0", "1"):
        monkeypatch.setenv("WL_RESJAC", "1" if which == "oracle" else which)
        e = _head_case(oracle.engine if which == "oracle" else cuda_engine, dims, perdir, mu0, x0, z)
        e.pois_head()
        got.append((e.level_download(0, "r")[inner], e.download("p")[inner], e.level_download(1, "r")[inner]))
    for k in range(3):
        assert rel_l2(got[2][k], got[1][k]) < 2e-6 and rel_l2(got[2][k], got[0][k]) < 2e-6, k


@pytest.mark.parametrize("perdir", [(), (1, 2, 3), (2,), (1, 3)])
@pytest.mark.parametrize("uniform", [False, True])
@pytest.mark.parametrize("mean", [0.0, 0.1])
def test_fused_solver_head_solve(perdir, uniform, mean, monkeypatch):
    # the whole solver! with the fused head against the separate kernels and against the oracle
    dims = (128, 16, 16)
    monkeypatch.setenv("WL_FUSED_SOLVE_CELLS", "0")   # per-level kernels on the finest level
    mu0, x0, z = _head_fields(dims, uniform, mean)
    out = []
    for which in ("oracle", "0", "1"):
        monkeypatch.setenv("WL_RESJAC", "1" if which == "oracle" else which)
        e = _head_case(oracle.engine if which == "oracle" else cuda_engine, dims, perdir, mu0, x0, z)
        l0 = e.launch_count() if which != "oracle" else 0
        n = e.pois_solve(itmx=2)
        out.append((n, e.download("p"), e.level_download(0, "r"), (e.launch_count() - l0) if which != "oracle" else 0))
    (no, xo, ro, _), (n0, xa, ra, la), (n1, xb, rb, lb) = out
    assert n0 == n1 == no
    tol = 2e-6 if mean == 0.0 else 2e-5   # with the mean correction the two paths round differently from the first sweep on
    assert rel_l2(xb, xa) < tol and rel_l2(xb, xo) < 2e-5 and rel_l2(xa, xo) < 2e-5
    assert lb < la   # the fused pass really ran


@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 128), 2), (lambda cls: sphere3d(cls, n=128, m=64, exitBC=True), 3)])
def test_fused_solver_head_mom_step(make, steps, monkeypatch):
    monkeypatch.setenv("WL_FUSED_SOLVE_CELLS", "0")
    monkeypatch.setenv("WL_RESJAC", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_RESJAC", "1")
    s1 = make(Simulation)
    for _ in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
    assert s0.pois.n == s1.pois.n
    for f in ("u", "p"):
        assert rel_l2(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()) < 2e-6, f


@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 128), 4), (lambda cls: sphere3d(cls, n=128, m=64, exitBC=True), 4)])
def test_head_takes_divided_pressure_and_div_u_on_the_fly_bit_identical(make, steps, monkeypatch):
    # inside mom_step! the fused head (wl_resjac.cu) evaluates z = div u itself and reads the pressure that the previous
    # pressure correction stored divided by dt, multiplying by the new dt while staging — against the stored z and the
    # whole-allocation scaling pass: same operations per cell, same bits.  Looking at p between two steps materialises the
    # pending division and must not disturb either path.
    monkeypatch.setenv("WL_FUSED_SOLVE_CELLS", "0")
    monkeypatch.setenv("WL_XDIV", "0")
    monkeypatch.setenv("WL_RESJAC_DIV", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_XDIV", "1")
    monkeypatch.setenv("WL_RESJAC_DIV", "1")
    s1 = make(Simulation)
    l0 = [s0._eng.launch_count(), s1._eng.launch_count()]
    for k in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
        if k == 1:
            assert np.array_equal(s1.flow.p.numpy(), s0.flow.p.numpy())
    assert s0.pois.n == s1.pois.n
    assert [float(x) for x in s0.flow.dt] == [float(x) for x in s1.flow.dt]
    for f in ("u", "p"):
        assert np.array_equal(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()), f
    assert s1._eng.launch_count() - l0[1] < s0._eng.launch_count() - l0[0]   # the two passes are really gone


@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 128), 2), (lambda cls: sphere3d(cls, exitBC=True), 4), (lambda cls: sphere3d(cls, n=128, m=64, exitBC=False), 3)])
def test_cfl_fused_into_pressure_correction_is_bit_identical(make, steps, monkeypatch):
    # k_project_correct_cfl + k_cfl_upper (the corrector's project! correction evaluates CFL's σ = flux_out, the corrected
    # field goes to the spare velocity buffer) against k_project_correct + k_cfl: same operations, same bits, same Δt
    monkeypatch.setenv("WL_CFL_FUSE", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_CFL_FUSE", "1")
    s1 = make(Simulation)
    for _ in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
    assert s0.pois.n == s1.pois.n
    assert [float(x) for x in s0.flow.dt] == [float(x) for x in s1.flow.dt]
    for f in ("u", "p", "sigma", "f"):
        assert np.array_equal(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()), f


@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 64), 3), (lambda cls: sphere3d(cls, exitBC=True), 4), (lambda cls: sphere3d(cls, n=128, m=64, exitBC=False), 2)])
@pytest.mark.parametrize("stage", ["0", "1"])   # z-window in registers / operand planes staged in shared memory by cp.async
def test_z_marching_kernels_match_row_kernels(make, steps, stage, monkeypatch):
    # k_*_m / k_*_s (for_columns, march_staged in wl_vec.cuh) against the (row-group, plane) kernels: same arithmetic
    # per cell, different reduction tree; the fused solver is switched off so that the per-level kernels run on every level
    monkeypatch.setenv("WL_FUSED_VCYCLE", "0")
    monkeypatch.setenv("WL_MARCH", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_MARCH", "1")
    monkeypatch.setenv("WL_STAGE", stage)
    s1 = make(Simulation)
    for _ in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
    assert s0.pois.n == s1.pois.n
    for f in ("u", "p"):
        assert rel_l2(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()) < 2e-6, f


@pytest.mark.parametrize("make,steps", [(lambda cls: tgv3d(cls, 32), 3), (lambda cls: sphere3d(cls, exitBC=True), 4), (lambda cls: sphere3d(cls, n=128, m=64, exitBC=False), 2)])
def test_bdim_stage2_folded_into_conv_diff_is_bit_identical(make, steps, monkeypatch):
    # On tile-planes away from every body the conv_diff! core kernel writes u = (u + f)·scale itself (CDArgs::ufused) and
    # BDIM! stage 2 skips those cells; the corrector assembles the new u in a spare buffer.  Same expressions → same bits,
    # including flow.f, which is recomputed on demand for the tiles that never stored it.
    monkeypatch.setenv("WL_BDIM_FUSE", "0")
    s0 = make(Simulation)
    monkeypatch.setenv("WL_BDIM_FUSE", "1")
    s1 = make(Simulation)
    for _ in range(steps):
        sim_step(s0, remeasure=False)
        sim_step(s1, remeasure=False)
        assert np.array_equal(s1.flow.u.numpy(), s0.flow.u.numpy())
    assert s0.pois.n == s1.pois.n and s0.flow.dt == s1.flow.dt
    inner = tuple(slice(1, -1) for _ in range(3))
    for f in ("u", "u0", "p", "sigma"):
        assert np.array_equal(getattr(s1.flow, f).numpy(), getattr(s0.flow, f).numpy()), f
    assert np.array_equal(s1.flow.f.numpy()[inner], s0.flow.f.numpy()[inner])
    # and stepping on after the on-demand f must not disturb anything
    sim_step(s0, remeasure=False)
    sim_step(s1, remeasure=False)
    assert np.array_equal(s1.flow.u.numpy(), s0.flow.u.numpy())


@pytest.mark.parametrize("dims,perdir", [((64, 32, 40), ()), ((64, 32, 32), (1, 2, 3)), ((136, 24, 16), (2,))])
def test_conv_diff_kernels_agree_bitwise(dims, perdir, monkeypatch):
    # the three 3-D core kernels of conv_diff! — untiled marching (WL_CD_TMA=0), TMA-staged one cell per thread, and the default
    # two cells per thread with packed FADD2/FMUL2/FFMA2 arithmetic (wl_flux.cuh: flux2) — must produce the same bits: every lane
    # of a packed instruction rounds like the scalar one, and no product is contracted into an fma the reference does not have
    outs = []
    for env in ({"WL_CD_TMA": "0"}, {"WL_CD_TMA": "1", "WL_CD_CELLS": "1"}, {"WL_CD_TMA": "1", "WL_CD_CELLS": "2"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ec = cuda_engine(dims, nu=0.013, dt0=0.25, perdir=perdir)
        rng = np.random.default_rng(11)
        for name in ("u", "u0", "V"):
            ec.upload(name, (rng.random(ec.shape(_cabi.FIELD_IDS[name])) * 2 - 1).astype(np.float32))
        ec.conv_diff("u")
        outs.append(ec.download("f"))
        ec.close()
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], outs[2])
