"""Size-independent properties of the path, checked on the CPU oracle (the checker must have them before it is used to judge
the CUDA kernels): conservation of the flux form, symmetry of the Poisson operator, sum-preserving restriction, idempotent
boundary conditions, divergence-free projection.  No GPU."""
import numpy as np
import pytest

from oracle import oracle
from waterlily_b200 import _cabi


def _inner(D):
    return tuple(slice(1, -1) for _ in range(D))


def _rand(shape, seed):
    return (np.random.default_rng(seed).random(shape) * 2 - 1).astype(np.float32)


@pytest.mark.parametrize("dims", [(16, 12), (12, 8, 10)])
def test_conv_diff_conserves_momentum_on_a_periodic_box(dims):
    # conv_diff! is in flux form: every face flux enters one cell and leaves its neighbour (Flow.jl:45-47, periodic ends :58-60),
    # so the interior sum of r[:,i] telescopes to zero — up to Float32 rounding of the cell sums
    D = len(dims)
    per = tuple(range(1, D + 1))
    e = oracle.engine(dims, nu=0.07, perdir=per, pois_kind=_cabi.POIS_SINGLE)
    Ng = tuple(n + 2 for n in dims)
    e.upload("u", np.asfortranarray(_rand(Ng + (D,), 1)))
    e.bc("u", (0.0,) * D, perdir=per)
    e.conv_diff("u")
    r = e.download("f").astype(np.float64)
    for i in range(D):
        ri = r[_inner(D) + (i,)]
        assert abs(ri.sum()) < 1e-5 * np.abs(ri).sum(), (i, ri.sum(), np.abs(ri).sum())


@pytest.mark.parametrize("dims,perdir", [((16, 16), ()), ((8, 8, 16), ()), ((8, 8, 8), (1, 2, 3))])
def test_poisson_operator_is_symmetric(dims, perdir):
    # mult (Poisson.jl:69-75) with D = -Σ(L_lower + L_upper): x·(A y) == y·(A x) for fields with consistent ghosts
    D = len(dims)
    Ng = tuple(n + 2 for n in dims)
    e = oracle.engine(dims, perdir=perdir, pois_kind=_cabi.POIS_SINGLE)
    mu0 = (0.1 + 0.9 * np.random.default_rng(3).random(Ng + (D,))).astype(np.float32)
    e.upload("mu0", np.asfortranarray(mu0))
    e.bc("mu0", (0.0,) * D, perdir=perdir)                      # BC!(L,0): closed walls on the non-periodic sides
    e.pois_update()
    fields = []
    for seed in (5, 6):
        x = np.zeros(Ng, dtype=np.float32)
        x[_inner(D)] = _rand(dims, seed)
        fields.append(x)
    Ax = []
    for x in fields:
        e.upload("p", np.asfortranarray(x))
        e.pois_mult()                                           # perBC!(x); z = A x
        Ax.append(e.download("sigma").astype(np.float64)[_inner(D)])
    x0, x1 = (f.astype(np.float64)[_inner(D)] for f in fields)
    a, b = (x0 * Ax[1]).sum(), (x1 * Ax[0]).sum()
    assert abs(a - b) < 1e-5 * (np.abs(x0 * Ax[1]).sum() + 1e-30), (a, b)
    # A annihilates constants: every row of A sums to zero (Neumann / periodic problem, Poisson.jl:42-46)
    e.upload("p", np.asfortranarray(np.ones(Ng, dtype=np.float32)))
    if not perdir:
        one = np.zeros(Ng, dtype=np.float32)
        one[_inner(D)] = 1
        e.upload("p", np.asfortranarray(one))
    e.pois_mult()
    assert np.abs(e.download("sigma")[_inner(D)]).max() < 1e-5


@pytest.mark.parametrize("perdir", [(), (1, 2, 3)])
def test_restriction_preserves_the_residual_sum(perdir):
    # restrict! sums the 2^D fine residuals of a coarse cell (MultiLevelPoisson.jl:33): Σ r_coarse == Σ r_fine
    dims = (16, 8, 8)
    Ng = tuple(n + 2 for n in dims)
    e = oracle.engine(dims, perdir=perdir, pois_kind=_cabi.POIS_MULTILEVEL)
    e.upload("mu0", np.asfortranarray(np.ones(Ng + (3,), dtype=np.float32)))
    e.bc("mu0", (0.0, 0.0, 0.0), perdir=perdir)
    e.pois_update()
    x = np.zeros(Ng, dtype=np.float32)
    x[_inner(3)] = _rand(dims, 7)
    z = np.zeros(Ng, dtype=np.float32)
    zi = _rand(dims, 8).astype(np.float64)
    z[_inner(3)] = (zi - zi.mean()).astype(np.float32)
    e.upload("p", np.asfortranarray(x))
    e.upload("sigma", np.asfortranarray(z))
    e.pois_head()                                               # residual! + Jacobi! + restrict!
    rf = e.level_download(0, "r").astype(np.float64)[_inner(3)]
    rc = e.level_download(1, "r").astype(np.float64)[_inner(3)]
    assert rc.shape == tuple(n // 2 for n in dims)
    assert abs(rc.sum() - rf.sum()) < 1e-5 * np.abs(rf).sum()
    blocks = rf.reshape(8, 2, 4, 2, 4, 2).sum(axis=(1, 3, 5))
    assert np.abs(blocks - rc).max() < 1e-5 * np.abs(rf).max()
    assert not e.level_download(1, "x").any()                   # fill!(coarse.x, 0) (MultiLevelPoisson.jl:75)


@pytest.mark.parametrize("dims,perdir", [((8, 12), ()), ((8, 12), (2,)), ((6, 8, 10), (1, 3))])
@pytest.mark.parametrize("saveexit", [False, True])
def test_boundary_conditions_are_idempotent(dims, perdir, saveexit):
    D = len(dims)
    Ng = tuple(n + 2 for n in dims)
    e = oracle.engine(dims, perdir=perdir, pois_kind=_cabi.POIS_SINGLE)
    e.upload("u", np.asfortranarray(_rand(Ng + (D,), 11)))
    A = (1.0, 0.5, -0.25)[:D]
    e.bc("u", A, saveexit=saveexit, perdir=perdir)
    once = e.download("u").copy()
    e.bc("u", A, saveexit=saveexit, perdir=perdir)
    assert np.array_equal(e.download("u"), once)
    e.upload("p", np.asfortranarray(_rand(Ng, 12)))
    e.per_bc("p", perdir)
    p1 = e.download("p").copy()
    e.per_bc("p", perdir)
    assert np.array_equal(e.download("p"), p1)
    for j in perdir:                                            # a[1] = a[N-1]; a[N] = a[2] (util.jl:227-231)
        lo = [slice(None)] * D
        hi = [slice(None)] * D
        lo[j - 1], hi[j - 1] = 0, -2
        assert np.array_equal(p1[tuple(lo)], p1[tuple(hi)])
        lo[j - 1], hi[j - 1] = -1, 1
        assert np.array_equal(p1[tuple(lo)], p1[tuple(hi)])


@pytest.mark.parametrize("dims,perdir", [((16, 16), ()), ((8, 8, 8), (1, 2, 3))])
def test_projection_leaves_a_divergence_free_field(dims, perdir):
    # project! (Flow.jl:137-145): after solver!(tol=1e-4) the residual of the pressure equation, i.e. the divergence of the
    # corrected velocity (times dt), satisfies Σ r² < 1e-4
    D = len(dims)
    Ng = tuple(n + 2 for n in dims)
    e = oracle.engine(dims, perdir=perdir, dt0=0.25)
    e.upload("mu0", np.asfortranarray(np.ones(Ng + (D,), dtype=np.float32)))
    e.bc("mu0", (0.0,) * D, perdir=perdir)
    e.pois_update()
    e.upload("u", np.asfortranarray(0.1 * _rand(Ng + (D,), 21)))
    e.bc("u", (0.0,) * D, perdir=perdir)
    n = e.project(1.0)
    assert 1 <= n <= 32
    e.bc("u", (0.0,) * D, perdir=perdir)
    u = e.download("u").astype(np.float64)
    div = np.zeros(dims)
    for i in range(D):
        up = [slice(1, -1)] * D
        up[i] = slice(2, None)
        div += u[tuple(up) + (i,)] - u[_inner(D) + (i,)]
    assert (div ** 2).sum() < 1.05e-4, (div ** 2).sum()          # = L₂(pois) of the last iteration (< tol) + rounding of the correction
    assert 0 < e.cfl() <= 10.0                                  # CFL (Flow.jl:172-175): Δt = min(10, 1/(max σ + 5ν))
