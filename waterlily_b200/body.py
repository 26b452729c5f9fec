"""Host-side body geometry: ``AbstractBody``, ``NoBody``, ``AutoBody`` and ``measure!``.

Mirrors reference src/Body.jl:17-76 and src/AutoBody.jl:1-56,101-131.  The north star keeps the
geometry on the host (the reference evaluates it in Julia with ForwardDiff); here the automatic
differentiation is ``torch.autograd`` / ``torch.func.jvp`` on CPU tensors, and the resulting
μ₀, μ₁, V arrays are uploaded through the C ABI when measured.  This is init-/remeasure-time
work, not part of the per-step CUDA path.
"""
from __future__ import annotations

import math

import numpy as np
import torch


class AbstractBody:
    """Reference src/Body.jl:17 — subtypes implement ``sdf(x,t)`` and ``measure(x,t)``."""

    def sdf(self, x, t=0.0, fastd2=math.inf):
        raise NotImplementedError

    def measure(self, x, t=0.0, fastd2=math.inf):
        raise NotImplementedError


class NoBody(AbstractBody):
    """Reference src/Body.jl:75-76."""


def _as_t(t, like):
    return torch.as_tensor(t, dtype=like.dtype)


class AutoBody(AbstractBody):
    """``AutoBody(sdf, map=(x,t)->x; compose=true)`` — reference src/AutoBody.jl:1-33.

    ``sdf(x,t)`` takes a tensor ``x`` of shape ``(..., D)`` and returns ``(...)``; ``map(x,t)``
    returns ``(..., D)``.  Both must be written with torch ops so they can be differentiated.
    """

    def __init__(self, sdf, map=None, compose=True):
        self._map = map if map is not None else (lambda x, t: x)
        self._has_map = map is not None
        comp = compose and map is not None
        self._sdf = (lambda x, t: sdf(self._map(x, t), t)) if comp else sdf

    def sdf(self, x, t=0.0, fastd2=math.inf):
        x = torch.as_tensor(x)
        return self._sdf(x, _as_t(t, x))

    def measure(self, x, t=0.0, fastd2=math.inf):
        """``d,n,V = measure(body,x,t;fastd²)`` — reference src/AutoBody.jl:115-131, batched."""
        x = torch.as_tensor(x).detach().clone().requires_grad_(True)
        tt = _as_t(t, x)
        d = self._sdf(x, tt)
        (n,) = torch.autograd.grad(d.sum(), x, allow_unused=True)
        if n is None:
            n = torch.zeros_like(x)
        d = d.detach()
        x = x.detach()
        skip = (d * d > fastd2) | torch.isnan(n).any(dim=-1)
        m = torch.sqrt((n * n).sum(dim=-1))
        m_safe = torch.where(skip | (m == 0), torch.ones_like(m), m)
        dn = torch.where(skip, d, d / m_safe)
        nn = torch.where(skip[..., None], torch.zeros_like(n), n / m_safe[..., None])
        V = torch.zeros_like(x)
        if self._has_map:
            # ẋ = -(∂m/∂x) \ ∂m/∂t
            J = torch.func.vmap(torch.func.jacrev(lambda y: self._map(y, tt)))(x.reshape(-1, x.shape[-1]))
            _, dot = torch.func.jvp(lambda s: self._map(x, s), (tt,), (torch.ones_like(tt),))
            Vall = -torch.linalg.solve(J, dot.reshape(-1, x.shape[-1], 1)).reshape(x.shape)
            V = torch.where(skip[..., None], V, Vall)
        return dn, nn, V

    # CSG (reference src/AutoBody.jl:35-45)
    def __add__(self, other):
        return _Combined(self, other, "+")

    def __sub__(self, other):
        return _Combined(self, other, "-")

    def __and__(self, other):
        return _Combined(self, other, "&")

    __or__ = __add__


class _Combined(AutoBody):
    """Union / difference / intersection of two bodies (src/AutoBody.jl:35-45,84-93)."""

    def __init__(self, a, b, op):
        self.a, self.b, self.op = a, b, op

    def _pick(self, da, db):
        if self.op == "+":
            return db < da, db
        if self.op == "-":
            return -db > da, -db
        return db > da, db

    def sdf(self, x, t=0.0, fastd2=math.inf):
        da, db = self.a.sdf(x, t), self.b.sdf(x, t)
        useb, dbs = self._pick(da, db)
        return torch.where(useb, dbs, da)

    def measure(self, x, t=0.0, fastd2=math.inf):
        da, na, Va = self.a.measure(x, t, fastd2)
        db, nb, Vb = self.b.measure(x, t, fastd2)
        if self.op == "-":  # a - b = a ∩ -b (src/AutoBody.jl:34): max(da, -db), with -b's flipped normal
            db, nb = -db, -nb
        useb = (db < da) if self.op == "+" else (db > da)
        return torch.where(useb, db, da), torch.where(useb[..., None], nb, na), torch.where(useb[..., None], Vb, Va)


# convolution kernel moments — reference src/Body.jl:56-61
def kern0(d):
    return 0.5 + 0.5 * d + 0.5 * np.sin(np.pi * d) / np.pi


def kern1(d):
    return 0.25 * (1 - d**2) - 0.5 * (d * np.sin(np.pi * d) + (1 + np.cos(np.pi * d)) / np.pi) / np.pi


def mu0_of(d, eps):
    return kern0(np.clip(d / eps, -1, 1))


def mu1_of(d, eps):
    return eps * kern1(np.clip(d / eps, -1, 1))


def loc_grid(i, Ng, dtype=np.float32):
    """``loc(i,I)`` for every cell of a padded grid (reference src/util.jl:160):
    ``I .- 1.5 .- 0.5δᵢ`` with 1-based I; ``i=0`` is the cell centre.  Returns ``(Ng..., D)``."""
    D = len(Ng)
    axes = [np.arange(1, n + 1, dtype=np.float64) - 1.5 for n in Ng]
    X = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1)
    if i > 0:
        X[..., i - 1] -= 0.5
    return X.astype(dtype)


class LocLazy:
    """``loc(i,I)`` for every cell without materialising the ``(Ng...,D)`` array: ``x[..., k]`` is a
    broadcastable 1-D coordinate, so ``uλ`` closures written with ``x[..., k]`` cost O(N) memory
    until their final broadcast (a 512³ initial condition would otherwise need several GB)."""

    def __init__(self, i, Ng, dtype=np.float32):
        self.shape = tuple(Ng) + (len(Ng),)
        self._ax = []
        for k, n in enumerate(Ng):
            a = np.arange(1, n + 1, dtype=np.float64) - 1.5 - (0.5 if (k == i - 1) else 0.0)
            shp = [1] * len(Ng)
            shp[k] = n
            self._ax.append(a.astype(dtype).reshape(shp))

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 2 and idx[0] is Ellipsis and isinstance(idx[1], int):
            return self._ax[idx[1]]
        return np.stack(np.broadcast_arrays(*self._ax), axis=-1)[idx]


def measure_fields(body, Ng, t=0.0, eps=1.0, dtype=np.float32):
    """The cell loop of ``measure!(flow,body;t,ϵ)`` — reference src/Body.jl:31-50.

    Returns ``(mu0, mu1, V, d)`` as padded Fortran-ordered arrays *before* the two ``BC!`` calls of
    src/Body.jl:51-52 (the caller applies those through the C ABI).
    """
    D = len(Ng)
    mu0 = np.ones(tuple(Ng) + (D,), dtype=dtype, order="F")
    mu1 = np.zeros(tuple(Ng) + (D, D), dtype=dtype, order="F")
    V = np.zeros(tuple(Ng) + (D,), dtype=dtype, order="F")
    dfield = np.zeros(tuple(Ng), dtype=dtype, order="F")
    if body is None or type(body) is NoBody:
        return mu0, mu1, V, dfield
    tdtype = torch.float32 if dtype == np.float32 else torch.float64
    d2lim = float((2 + eps) ** 2)
    inside = tuple(slice(1, n - 1) for n in Ng)
    xc = torch.from_numpy(np.ascontiguousarray(loc_grid(0, Ng, dtype)[inside])).to(tdtype)
    d = body.sdf(xc, t, fastd2=d2lim).detach().numpy().astype(dtype)
    dfield[inside] = d
    near = d.astype(np.float64) ** 2 < d2lim
    solid = (~near) & (d < 0)
    idx = np.nonzero(near)
    for i in range(1, D + 1):
        m0 = mu0[inside + (i - 1,)]
        m0[solid] = 0
        if idx[0].size:
            xi = torch.from_numpy(np.ascontiguousarray(loc_grid(i, Ng, dtype)[inside][idx])).to(tdtype)
            di, ni, Vi = body.measure(xi, t, fastd2=d2lim)
            di = di.detach().numpy().astype(np.float64)
            ni = ni.detach().numpy().astype(np.float64)
            Vi = Vi.detach().numpy().astype(np.float64)
            m0[idx] = mu0_of(di, eps).astype(dtype)
            Vc = V[inside + (i - 1,)]
            Vc[idx] = Vi[:, i - 1].astype(dtype)
            V[inside + (i - 1,)] = Vc
            k1 = mu1_of(di, eps)
            for j in range(1, D + 1):
                mc = mu1[inside + (i - 1, j - 1)]
                mc[idx] = (k1 * ni[:, j - 1]).astype(dtype)
                mu1[inside + (i - 1, j - 1)] = mc
        mu0[inside + (i - 1,)] = m0
    return mu0, mu1, V, dfield


def kern(d):
    """``kern(d) = 0.5+0.5cos(πd)`` — reference src/Body.jl:56."""
    return 0.5 + 0.5 * np.cos(np.pi * d)


def nds_field(body, Ng, t=0.0):
    """``nds(body, loc(0,I,Float64), t)`` for every cell of the padded grid — reference
    src/Metrics.jl:84-87: ``d,n,_ = measure(body,x,t,fastd²=1); n*kern(clamp(d,-1,1))``.
    Returns a Float64 Fortran-ordered array ``(Ng..., D)`` (zero for ``NoBody``)."""
    D = len(Ng)
    out = np.zeros(tuple(Ng) + (D,), dtype=np.float64, order="F")
    if body is None or type(body) is NoBody:
        return out
    x = torch.from_numpy(np.ascontiguousarray(loc_grid(0, Ng, np.float64)))
    d, n, _ = body.measure(x, t, fastd2=1.0)
    d = d.detach().numpy().astype(np.float64)
    n = n.detach().numpy().astype(np.float64)
    out[...] = n * kern(np.clip(d, -1, 1))[..., None]
    return out
