"""ctypes binding of the C ABI declared in ``include/waterlily_b200.h``.

The same table binds ``libwaterlily_b200.so`` (prefix ``wl_``, the product) and — from the test
tree only — the CPU oracle (prefix ``wlo_``).  Nothing in this package ever loads the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
H = C.c_void_p

# name -> (restype, argtypes); mirrors include/waterlily_b200.h one to one
SIGNATURES = {
    "create": (C.c_int, [C.POINTER(H), C.c_int, c_int_p, C.c_double, C.c_double, c_int_p, C.c_int, C.c_int, C.c_int]),
    "dist_unique_id": (C.c_int, [C.c_char_p]),
    "create_dist": (C.c_int, [C.POINTER(H), C.c_int, c_int_p, C.c_double, C.c_double, c_int_p, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_char_p]),
    "dist_info": (C.c_int, [H, c_int_p, c_int_p, c_int_p, c_int_p, c_i64_p]),
    "dist_finalize": (C.c_int, []),
    "dist_plan": (C.c_int, [C.c_int, c_int_p, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "destroy": (C.c_int, [H]),
    "last_error": (C.c_char_p, []),
    "set_stream": (C.c_int, [H, C.c_void_p]),
    "sync": (C.c_int, [H]),
    "field_count": (C.c_int, [H, C.c_int, c_i64_p]),
    "upload": (C.c_int, [H, C.c_int, c_flt_p, C.c_int64]),
    "download": (C.c_int, [H, C.c_int, c_flt_p, C.c_int64]),
    "nlevels": (C.c_int, [H]),
    "level_dims": (C.c_int, [H, C.c_int, c_int_p]),
    "level_upload": (C.c_int, [H, C.c_int, C.c_int, c_flt_p, C.c_int64]),
    "level_download": (C.c_int, [H, C.c_int, C.c_int, c_flt_p, C.c_int64]),
    "bc": (C.c_int, [H, C.c_int, c_dbl_p, C.c_int, c_int_p]),
    "exit_bc": (C.c_int, [H, c_dbl_p, C.c_double]),
    "per_bc": (C.c_int, [H, C.c_int, c_int_p]),
    "conv_diff": (C.c_int, [H, C.c_int]),
    "accelerate": (C.c_int, [H, c_dbl_p]),
    "bdim": (C.c_int, [H]),
    "scale_u": (C.c_int, [H, C.c_double]),
    "project": (C.c_int, [H, C.c_double, c_int_p]),
    "cfl": (C.c_int, [H, c_dbl_p]),
    "pois_update": (C.c_int, [H]),
    "pois_mult": (C.c_int, [H]),
    "pois_residual": (C.c_int, [H, C.c_int, c_dbl_p]),
    "pois_jacobi": (C.c_int, [H, C.c_int, C.c_int]),
    "pois_pcg": (C.c_int, [H, C.c_int, C.c_int]),
    "pois_vcycle": (C.c_int, [H, C.c_int]),
    "pois_solve": (C.c_int, [H, C.c_double, C.c_int, c_int_p]),
    "pressure_force": (C.c_int, [H, c_dbl_p, C.c_int64, c_dbl_p]),
    "viscous_force": (C.c_int, [H, c_dbl_p, C.c_int64, c_dbl_p]),
    "pressure_moment": (C.c_int, [H, c_dbl_p, c_dbl_p, C.c_int64, c_dbl_p]),
    "derived": (C.c_int, [H, C.c_int, C.c_int, c_dbl_p, c_dbl_p, c_dbl_p]),
    "mom_step": (C.c_int, [H, c_dbl_p, c_dbl_p, c_dbl_p, c_dbl_p, c_int_p, c_int_p]),
    "pois_head": (C.c_int, [H]),
    "get_dt": (C.c_int, [H, c_dbl_p]),
    "set_dt": (C.c_int, [H, C.c_double]),
    "launch_count": (C.c_int64, [H]),
    "profile": (C.c_int, [H, C.c_int]),
    "profile_count": (C.c_int, []),
    "profile_name": (C.c_char_p, [C.c_int]),
    "profile_read": (C.c_int, [H, C.c_int, c_dbl_p, c_i64_p, c_dbl_p]),
}

# enum wl_field / wl_level_field / wl_pois_kind
U, U0, F, P, SIGMA, V, MU0, MU1 = range(8)
LVL_L, LVL_D, LVL_ID, LVL_X, LVL_EPS, LVL_R, LVL_Z = range(7)
POIS_MULTILEVEL, POIS_SINGLE = 0, 1
FIELD_IDS = {"u": U, "u0": U0, "f": F, "p": P, "sigma": SIGMA, "V": V, "mu0": MU0, "mu1": MU1}
LEVEL_IDS = {"L": LVL_L, "D": LVL_D, "iD": LVL_ID, "x": LVL_X, "eps": LVL_EPS, "r": LVL_R, "z": LVL_Z}


class WLError(RuntimeError):
    """A non-zero status from the C ABI (the reference raises AssertionError / ErrorException)."""


class CABI:
    """Loads one shared library and binds every entry point of the header under ``prefix``."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "wl_" else C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, prefix + name)  # AttributeError = a declared symbol is missing
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def check(self, status: int):
        if status != 0:
            msg = self.last_error()
            raise WLError((msg.decode("utf-8") if msg else "") + f" (status {status})")


def nccl_unique_id(cabi: CABI) -> bytes:
    """128-byte NCCL id made by rank 0 and broadcast by the launcher (wl_dist_unique_id)."""
    buf = C.create_string_buffer(128)
    cabi.check(cabi.dist_unique_id(buf))
    return buf.raw


def dist_plan(cabi: CABI, dims, perdir, rank, nranks, pois_kind=0):
    """Slab plan of one rank (host-only, wl_dist_plan): list of per-level dicts."""
    D = len(dims)
    per = [1 if (j + 1) in tuple(perdir) else 0 for j in range(D)]
    nl = C.c_int(0)
    arrs = [(C.c_int * 16)() for _ in range(5)]
    cabi.check(cabi.dist_plan(D, _ints(dims), _ints(per), pois_kind, rank, nranks, 16, C.byref(nl), *arrs))
    keys = ("distributed", "nz_local", "zoff", "zlo_halo", "zhi_halo")
    return [{k: int(a[l]) for k, a in zip(keys, arrs)} for l in range(nl.value)]


def _ints(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def _dbls(vals):
    return (C.c_double * len(vals))(*[float(v) for v in vals])


class Engine:
    """One ``wl_sim`` handle: owns the device (or oracle) state of a Flow + its Poisson solver.

    Arrays cross the boundary in the reference's layout — padded, column-major, components
    last — represented in numpy as Fortran-ordered float32 arrays of shape ``(Ng..., D)``.
    """

    def __init__(self, cabi: CABI, dims, nu=0.0, dt0=0.25, perdir=(), exitBC=False, pois_kind=POIS_MULTILEVEL, device=0, dist=None):
        """``dist=(rank, nranks, nccl_id_bytes)`` makes this handle one z-slab of a multi-GPU job."""
        self.cabi = cabi
        self.D = len(dims)
        self.N = tuple(int(n) for n in dims)
        self.Ng = tuple(n + 2 for n in self.N)
        self.perdir = tuple(sorted(int(j) for j in perdir))  # 1-based, like the reference
        self.per_flags = [1 if (j + 1) in self.perdir else 0 for j in range(self.D)]
        self.exitBC = bool(exitBC)
        self.h = H()
        self.dist = dist
        if dist is None or dist[1] == 1:
            cabi.check(cabi.create(C.byref(self.h), self.D, _ints(self.N), float(nu), float(dt0), _ints(self.per_flags),
                                   int(self.exitBC), int(pois_kind), int(device)))
        else:
            rank, nranks, ident = dist
            cabi.check(cabi.create_dist(C.byref(self.h), self.D, _ints(self.N), float(nu), float(dt0), _ints(self.per_flags),
                                        int(self.exitBC), int(pois_kind), int(device), int(rank), int(nranks), bytes(ident)))

    # -- lifetime
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.cabi.destroy(self.h)
            self.h = H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- shapes
    def shape(self, field: int):
        if field in (P, SIGMA):
            return self.Ng
        if field == MU1:
            return self.Ng + (self.D, self.D)
        return self.Ng + (self.D,)

    # -- transfer
    def upload(self, field, arr):
        field = FIELD_IDS.get(field, field)
        a = np.asfortranarray(arr, dtype=np.float32)
        if a.shape != self.shape(field):
            raise ValueError(f"field {field}: expected shape {self.shape(field)}, got {a.shape}")
        self.cabi.check(self.cabi.upload(self.h, field, a.ctypes.data_as(c_flt_p), a.size))

    def download(self, field):
        field = FIELD_IDS.get(field, field)
        a = np.zeros(self.shape(field), dtype=np.float32, order="F")  # a slab handle fills only the planes it owns
        self.cabi.check(self.cabi.download(self.h, field, a.ctypes.data_as(c_flt_p), a.size))
        return a

    def dist_info(self):
        r, n, z, nz, c = C.c_int(0), C.c_int(1), C.c_int(0), C.c_int(0), C.c_int64(0)
        self.cabi.check(self.cabi.dist_info(self.h, C.byref(r), C.byref(n), C.byref(z), C.byref(nz), C.byref(c)))
        return {"rank": r.value, "nranks": n.value, "zoff": z.value, "nz_local": nz.value, "collectives": c.value}

    def nlevels(self):
        return int(self.cabi.nlevels(self.h))

    def level_dims(self, level):
        out = (C.c_int * self.D)()
        self.cabi.check(self.cabi.level_dims(self.h, level, out))
        return tuple(out)

    def level_shape(self, level, lfield):
        d = self.level_dims(level)
        return d + (self.D,) if lfield == LVL_L else d

    def level_upload(self, level, lfield, arr):
        lfield = LEVEL_IDS.get(lfield, lfield)
        a = np.asfortranarray(arr, dtype=np.float32)
        if a.shape != self.level_shape(level, lfield):
            raise ValueError(f"level {level} field {lfield}: expected {self.level_shape(level, lfield)}, got {a.shape}")
        self.cabi.check(self.cabi.level_upload(self.h, level, lfield, a.ctypes.data_as(c_flt_p), a.size))

    def level_download(self, level, lfield):
        lfield = LEVEL_IDS.get(lfield, lfield)
        a = np.zeros(self.level_shape(level, lfield), dtype=np.float32, order="F")
        self.cabi.check(self.cabi.level_download(self.h, level, lfield, a.ctypes.data_as(c_flt_p), a.size))
        return a

    # -- BCs
    def _per(self, perdir):
        if perdir is None:
            return None
        return _ints([1 if (j + 1) in tuple(perdir) else 0 for j in range(self.D)])

    def bc(self, field, A, saveexit=False, perdir=None):
        self.cabi.check(self.cabi.bc(self.h, FIELD_IDS.get(field, field), _dbls(A), int(saveexit), self._per(perdir)))

    def exit_bc(self, U, dt):
        self.cabi.check(self.cabi.exit_bc(self.h, _dbls(U), float(dt)))

    def per_bc(self, field, perdir=None):
        self.cabi.check(self.cabi.per_bc(self.h, FIELD_IDS.get(field, field), self._per(perdir)))

    # -- momentum pieces
    def conv_diff(self, src="u"):
        self.cabi.check(self.cabi.conv_diff(self.h, FIELD_IDS.get(src, src)))

    def accelerate(self, g):
        self.cabi.check(self.cabi.accelerate(self.h, None if g is None else _dbls(g)))

    def bdim(self):
        self.cabi.check(self.cabi.bdim(self.h))

    def scale_u(self, scale):
        self.cabi.check(self.cabi.scale_u(self.h, float(scale)))

    def project(self, w=1.0):
        n = C.c_int(0)
        self.cabi.check(self.cabi.project(self.h, float(w), C.byref(n)))
        return n.value

    def cfl(self):
        dt = C.c_double(0)
        self.cabi.check(self.cabi.cfl(self.h, C.byref(dt)))
        return dt.value

    # -- Metrics (src/Metrics.jl:94-141); nds: Float64 (Ng...,D) from the host geometry, None = reuse
    def _nds_args(self, nds):
        if nds is None:
            return None, 0
        a = np.asfortranarray(nds, dtype=np.float64)
        if a.shape != tuple(self.Ng) + (self.D,):
            raise ValueError(f"nds: expected shape {tuple(self.Ng) + (self.D,)}, got {a.shape}")
        self._nds_keep = a
        return a.ctypes.data_as(c_dbl_p), a.size

    def pressure_force(self, nds=None):
        ptr, cnt = self._nds_args(nds)
        out = (C.c_double * self.D)()
        self.cabi.check(self.cabi.pressure_force(self.h, ptr, cnt, out))
        return np.array(out[:], dtype=np.float64)

    def viscous_force(self, nds=None):
        ptr, cnt = self._nds_args(nds)
        out = (C.c_double * self.D)()
        self.cabi.check(self.cabi.viscous_force(self.h, ptr, cnt, out))
        return np.array(out[:], dtype=np.float64)

    def pressure_moment(self, x0, nds=None):
        ptr, cnt = self._nds_args(nds)
        out = (C.c_double * self.D)()
        self.cabi.check(self.cabi.pressure_moment(self.h, _dbls(x0), ptr, cnt, out))
        return np.array(out[:], dtype=np.float64)

    def derived(self, kind, comp=0, U=None, axis=None, center=None):
        """``@inside σ[I] = f(I,u)`` for a Metrics.jl diagnostic (wl_derived); returns σ."""
        kinds = {"ke": 0, "curl": 1, "omega_mag": 2, "lambda2": 3, "omega_theta": 4}
        k = kinds.get(kind, kind)
        self.cabi.check(self.cabi.derived(self.h, int(k), int(comp), None if U is None else _dbls(U),
                                          None if axis is None else _dbls(axis), None if center is None else _dbls(center)))
        return self.download(SIGMA)

    # -- Poisson
    def pois_update(self):
        self.cabi.check(self.cabi.pois_update(self.h))

    def pois_mult(self):
        self.cabi.check(self.cabi.pois_mult(self.h))

    def pois_residual(self, level=0):
        r2 = C.c_double(0)
        self.cabi.check(self.cabi.pois_residual(self.h, level, C.byref(r2)))
        return r2.value

    def pois_jacobi(self, level=0, it=1):
        self.cabi.check(self.cabi.pois_jacobi(self.h, level, it))

    def pois_pcg(self, level=0, it=6):
        self.cabi.check(self.cabi.pois_pcg(self.h, level, it))

    def pois_vcycle(self, level=0):
        self.cabi.check(self.cabi.pois_vcycle(self.h, level))

    def pois_head(self):
        """residual! + Jacobi! + restrict! + fill!(coarse.x,0): the head of solver!(ml) on the finest level."""
        self.cabi.check(self.cabi.pois_head(self.h))

    def pois_solve(self, tol=-1.0, itmx=-1):
        n = C.c_int(0)
        self.cabi.check(self.cabi.pois_solve(self.h, float(tol), int(itmx), C.byref(n)))
        return n.value

    # -- hot path
    def mom_step(self, U, g_pred=None, g_corr=None):
        dt = C.c_double(0)
        n1, n2 = C.c_int(0), C.c_int(0)
        self.cabi.check(self.cabi.mom_step(self.h, _dbls(U), None if g_pred is None else _dbls(g_pred),
                                           None if g_corr is None else _dbls(g_corr), C.byref(dt), C.byref(n1), C.byref(n2)))
        return dt.value, n1.value, n2.value

    def get_dt(self):
        dt = C.c_double(0)
        self.cabi.check(self.cabi.get_dt(self.h, C.byref(dt)))
        return dt.value

    def set_dt(self, dt):
        self.cabi.check(self.cabi.set_dt(self.h, float(dt)))

    def set_stream(self, stream_ptr):
        self.cabi.check(self.cabi.set_stream(self.h, C.c_void_p(stream_ptr)))

    def sync(self):
        self.cabi.check(self.cabi.sync(self.h))

    def launch_count(self):
        return int(self.cabi.launch_count(self.h))

    def profile(self, on=True):
        self.cabi.check(self.cabi.profile(self.h, int(on)))

    def profile_classes(self):
        return [self.cabi.profile_name(i).decode() for i in range(self.cabi.profile_count())]

    def profile_read(self, which):
        ms, n, b = C.c_double(0), C.c_int64(0), C.c_double(0)
        self.cabi.check(self.cabi.profile_read(self.h, which, C.byref(ms), C.byref(n), C.byref(b)))
        return ms.value, n.value, b.value
