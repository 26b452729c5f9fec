// wl_poisson.cu — pressure solver kernels: Poisson (reference src/Poisson.jl) and
// MultiLevelPoisson (reference src/MultiLevelPoisson.jl), hand-written for sm_100a.
//
// Differences in *structure* from the reference (results are the same, see tests/):
//  * every reduction (`⋅`, `sum`) is fused into the kernel that produces its operands and is
//    finished on the device by the last block to arrive; the data-dependent decisions of pcg!
//    (|ρ|<10eps, |α|∉[1e-2,1e2]) are taken there and kept in device memory (LevelScalars), so
//    a whole V-cycle is enqueued without a single host synchronisation;
//  * kernels of an exited pcg! are predicated off by `sc->active`;
//  * L₂(p)=r⋅r is refreshed by whichever kernel last wrote r — no separate dot pass.
#include <algorithm>
#include <cstring>

#include "wl_internal.h"

namespace {

template <int D>
__device__ __forceinline__ float mult_at(const Lvl& l, int64_t I, const float* __restrict__ x) {
  // mult (Poisson.jl:69-75)
  float s = x[I] * l.Dg[I];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int64_t st = l.g.s[i];
    const float* Li = l.L + i * l.g.comp;
    s += (x[I - st] * Li[I] + x[I + st] * Li[I + st]);
  }
  return s;
}

// set_diag! (Poisson.jl:42-45) with diag (:48-54)
template <int D>
__global__ void __launch_bounds__(256) k_set_diag(Lvl l) {
  const Box b = box_interior(l.g);
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float* Li = l.L + i * l.g.comp;
      s -= (Li[I] + Li[I + l.g.s[i]]);
    }
    l.Dg[I] = s;
    l.iD[I] = (s * s < 2 * WL_EPS32) ? 0.f : 1.f / s;
  });
}

// restrictL! interior part (MultiLevelPoisson.jl:26-30) with restrictL (:10-16), up(I,i) (:1)
template <int D>
__global__ void __launch_bounds__(256) k_restrictL(Lvl c, Lvl f) {
  const Box b = box_interior(c.g);
  for_cells(c.g, b, [&](int i0, int j0, int k0, int64_t I) {
    const int cc[3] = {i0, j0, k0};
#pragma unroll
    for (int i = 0; i < D; ++i) {
      int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < D; ++k) { lo[k] = 2 * cc[k] - 1; hi[k] = 2 * cc[k] - (k == i ? 1 : 0); }
      const float* Lf = f.L + i * f.g.comp;
      float s = 0.f;
      for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
          for (int x = lo[0]; x <= hi[0]; ++x) s += Lf[x + y * f.g.s[1] + z * f.g.s[2]];
      c.L[I + i * c.g.comp] = 0.5f * s;
    }
  });
}

// residual! first loop + sum (Poisson.jl:93-94)
template <int D>
__global__ void __launch_bounds__(256) k_residual(Lvl l, ReduceWS ws) {
  const Box b = box_interior(l.g);
  double acc[2] = {0.0, 0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    const float r = l.iD[I] == 0.f ? 0.f : l.z[I] - mult_at<D>(l, I, l.x);
    l.r[I] = r;
    acc[0] += (double)r;
    acc[1] += (double)r * (double)r;
  });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) {
    const float s = (float)tot[0] / (float)l.g.interior;
    l.sc->rsum = s;
    l.sc->need_sub = fabsf(s) > 2 * WL_EPS32 ? 1 : 0;
    l.sc->r2 = (float)tot[1];
  }
}
// residual! second loop (Poisson.jl:96), predicated on |s| > 2eps
__global__ void __launch_bounds__(256) k_sub_mean(Lvl l, ReduceWS ws) {
  if (!l.sc->need_sub) return;
  const float s = l.sc->rsum;
  const Box b = box_interior(l.g);
  double acc[1] = {0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    const float r = l.r[I] - s;
    l.r[I] = r;
    acc[0] += (double)r * (double)r;
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) l.sc->r2 = (float)tot[0];
}

// mult! interior (Poisson.jl:66)
template <int D>
__global__ void __launch_bounds__(256) k_mult(Lvl l) {
  const Box b = box_interior(l.g);
  for_cells(l.g, b, [&](int, int, int, int64_t I) { l.z[I] = mult_at<D>(l, I, l.x); });
}

// Jacobi! (Poisson.jl:111): ϵ = r·iD
__global__ void __launch_bounds__(256) k_jacobi_eps(Lvl l) {
  const Box b = box_interior(l.g);
  for_cells(l.g, b, [&](int, int, int, int64_t I) { l.eps[I] = l.r[I] * l.iD[I]; });
}

// increment! (Poisson.jl:101-102) + L₂
template <int D>
__global__ void __launch_bounds__(256) k_increment(Lvl l, ReduceWS ws) {
  const Box b = box_interior(l.g);
  double acc[1] = {0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    const float r = l.r[I] - mult_at<D>(l, I, l.eps);
    l.r[I] = r;
    l.x[I] = l.x[I] + l.eps[I];
    acc[0] += (double)r * (double)r;
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) l.sc->r2 = (float)tot[0];
}

// restrict! (MultiLevelPoisson.jl:33) + fill!(coarse.x,0) (:75)
template <int D>
__global__ void __launch_bounds__(256) k_restrict(Lvl c, Lvl f) {
  const Box b = box_interior(c.g);
  for_cells(c.g, b, [&](int i0, int j0, int k0, int64_t I) {
    const int cc[3] = {i0, j0, k0};
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < D; ++k) { lo[k] = 2 * cc[k] - 1; hi[k] = 2 * cc[k]; }
    float s = 0.f;
    for (int z = lo[2]; z <= hi[2]; ++z)
      for (int y = lo[1]; y <= hi[1]; ++y)
        for (int x = lo[0]; x <= hi[0]; ++x) s += f.r[x + y * f.g.s[1] + z * f.g.s[2]];
    c.r[I] = s;
    c.x[I] = 0.f;
  });
}

// prolongate! (MultiLevelPoisson.jl:34) with down(I) (:2)
template <int D>
__global__ void __launch_bounds__(256) k_prolongate(Lvl f, Lvl c) {
  const Box b = box_interior(f.g);
  for_cells(f.g, b, [&](int i0, int j0, int k0, int64_t I) {
    const int64_t J = (int64_t)((i0 + 1) >> 1) + (D > 1 ? ((j0 + 1) >> 1) * c.g.s[1] : 0) + (D > 2 ? ((k0 + 1) >> 1) * c.g.s[2] : 0);
    f.eps[I] = c.x[J];
  });
}

// ---- pcg! (Poisson.jl:123-144) -----------------------------------------------------------
// :125-127  z = ϵ = r·iD ; ρ = r⋅z ; |ρ|<10eps → return      (z is rewritten by k_pcg_mult before
// any read, so only ϵ is stored)
__global__ void __launch_bounds__(256) k_pcg_init(Lvl l, ReduceWS ws) {
  const Box b = box_interior(l.g);
  double acc[1] = {0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    const float r = l.r[I];
    const float e = r * l.iD[I];
    l.eps[I] = e;
    acc[0] += (double)r * (double)e;
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) {
    const float rho = (float)tot[0];
    l.sc->rho = rho;
    l.sc->active = fabsf(rho) < 10 * WL_EPS32 ? 0 : 1;
  }
}
// :130-132  z = Aϵ ; α = ρ/(z⋅ϵ) ; |α|∉[1e-2,1e2] → return.  The dot runs over the padded arrays
// in the reference; the ghost-shell part (non-zero only at level 1) is in sc->ghost_dot.
template <int D>
__global__ void __launch_bounds__(256) k_pcg_mult(Lvl l, ReduceWS ws) {
  if (!l.sc->active) return;
  const Box b = box_interior(l.g);
  double acc[1] = {0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    const float z = mult_at<D>(l, I, l.eps);
    l.z[I] = z;
    acc[0] += (double)z * (double)l.eps[I];
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) {
    ++l.sc->cnt_mult;
    const float ze = (float)(tot[0] + l.sc->ghost_dot);
    const float alpha = l.sc->rho / ze;
    l.sc->alpha = alpha;
    const double aa = fabs((double)alpha);
    if (aa < 1e-2 || aa > 1e2) l.sc->active = 0;
  }
}
// :133-139  x += αϵ ; r -= αz ; [last sweep → return] ; z = r·iD ; ρ₂ = r⋅z ; |ρ₂|<10eps → return ;
// β = ρ₂/ρ ; ρ = ρ₂
__global__ void __launch_bounds__(256) k_pcg_update(Lvl l, ReduceWS ws, int last) {
  if (!l.sc->active) return;
  const float alpha = l.sc->alpha;
  const Box b = box_interior(l.g);
  double acc[2] = {0.0, 0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) {
    l.x[I] += alpha * l.eps[I];
    const float r = l.r[I] - alpha * l.z[I];
    l.r[I] = r;
    acc[0] += (double)r * (double)(r * l.iD[I]);
    acc[1] += (double)r * (double)r;
  });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) {
    l.sc->r2 = (float)tot[1];
    ++l.sc->cnt_update;
    if (last) {
      l.sc->active = 0;
    } else {
      const float rho2 = (float)tot[0];
      if (fabsf(rho2) < 10 * WL_EPS32) {
        l.sc->active = 0;
      } else {
        l.sc->beta = rho2 / l.sc->rho;
        l.sc->rho = rho2;
      }
    }
  }
}
// :140  ϵ = βϵ + z  with z = r·iD recomputed
__global__ void __launch_bounds__(256) k_pcg_dir(Lvl l) {
  if (!l.sc->active) return;
  const float beta = l.sc->beta;
  if (blockIdx.x == 0 && threadIdx.x + threadIdx.y + threadIdx.z == 0) atomicAdd(&l.sc->cnt_dir, 1);
  const Box b = box_interior(l.g);
  for_cells(l.g, b, [&](int, int, int, int64_t I) { l.eps[I] = beta * l.eps[I] + l.r[I] * l.iD[I]; });
}

// Σ_shell z·ϵ  → sc->ghost_dot  (level 1: z ≡ σ keeps the stale Φ of conv_diff! on its upper ghost
// layer, and ϵ's ghosts are non-zero after perBC! — reference quirk Q2, must be reproduced)
__global__ void __launch_bounds__(256) k_shell_dot(Lvl l, ReduceWS ws, int64_t count) {
  if (!l.sc->active) return;
  double acc[1] = {0.0};
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    int c[3];
    if (shell_decode(l.g, idx, c)) {
      const int64_t I = c[0] + c[1] * l.g.s[1] + c[2] * l.g.s[2];
      acc[0] += (double)l.z[I] * (double)l.eps[I];
    }
  }
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) l.sc->ghost_dot = tot[0];
}

__global__ void __launch_bounds__(256) k_dot_rr(Lvl l, ReduceWS ws, double* out) {
  const Box b = box_interior(l.g);
  double acc[1] = {0.0};
  for_cells(l.g, b, [&](int, int, int, int64_t I) { acc[0] += (double)l.r[I] * (double)l.r[I]; });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) *out = tot[0];
}

}  // namespace

#define LAUNCH(s, kern, ntiles, blk, ...)                                        \
  do {                                                                            \
    kern<<<wl_grid_for((s), (ntiles)), (blk), 0, (s).stream>>>(__VA_ARGS__);      \
    ++(s).launches;                                                               \
  } while (0)
#define LAUNCH_D(s, kern, D, ntiles, blk, ...)                                   \
  do {                                                                            \
    if ((D) == 2) LAUNCH(s, kern<2>, ntiles, blk, __VA_ARGS__);                   \
    else LAUNCH(s, kern<3>, ntiles, blk, __VA_ARGS__);                            \
  } while (0)

static inline int64_t interior_tiles(const GridD& g, dim3 blk) { return box_tiles(box_interior(g), blk); }
static inline double cellsB(const Lvl& l, int words) { return 4.0 * words * (double)l.g.interior; }

int pois_set_diag(Sim& s, Lvl& l) {
  const dim3 blk = wl_pick_block(l.g);
  LAUNCH_D(s, k_set_diag, s.D, interior_tiles(l.g, blk), blk, l);
  return wl_cuda_check(s, cudaGetLastError(), "k_set_diag");
}
int pois_restrictL(Sim& s, Lvl& c, const Lvl& f) {
  const dim3 blk = wl_pick_block(c.g);
  LAUNCH_D(s, k_restrictL, s.D, interior_tiles(c.g, blk), blk, c, f);
  const double zero[3] = {0, 0, 0};
  return bc_vec(s, c.g, c.L, zero, false, f.g.per);  // BC!(a,0,false,perdir) (MultiLevelPoisson.jl:31)
}
// update!(ml) (MultiLevelPoisson.jl:62-68)
int pois_update(Sim& s) {
  int rc = pois_set_diag(s, s.levels[0]);
  for (size_t l = 1; l < s.levels.size() && !rc; ++l) {
    rc = pois_restrictL(s, s.levels[l], s.levels[l - 1]);
    if (!rc) rc = pois_set_diag(s, s.levels[l]);
  }
  return rc;
}
// mult!(p,x) (Poisson.jl:62-68)
int pois_mult(Sim& s) {
  Lvl& l = s.levels[0];
  int rc = bc_per(s, l.g, l.x, l.g.per);
  if (rc) return rc;
  if ((rc = wl_fill(s, l.z - WL_OX, 0.f, l.g.comp))) return rc;
  const dim3 blk = wl_pick_block(l.g);
  LAUNCH_D(s, k_mult, s.D, interior_tiles(l.g, blk), blk, l);
  return wl_cuda_check(s, cudaGetLastError(), "k_mult");
}
// residual!(p) (Poisson.jl:91-97)
int pois_residual(Sim& s, Lvl& l) {
  int rc = bc_per(s, l.g, l.x, l.g.per);
  if (rc) return rc;
  const dim3 blk = wl_pick_block(l.g);
  const int64_t nt = interior_tiles(l.g, blk);
  const bool fine = &l == &s.levels[0];
  if (fine) wl_prof_begin(s, WL_PROF_RESIDUAL);
  LAUNCH_D(s, k_residual, s.D, nt, blk, l, s.ws);
  LAUNCH(s, k_sub_mean, nt, blk, l, s.ws);
  if (fine) wl_prof_end(s, WL_PROF_RESIDUAL, cellsB(l, 8));
  return wl_cuda_check(s, cudaGetLastError(), "k_residual");
}
// increment!(p) (Poisson.jl:99-103)
static int pois_increment(Sim& s, Lvl& l) {
  int rc = bc_per(s, l.g, l.eps, l.g.per);
  if (rc) return rc;
  const dim3 blk = wl_pick_block(l.g);
  const bool fine = &l == &s.levels[0];
  if (fine) wl_prof_begin(s, WL_PROF_INCREMENT);
  LAUNCH_D(s, k_increment, s.D, interior_tiles(l.g, blk), blk, l, s.ws);
  if (fine) wl_prof_end(s, WL_PROF_INCREMENT, cellsB(l, 9));
  return wl_cuda_check(s, cudaGetLastError(), "k_increment");
}
// Jacobi!(p;it) (Poisson.jl:110-113)
int pois_jacobi(Sim& s, Lvl& l, int it) {
  const dim3 blk = wl_pick_block(l.g);
  for (int t = 0; t < it; ++t) {
    LAUNCH(s, k_jacobi_eps, interior_tiles(l.g, blk), blk, l);
    const int rc = pois_increment(s, l);
    if (rc) return rc;
  }
  return 0;
}
// pcg!(p;it) (Poisson.jl:123-144) — enqueued in full; early exits are device-side predicates
int pois_pcg(Sim& s, int level, int it) {
  Lvl& l = s.levels[level];
  const dim3 blk = wl_pick_block(l.g);
  const int64_t nt = interior_tiles(l.g, blk);
  const bool fine = level == 0;
  bool anyper = false;
  for (int k = 0; k < s.D; ++k) anyper |= l.g.per[k] != 0;
  if (fine) wl_prof_begin(s, WL_PROF_PCG_INIT);
  LAUNCH(s, k_pcg_init, nt, blk, l, s.ws);
  if (fine) wl_prof_end(s, WL_PROF_PCG_INIT, cellsB(l, 3));
  for (int i = 1; i <= it; ++i) {
    if (anyper) {
      const int rc = bc_per(s, l.g, l.eps, l.g.per, &l.sc->active);
      if (rc) return rc;
      if (fine) {  // only level 1 has non-zero z ghosts (z ≡ σ)
        const int64_t cnt = shell_count(l.g);
        k_shell_dot<<<(int)std::min<int64_t>((cnt + 255) / 256, 256), 256, 0, s.stream>>>(l, s.ws, cnt);
        ++s.launches;
      }
    }
    if (fine) wl_prof_begin(s, WL_PROF_PCG_MULT);
    LAUNCH_D(s, k_pcg_mult, s.D, nt, blk, l, s.ws);
    if (fine) { wl_prof_end(s, WL_PROF_PCG_MULT, 0); wl_prof_begin(s, WL_PROF_PCG_UPDATE); }
    LAUNCH(s, k_pcg_update, nt, blk, l, s.ws, i == it ? 1 : 0);
    if (fine) wl_prof_end(s, WL_PROF_PCG_UPDATE, 0);
    if (i == it) break;
    if (fine) wl_prof_begin(s, WL_PROF_PCG_DIR);
    LAUNCH(s, k_pcg_dir, nt, blk, l);
    if (fine) wl_prof_end(s, WL_PROF_PCG_DIR, 0);
  }
  return wl_cuda_check(s, cudaGetLastError(), "pcg");
}
// Vcycle!(ml;l) (MultiLevelPoisson.jl:70-82)
int pois_vcycle(Sim& s, int level) {
  Lvl& fine = s.levels[level];
  Lvl& coarse = s.levels[level + 1];
  int rc = pois_jacobi(s, fine, 1);
  if (rc) return rc;
  {
    const dim3 blk = wl_pick_block(coarse.g);
    LAUNCH_D(s, k_restrict, s.D, interior_tiles(coarse.g, blk), blk, coarse, fine);
  }
  if (level + 2 < (int)s.levels.size() && (rc = pois_vcycle(s, level + 1))) return rc;
  if ((rc = pois_pcg(s, level + 1, 6))) return rc;
  {
    const dim3 blk = wl_pick_block(fine.g);
    LAUNCH_D(s, k_prolongate, s.D, interior_tiles(fine.g, blk), blk, fine, coarse);
  }
  return pois_increment(s, fine);
}
// solver!(ml;tol,itmx) (MultiLevelPoisson.jl:87-99) / solver!(p::Poisson;tol,itmx) (Poisson.jl:162-172)
int pois_solve(Sim& s, double tol, int itmx, int* n_out) {
  Lvl& p = s.levels[0];
  wl_prof_begin(s, WL_PROF_SOLVER);
  int rc = pois_residual(s, p);
  if (rc) return rc;
  int n = 0;
  while (n < itmx) {
    if (s.pois_kind == WL_POIS_MULTILEVEL && (rc = pois_vcycle(s, 0))) return rc;
    if ((rc = pois_pcg(s, 0, 6))) return rc;
    ++n;
    // r₂ = L₂(p): one host read per solver iteration (the reference needs ~100)
    if ((rc = wl_read_scalars(s, &p.sc->r2, sizeof(float)))) return rc;
    if ((double)s.h_pinned[0] < tol) break;
  }
  rc = bc_per(s, p.g, p.x, p.g.per);
  wl_prof_end(s, WL_PROF_SOLVER, 0.0);
  s.n_log.push_back(n);
  if (n_out) *n_out = n;
  return rc;
}
int pois_dot_rr(Sim& s, Lvl& l, double* out) {
  const dim3 blk = wl_pick_block(l.g);
  double* d = s.ws.partials + (size_t)WL_MAXBLOCKS * WL_MAXRED;  // spare slot behind the partials
  LAUNCH(s, k_dot_rr, interior_tiles(l.g, blk), blk, l, s.ws, d);
  int rc = wl_read_scalars(s, d, sizeof(double));
  if (rc) return rc;
  double v;
  memcpy(&v, s.h_pinned, sizeof(double));
  *out = v;
  return 0;
}
