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
#include <cstdlib>
#include <cstring>

#include "wl_internal.h"
#include "wl_vec.cuh"

namespace {

// mult (Poisson.jl:69-75) for V consecutive cells starting at I; xc = x[I..I+V)
template <int D, int V>
__device__ __forceinline__ Pack<V> mult_row(const Lvl& l, int64_t I, const float* __restrict__ x, const Pack<V>& xc) {
  const Pack<V> Dc = ldp<V>(l.Dg + I);
  Pack<V> s;
#pragma unroll
  for (int q = 0; q < V; ++q) s.v[q] = xc.v[q] * Dc.v[q];
  {
    const Pack<V> Lc = ldp<V>(l.L + I), Lp = ld_xp<V>(l.L + I, Lc), xm = ld_xm<V>(x + I, xc), xp = ld_xp<V>(x + I, xc);
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
#pragma unroll
  for (int i = 1; i < D; ++i) {
    const int64_t st = l.g.s[i];
    const float* __restrict__ Li = l.L + i * l.g.comp;
    const Pack<V> Lc = ldp<V>(Li + I), Lp = ldp<V>(Li + I + st), xm = ldp<V>(x + I - st), xp = ldp<V>(x + I + st);
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
  return s;
}

// set_diag! (Poisson.jl:42-45) with diag (:48-54)
// Also classifies every aligned pack of four cells: umask = 0 when all its 2D·4 face coefficients equal
// uL (then D = uD and iD = uiD exactly, see Lvl) — `umask` is the array to fill (or null), `nonuni`
// counts the other packs (profiling: algorithmic bytes depend on it).
template <int D, int V>
__global__ void __launch_bounds__(256) k_set_diag(Lvl l, unsigned char* __restrict__ umask, unsigned int* __restrict__ nonuni) {
  unsigned int mine = 0;
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    Pack<V> s;
    bool uni = true;
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] = 0.f;
    {
      const Pack<V> Lc = ldp<V>(l.L + I), Lp = ld_xp<V>(l.L + I, Lc);
#pragma unroll
      for (int q = 0; q < V; ++q) { s.v[q] -= (Lc.v[q] + Lp.v[q]); uni &= (Lc.v[q] == l.uL) & (Lp.v[q] == l.uL); }
    }
#pragma unroll
    for (int i = 1; i < D; ++i) {
      const float* __restrict__ Li = l.L + i * l.g.comp;
      const Pack<V> Lc = ldp<V>(Li + I), Lp = ldp<V>(Li + I + l.g.s[i]);
#pragma unroll
      for (int q = 0; q < V; ++q) { s.v[q] -= (Lc.v[q] + Lp.v[q]); uni &= (Lc.v[q] == l.uL) & (Lp.v[q] == l.uL); }
    }
    Pack<V> inv;
#pragma unroll
    for (int q = 0; q < V; ++q) inv.v[q] = (s.v[q] * s.v[q] < 2 * WL_EPS32) ? 0.f : 1.f / s.v[q];
    stp<V>(l.Dg + I, s);
    stp<V>(l.iD + I, inv);
    if (V == 4 && umask) {
      umask[(I + l.g.lead) >> 2] = uni ? 0 : 1;
      mine += uni ? 0u : 1u;
    }
  });
  if (V == 4 && nonuni) {
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (((threadIdx.x + blockDim.x * threadIdx.y) & 31) == 0 && mine) atomicAdd(nonuni, mine);
  }
}

// restrictL! interior part (MultiLevelPoisson.jl:26-30) with restrictL (:10-16), up(I,i) (:1)
template <int D, int V>
__global__ void __launch_bounds__(256) k_restrictL(Lvl c, Lvl f, int ck0, int ck1, int fshift) {
  // ck0..ck1: coarse planes this rank computes; fshift: fine local k = fine global k + fshift
  // (both trivial unless the fine level is z-distributed and the coarse one replicated)
  for_rows<1>(c.g, [&](int i0, int j0, int k0, int64_t I) {
    if (D > 2 && (k0 < ck0 || k0 > ck1)) return;
    const int cc[3] = {i0, j0, k0};
#pragma unroll
    for (int i = 0; i < D; ++i) {
      int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < D; ++k) { lo[k] = 2 * cc[k] - 1; hi[k] = 2 * cc[k] - (k == i ? 1 : 0); }
      if (D > 2) { lo[2] += fshift; hi[2] += fshift; }
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
// `divu` non-null: z is not read but evaluated on the spot as div(u) (project! inside mom_step!: the separate z = div u
// pass and its 4 words per cell disappear; z's interior is rewritten by pcg! before anything reads it)
template <int D, int V>
__global__ void __launch_bounds__(256) k_residual(Lvl l, ReduceWS ws, FinArgs fa, int write_eps, const float* __restrict__ divu) {
  double acc[2] = {0.0, 0.0};
  for_rows<V>(l.g, [&](int i, int j, int k, int64_t I) {
    const bool uni = pack_uniform<V>(l, I);
    const Pack<V> xc = ldp<V>(l.x + I), ax = mult_w<D, V>(l, l.x, i, j, k, I, xc, uni), zc = divu ? div_pack<D, V>(l.g, divu, I) : ldp<V>(l.z + I),
                  id = ld_iD<V>(l, I, uni);
    Pack<V> r;
#pragma unroll
    for (int q = 0; q < V; ++q) {
      r.v[q] = id.v[q] == 0.f ? 0.f : zc.v[q] - ax.v[q];
      acc[0] += (double)r.v[q];
      acc[1] = dfma(r.v[q], r.v[q], acc[1]);
    }
    stp<V>(l.r + I, r);
    if (write_eps) {  // ϵ = r·iD of the Jacobi! that opens the V-cycle (Poisson.jl:111): saves its own pass over r, iD
      Pack<V> e;
#pragma unroll
      for (int q = 0; q < V; ++q) e.v[q] = r.v[q] * id.v[q];
      stp<V>(l.eps + I, e);
    }
  });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(FIN_RESIDUAL, fa, tot[0], tot[1]);
}
// residual! second loop (Poisson.jl:96), predicated on |s| > 2eps
template <int D, int V>
__global__ void __launch_bounds__(256) k_sub_mean(Lvl l, ReduceWS ws, FinArgs fa, int write_eps) {
  if (!l.sc->need_sub) return;
  const float s = l.sc->rsum;
  double acc[1] = {0.0};
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    Pack<V> r = ldp<V>(l.r + I);
#pragma unroll
    for (int q = 0; q < V; ++q) { r.v[q] = r.v[q] - s; acc[0] = dfma(r.v[q], r.v[q], acc[0]); }
    stp<V>(l.r + I, r);
    if (write_eps) {
      const Pack<V> id = ld_iD<V>(l, I, pack_uniform<V>(l, I));
      Pack<V> e;
#pragma unroll
      for (int q = 0; q < V; ++q) e.v[q] = r.v[q] * id.v[q];
      stp<V>(l.eps + I, e);
    }
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_R2_IF_SUB, fa, tot[0], 0.0);
}

// mult! interior (Poisson.jl:66)
template <int D, int V>
__global__ void __launch_bounds__(256) k_mult(Lvl l) {
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    const Pack<V> xc = ldp<V>(l.x + I);
    stp<V>(l.z + I, mult_row<D, V>(l, I, l.x, xc));
  });
}

// Jacobi! (Poisson.jl:111): ϵ = r·iD
template <int D, int V>
__global__ void __launch_bounds__(256) k_jacobi_eps(Lvl l) {
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    const Pack<V> r = ldp<V>(l.r + I), id = ld_iD<V>(l, I, pack_uniform<V>(l, I));
    Pack<V> e;
#pragma unroll
    for (int q = 0; q < V; ++q) e.v[q] = r.v[q] * id.v[q];
    stp<V>(l.eps + I, e);
  });
}

// increment! (Poisson.jl:101-102) + L₂
template <int D, int V>
__global__ void __launch_bounds__(256) k_increment(Lvl l, ReduceWS ws, FinArgs fa) {
  double acc[1] = {0.0};
  for_rows<V>(l.g, [&](int i, int j, int k, int64_t I) {
    const Pack<V> ec = ldp<V>(l.eps + I), ae = mult_w<D, V>(l, l.eps, i, j, k, I, ec, pack_uniform<V>(l, I));
    Pack<V> r = ldp<V>(l.r + I), x = ldp<V>(l.x + I);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      r.v[q] = r.v[q] - ae.v[q];
      x.v[q] = x.v[q] + ec.v[q];
      acc[0] = dfma(r.v[q], r.v[q], acc[0]);
    }
    stp<V>(l.r + I, r);
    stp<V>(l.x + I, x);
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_R2, fa, tot[0], 0.0);
}

// restrict! (MultiLevelPoisson.jl:33) + fill!(coarse.x,0) (:75)
template <int D, int V>
__global__ void __launch_bounds__(256) k_restrict(Lvl c, Lvl f, int ck0, int ck1, int fshift) {
  for_rows<1>(c.g, [&](int i0, int j0, int k0, int64_t I) {
    c.x[I] = 0.f;
    if (D > 2 && (k0 < ck0 || k0 > ck1)) return;
    const int cc[3] = {i0, j0, k0};
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < D; ++k) { lo[k] = 2 * cc[k] - 1; hi[k] = 2 * cc[k]; }
    if (D > 2) { lo[2] += fshift; hi[2] += fshift; }
    float s = 0.f;
    for (int z = lo[2]; z <= hi[2]; ++z)
      for (int y = lo[1]; y <= hi[1]; ++y)
        for (int x = lo[0]; x <= hi[0]; ++x) s += f.r[x + y * f.g.s[1] + z * f.g.s[2]];
    c.r[I] = s;
  });
}

// prolongate! (MultiLevelPoisson.jl:34) with down(I) (:2)
template <int D, int V>
__global__ void __launch_bounds__(256) k_prolongate(Lvl f, Lvl c, int fzadd) {
  // fzadd: fine global k = fine local k + fzadd when the coarse level is replicated (else 0)
  for_rows<V>(f.g, [&](int i0, int j0, int k0, int64_t I) {
    const int64_t J = (D > 1 ? ((j0 + 1) >> 1) * c.g.s[1] : 0) + (D > 2 ? ((k0 + fzadd + 1) >> 1) * c.g.s[2] : 0);
    Pack<V> e;
#pragma unroll
    for (int q = 0; q < V; ++q) e.v[q] = c.x[J + ((i0 + q + 1) >> 1)];
    stp<V>(f.eps + I, e);
  });
}

// prolongate! + increment! + ϵ = r·iD, ρ = r⋅ϵ of the pcg! that follows (see wl_vec.cuh)
template <int D, int V>
__global__ void __launch_bounds__(256) k_prolong_increment_init(Lvl f, Lvl c, ProMap m, const float* __restrict__ xsrc, ReduceWS ws, FinArgs fa) {
  double acc[2] = {0.0, 0.0};
  for_rows<V>(f.g, [&](int i, int j, int k, int64_t I) { prolong_increment_init<D, V>(f, c, m, xsrc, i, j, k, I, acc[0], acc[1]); });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(FIN_INC_INIT, fa, tot[0], tot[1]);
}

// ---- z-marching variants of the stencil kernels (3-D, V = 4; see for_columns in wl_vec.cuh) ---------
// Same arithmetic as k_residual / k_increment / k_pcg_mult / k_prolong_increment_init, cell for cell.
__global__ void __launch_bounds__(256) k_residual_m(Lvl l, MarchGeom mg, ReduceWS ws, FinArgs fa, int write_eps, const float* __restrict__ divu) {
  double acc[2] = {0.0, 0.0};
  const GridD& g = l.g;
  for_columns(g, mg, [&](int i, int j, int k0, int k1) {
    int64_t I = i + j * g.s[1] + (int64_t)k0 * g.s[2];
    Pack<4> xm = ldp<4>(l.x + I + zoff_m(g, k0)), xc = ldp<4>(l.x + I);
    for (int k = k0; k <= k1; ++k, I += g.s[2]) {
      const Pack<4> xp = ldp<4>(l.x + I + zoff_p(g, k));
      const bool uni = pack_uniform<4>(l, I);
      const Pack<4> ax = mult_zreg<4>(l, l.x, i, j, I, xc, xm, xp, uni), zc = divu ? div_pack<3, 4>(g, divu, I) : ldp<4>(l.z + I), id = ld_iD<4>(l, I, uni);
      Pack<4> r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        r.v[q] = id.v[q] == 0.f ? 0.f : zc.v[q] - ax.v[q];
        acc[0] += (double)r.v[q];
        acc[1] = dfma(r.v[q], r.v[q], acc[1]);
      }
      stp<4>(l.r + I, r);
      if (write_eps) {
        Pack<4> e;
#pragma unroll
        for (int q = 0; q < 4; ++q) e.v[q] = r.v[q] * id.v[q];
        stp<4>(l.eps + I, e);
      }
      xm = xc; xc = xp;
    }
  });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(FIN_RESIDUAL, fa, tot[0], tot[1]);
}
__global__ void __launch_bounds__(256) k_increment_m(Lvl l, MarchGeom mg, ReduceWS ws, FinArgs fa) {
  double acc[1] = {0.0};
  const GridD& g = l.g;
  for_columns(g, mg, [&](int i, int j, int k0, int k1) {
    int64_t I = i + j * g.s[1] + (int64_t)k0 * g.s[2];
    Pack<4> em = ldp<4>(l.eps + I + zoff_m(g, k0)), ec = ldp<4>(l.eps + I);
    for (int k = k0; k <= k1; ++k, I += g.s[2]) {
      const Pack<4> ep = ldp<4>(l.eps + I + zoff_p(g, k));
      const Pack<4> ae = mult_zreg<4>(l, l.eps, i, j, I, ec, em, ep, pack_uniform<4>(l, I));
      Pack<4> r = ldp<4>(l.r + I), x = ldp<4>(l.x + I);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        r.v[q] = r.v[q] - ae.v[q];
        x.v[q] = x.v[q] + ec.v[q];
        acc[0] = dfma(r.v[q], r.v[q], acc[0]);
      }
      stp<4>(l.r + I, r);
      stp<4>(l.x + I, x);
      em = ec; ec = ep;
    }
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_R2, fa, tot[0], 0.0);
}
__global__ void __launch_bounds__(256) k_pcg_mult_m(Lvl l, MarchGeom mg, ReduceWS ws, FinArgs fa) {
  if (!l.sc->active) return;
  double acc[1] = {0.0};
  const GridD& g = l.g;
  for_columns(g, mg, [&](int i, int j, int k0, int k1) {
    int64_t I = i + j * g.s[1] + (int64_t)k0 * g.s[2];
    Pack<4> em = ldp<4>(l.eps + I + zoff_m(g, k0)), ec = ldp<4>(l.eps + I);
    for (int k = k0; k <= k1; ++k, I += g.s[2]) {
      const Pack<4> ep = ldp<4>(l.eps + I + zoff_p(g, k));
      const Pack<4> z = mult_zreg<4>(l, l.eps, i, j, I, ec, em, ep, pack_uniform<4>(l, I));
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[0] = dfma(z.v[q], ec.v[q], acc[0]);
      stp<4>(l.z + I, z);
      em = ec; ec = ep;
    }
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_PCG_MULT, fa, tot[0], 0.0);
}
// ---- the same with the stencil operand staged through shared memory (march_staged / tile_* in wl_vec.cuh) ------
#ifndef WL_STAGE_MINB
#define WL_STAGE_MINB 3  // 80 registers, no spills: three resident blocks of the staged kernels per SM (measured: 2 and 4 are slower)
#endif
extern __shared__ __align__(16) float wl_ring[];
__global__ void __launch_bounds__(256, WL_STAGE_MINB) k_pcg_mult_s(Lvl l, MarchGeom mg, ReduceWS ws, FinArgs fa) {
  if (!l.sc->active) return;
  double acc[1] = {0.0};
  for_columns(l.g, mg, [&](int i, int j, int k0, int k1) { tile_pcg_mult(l, mg, wl_ring, i, j, k0, k1, acc[0]); });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_PCG_MULT, fa, tot[0], 0.0);
}
__global__ void __launch_bounds__(256, WL_STAGE_MINB) k_increment_s(Lvl l, MarchGeom mg, ReduceWS ws, FinArgs fa) {
  double acc[1] = {0.0};
  for_columns(l.g, mg, [&](int i, int j, int k0, int k1) { tile_increment(l, mg, wl_ring, i, j, k0, k1, acc[0]); });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_R2, fa, tot[0], 0.0);
}
__global__ void __launch_bounds__(256, WL_STAGE_MINB) k_residual_s(Lvl l, MarchGeom mg, ReduceWS ws, FinArgs fa, int write_eps, const float* __restrict__ divu) {
  double acc[2] = {0.0, 0.0};
  for_columns(l.g, mg, [&](int i, int j, int k0, int k1) { tile_residual(l, mg, wl_ring, i, j, k0, k1, write_eps, divu, acc[0], acc[1]); });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(FIN_RESIDUAL, fa, tot[0], tot[1]);
}
#ifndef WL_PROLONG_MINB
#define WL_PROLONG_MINB 5  // 48 instead of 64 registers, five resident blocks: the pass is latency-bound on its L1-served coarse-x reads (512³: 0.65 → 0.54 ms)
#endif
__global__ void __launch_bounds__(256, WL_PROLONG_MINB) k_prolong_increment_init_m(Lvl f, Lvl c, ProMap m, MarchGeom mg, const float* __restrict__ xsrc, ReduceWS ws, FinArgs fa) {
  double acc[2] = {0.0, 0.0};
  for_columns(f.g, mg, [&](int i, int j, int k0, int k1) { tile_prolong(f, c, m, mg, xsrc, i, j, k0, k1, acc[0], acc[1]); });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(FIN_INC_INIT, fa, tot[0], tot[1]);
}

// ---- pcg! (Poisson.jl:123-144) -----------------------------------------------------------
// :125-127  z = ϵ = r·iD ; ρ = r⋅z ; |ρ|<10eps → return      (z is rewritten by k_pcg_mult before
// any read, so only ϵ is stored)
template <int D, int V>
__global__ void __launch_bounds__(256) k_pcg_init(Lvl l, ReduceWS ws, FinArgs fa) {
  double acc[1] = {0.0};
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    const Pack<V> r = ldp<V>(l.r + I), id = ld_iD<V>(l, I, pack_uniform<V>(l, I));
    Pack<V> e;
#pragma unroll
    for (int q = 0; q < V; ++q) { e.v[q] = r.v[q] * id.v[q]; acc[0] = dfma(r.v[q], e.v[q], acc[0]); }
    stp<V>(l.eps + I, e);
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_PCG_INIT, fa, tot[0], 0.0);
}
// :130-132  z = Aϵ ; α = ρ/(z⋅ϵ) ; |α|∉[1e-2,1e2] → return.  The dot runs over the padded arrays
// in the reference; the ghost-shell part (non-zero only at level 1) is in sc->ghost_dot.
template <int D, int V>
__global__ void __launch_bounds__(256) k_pcg_mult(Lvl l, ReduceWS ws, FinArgs fa) {
  if (!l.sc->active) return;
  double acc[1] = {0.0};
  for_rows<V>(l.g, [&](int i, int j, int k, int64_t I) {
    const Pack<V> ec = ldp<V>(l.eps + I), z = mult_w<D, V>(l, l.eps, i, j, k, I, ec, pack_uniform<V>(l, I));
#pragma unroll
    for (int q = 0; q < V; ++q) acc[0] = dfma(z.v[q], ec.v[q], acc[0]);
    stp<V>(l.z + I, z);
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_PCG_MULT, fa, tot[0], 0.0);
}
// :133-139  x += αϵ ; r -= αz ; [last sweep → return] ; z = r·iD ; ρ₂ = r⋅z ; |ρ₂|<10eps → return ;
// β = ρ₂/ρ ; ρ = ρ₂
#ifndef WL_UPD_MINB
#define WL_UPD_MINB 1
#endif
template <int D, int V>
__global__ void __launch_bounds__(256, WL_UPD_MINB) k_pcg_update(Lvl l, ReduceWS ws, FinArgs fa, int last) {
  if (!l.sc->active) return;
  const float alpha = l.sc->alpha;
  double acc[2] = {0.0, 0.0};
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    const Pack<V> e = ldp<V>(l.eps + I), z = ldp<V>(l.z + I), id = ld_iD<V>(l, I, pack_uniform<V>(l, I));
    Pack<V> x = ldp<V>(l.x + I), r = ldp<V>(l.r + I);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      x.v[q] += alpha * e.v[q];
      r.v[q] -= alpha * z.v[q];
      acc[0] = dfma(r.v[q], (r.v[q] * id.v[q]), acc[0]);
      acc[1] = dfma(r.v[q], r.v[q], acc[1]);
    }
    stp<V>(l.x + I, x);
    stp<V>(l.r + I, r);
  });
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(last ? FIN_PCG_UPDATE_LAST : FIN_PCG_UPDATE, fa, tot[0], tot[1]);
}
// :140  ϵ = βϵ + z  with z = r·iD recomputed
template <int D, int V>
__global__ void __launch_bounds__(256) k_pcg_dir(Lvl l) {
  if (!l.sc->active) return;
  const float beta = l.sc->beta;
  if (blockIdx.x == 0 && threadIdx.x + threadIdx.y + threadIdx.z == 0) atomicAdd(&l.sc->cnt_dir, 1);
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    const Pack<V> r = ldp<V>(l.r + I), id = ld_iD<V>(l, I, pack_uniform<V>(l, I));
    Pack<V> e = ldp<V>(l.eps + I);
#pragma unroll
    for (int q = 0; q < V; ++q) e.v[q] = beta * e.v[q] + r.v[q] * id.v[q];
    stp<V>(l.eps + I, e);
  });
}

// Σ_shell z·ϵ  → sc->ghost_dot  (level 1: z ≡ σ keeps the stale Φ of conv_diff! on its upper ghost
// layer, and ϵ's ghosts are non-zero after perBC! — reference quirk Q2, must be reproduced)
__global__ void __launch_bounds__(256) k_shell_dot(Lvl l, ReduceWS ws, FinArgs fa, int64_t count) {
  if (!l.sc->active) return;
  double acc[1] = {0.0};
  for_global_ghosts(l.g, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, [&](const int (&c)[3]) {
    const int64_t I = c[0] + c[1] * l.g.s[1] + c[2] * l.g.s[2];
    // ϵ's periodic ghosts are not materialised (the stencils wrap their index): read the image perBC! would copy
    int w[3] = {c[0], c[1], c[2]};
    for (int d = 0; d < l.g.D; ++d)
      if (l.g.per[d]) w[d] = c[d] == 0 ? l.g.n[d] - 2 : (c[d] == l.g.n[d] - 1 ? 1 : c[d]);
    acc[0] += (double)l.z[I] * (double)l.eps[w[0] + w[1] * l.g.s[1] + w[2] * l.g.s[2]];
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_GHOST_DOT, fa, tot[0], 0.0);
}

template <int D, int V>
__global__ void __launch_bounds__(256) k_dot_rr(Lvl l, ReduceWS ws, FinArgs fa) {
  double acc[1] = {0.0};
  for_rows<V>(l.g, [&](int, int, int, int64_t I) {
    const Pack<V> r = ldp<V>(l.r + I);
#pragma unroll
    for (int q = 0; q < V; ++q) acc[0] = dfma(r.v[q], r.v[q], acc[0]);
  });
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_DOT, fa, tot[0], 0.0);
}

__global__ void k_finalize(int stage, FinArgs fa) {
  if (fin_predicated_off(stage, fa)) return;
  finalize(stage, fa, fa.raw[0], fa.raw[1]);
}

}  // namespace


static int level_of(const Sim& s, const Lvl& l) { return (int)(&l - &s.levels[0]); }

// z-marching geometry of a level, or ok = false (2-D, small or odd-sized levels, WL_MARCH=0)
static MarchLaunch march_for(const Sim& s, const GridD& g) {
  if (!s.march_on) { MarchLaunch m{}; return m; }
  return march_launch(g, s.nsm);
}

FinArgs fin_level(const Sim& s, int level) {
  FinArgs fa{};
  const Lvl& l = s.levels[level];
  fa.sc = l.sc;
  fa.misc = s.d_misc;
  fa.raw = &l.sc->raw[0];
  fa.out = s.ws.partials + (size_t)WL_MAXBLOCKS * WL_MAXRED;  // spare slot behind the partials
  fa.count = (double)l.g.interior_global;
  fa.defer = (s.dist.nranks > 1 && s.lvl_dist[level]) ? (s.dist.p2p ? 2 : 1) : 0;
  fa.p2p = s.dist.dev;
  return fa;
}
FinArgs fin_misc(const Sim& s) {
  FinArgs fa = fin_level(s, 0);
  fa.raw = &s.d_misc->raw[0];
  return fa;
}
// deferred finalisation: all-reduce the parked totals, then run finalize on every rank
int post_reduce(Sim& s, int stage, const FinArgs& fa, bool is_max) {
  if (fa.defer != 1) return 0;  // local, or all-reduced over peer memory by the producing kernel itself
  int rc = dist_allreduce(s, fa.raw, 2, is_max);
  if (rc) return rc;
  k_finalize<<<1, 1, 0, s.stream>>>(stage, fa);
  ++s.launches;
  return wl_cuda_check(s, cudaGetLastError(), "k_finalize");
}

int pois_set_diag(Sim& s, Lvl& l) {
  const int lv = level_of(s, l);
  unsigned char* um = s.umask_store[lv];               // null: rows of this level are not a multiple of four
  unsigned int* cnt = um ? s.d_nonuni + lv : nullptr;
  if (cnt) cudaMemsetAsync(cnt, 0, sizeof(unsigned int), s.stream);
  RLAUNCH(s, k_set_diag, l.g, l, um, cnt);
  l.umask = s.uniform_path ? um : nullptr;             // the mask now describes L, D, iD of this level
  s.nonuni_known = false;
  return wl_cuda_check(s, cudaGetLastError(), "k_set_diag");
}
int pois_uniform_fraction(Sim& s) {
  s.uni_frac0 = 0.0;
  if (s.levels.empty() || !s.levels[0].umask) return 0;
  if (!s.nonuni_known) {
    const int rc = wl_read_scalars(s, s.d_nonuni, sizeof(unsigned int));
    if (rc) return rc;
    unsigned int c;
    memcpy(&c, s.h_pinned, sizeof(c));
    s.nonuni0 = c;
    s.nonuni_known = true;
  }
  const double packs = (double)s.levels[0].g.interior / 4.0;
  s.uni_frac0 = packs > 0 ? 1.0 - (double)s.nonuni0 / packs : 0.0;
  return 0;
}
// L, D or iD of a level changed behind set_diag!'s back (upload, BC! on μ₀): its kernels go back to loading them
void pois_invalidate_uniform(Sim& s, int level) {
  if (level >= 0 && level < (int)s.levels.size()) s.levels[level].umask = nullptr;
}
// restrictL! (MultiLevelPoisson.jl:26-32).  When the fine level is z-distributed and the coarse one
// replicated, every rank restricts its own planes and the planes are all-gathered.
int pois_restrictL(Sim& s, Lvl& c, const Lvl& f) {
  const int lc = level_of(s, c);
  const bool gather = s.dist.nranks > 1 && s.lvl_dist[lc - 1] && !s.lvl_dist[lc];
  int ck0 = 1, ck1 = c.g.n[2] - 2, fshift = 0;
  if (gather) {
    const int cp = (c.g.n[2] - 2) / s.dist.nranks;  // coarse planes per rank
    ck0 = 1 + s.dist.rank * cp; ck1 = ck0 + cp - 1; fshift = -f.g.zoff;
  }
  RLAUNCH(s, k_restrictL, c.g, c, f, ck0, ck1, fshift);
  if (gather) {
    const size_t cnt = (size_t)((c.g.n[2] - 2) / s.dist.nranks) * c.g.s[2];
    for (int i = 0; i < s.D; ++i) {
      const int rc = dist_allgather(s, c.L + i * c.g.comp + c.g.s[2], cnt);
      if (rc) return rc;
    }
  }
  const double zero[3] = {0, 0, 0};
  return bc_vec(s, c.g, c.L, zero, false, c.g.per, 1);  // BC!(a,0,false,perdir) (MultiLevelPoisson.jl:31)
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
  if ((rc = wl_fill(s, l.z - l.g.lead, 0.f, l.g.comp))) return rc;
  RLAUNCH(s, k_mult, l.g, l);
  return wl_cuda_check(s, cudaGetLastError(), "k_mult");
}
// residual!(p) (Poisson.jl:91-97)
int pois_residual(Sim& s, Lvl& l, bool need_r2, bool ghosts_valid, bool jacobi_eps, const float* divu) {
  // perBC!(x) (Poisson.jl:92).  project! hands over an x whose ghost cells and halo planes were scaled
  // together with the interior, i.e. are already consistent: nothing to refresh then.
  int rc = ghosts_valid ? 0 : bc_per(s, l.g, l.x, l.g.per);
  if (rc) return rc;
  const int lv = level_of(s, l);
  const bool fine = lv == 0;
  const FinArgs fa = fin_level(s, lv);
  if (fine) wl_prof_begin(s, WL_PROF_RESIDUAL);
  const MarchLaunch ml = march_for(s, l.g);
  if (ml.ok && ml.staged && s.stage_on) { k_residual_s<<<ml.grid, ml.blk, ml.smem, s.stream>>>(l, ml.mg, s.ws, fa, jacobi_eps ? 1 : 0, divu); ++s.launches; }
  else if (ml.ok) { k_residual_m<<<ml.grid, ml.blk, 0, s.stream>>>(l, ml.mg, s.ws, fa, jacobi_eps ? 1 : 0, divu); ++s.launches; }
  else RLAUNCH(s, k_residual, l.g, l, s.ws, fa, jacobi_eps ? 1 : 0, divu);
  if ((rc = post_reduce(s, FIN_RESIDUAL, fa))) return rc;
  {
    FinArgs fb = fa;
    if (!need_r2 && fb.defer == 2) fb.defer = 0;  // not needed across ranks: keep the partial local
    RLAUNCH(s, k_sub_mean, l.g, l, s.ws, fb, jacobi_eps ? 1 : 0);
  }
  s.eps_is_jacobi = jacobi_eps;
  // r⋅r after the mean subtraction is only reported, never used for control: solver! skips its all-reduce
  if (need_r2 && (rc = post_reduce(s, FIN_R2_IF_SUB, fa))) return rc;
  if (fine) wl_prof_end(s, WL_PROF_RESIDUAL, 4.0 * (double)l.g.interior * wl_words(s, s.D + 5 + (jacobi_eps ? 1 : 0) + (divu ? s.D - 1 : 0), s.D + 2));
  return wl_cuda_check(s, cudaGetLastError(), "k_residual");
}
// increment!(p) (Poisson.jl:99-103)
static int pois_increment(Sim& s, Lvl& l) {
  // perBC!(ϵ) (Poisson.jl:100): the stencil reads periodic neighbours through a wrapped index, so only
  // the slab faces need data (the periodic ghost copies are materialised on demand, see wl_api.cu)
  int rc = dist_halo(s, l.g, l.eps, 1, 1);
  if (rc) return rc;
  const int lv = level_of(s, l);
  const bool fine = lv == 0;
  FinArgs fa = fin_level(s, lv);
  if (!fine && fa.defer == 2) fa.defer = 0;  // L₂ is only ever read on the finest level
  if (fine) wl_prof_begin(s, WL_PROF_INCREMENT);
  const MarchLaunch ml = march_for(s, l.g);
  if (ml.ok && ml.staged && s.stage_on) { k_increment_s<<<ml.grid, ml.blk, ml.smem, s.stream>>>(l, ml.mg, s.ws, fa); ++s.launches; }
  else if (ml.ok) { k_increment_m<<<ml.grid, ml.blk, 0, s.stream>>>(l, ml.mg, s.ws, fa); ++s.launches; }
  else RLAUNCH(s, k_increment, l.g, l, s.ws, fa);
  if (fine && (rc = post_reduce(s, FIN_R2, fa))) return rc;  // L₂ is only ever read on the finest level
  if (fine) wl_prof_end(s, WL_PROF_INCREMENT, 4.0 * (double)l.g.interior * wl_words(s, s.D + 6, s.D + 1));
  return wl_cuda_check(s, cudaGetLastError(), "k_increment");
}
// Jacobi!(p;it) (Poisson.jl:110-113)
int pois_jacobi(Sim& s, Lvl& l, int it) {
  for (int t = 0; t < it; ++t) {
    const bool ready = t == 0 && &l == &s.levels[0] && s.eps_is_jacobi;  // residual! already stored ϵ = r·iD
    s.eps_is_jacobi = false;
    if (!ready) RLAUNCH(s, k_jacobi_eps, l.g, l);
    const int rc = pois_increment(s, l);
    if (rc) return rc;
  }
  return 0;
}
// pcg!(p;it) (Poisson.jl:123-144) — enqueued in full; early exits are device-side predicates
int pois_pcg(Sim& s, int level, int it) {
  Lvl& l = s.levels[level];
  const bool fine = level == 0;
  FinArgs fa = fin_level(s, level);
  int rc;
  bool ghosts = l.g.zlo_halo || l.g.zhi_halo;  // ϵ ghosts need refreshing: periodic copies or slab halos
  for (int k = 0; k < s.D; ++k) ghosts |= l.g.per[k] != 0;
  fa.has_ghost = (fine && ghosts) ? 1 : 0;
  if (s.pcg_init_done == level) {
    rc = 0;  // ϵ = r·iD and ρ were produced by the fused prolongate!/increment! pass of the V-cycle
  } else {
    if (fine) wl_prof_begin(s, WL_PROF_PCG_INIT);
    RLAUNCH(s, k_pcg_init, l.g, l, s.ws, fa);
    if ((rc = post_reduce(s, FIN_PCG_INIT, fa))) return rc;
    if (fine) wl_prof_end(s, WL_PROF_PCG_INIT, 4.0 * (double)l.g.interior * wl_words(s, 3, 1));
  }
  s.pcg_init_done = -1;
  // Sweeps are enqueued speculatively: as many as the previous pcg! on this level needed, then ONE
  // host read of {active, sweep counter} decides whether more are launched.  Exited sweeps would be
  // predicated off on the device anyway; this only avoids enqueueing (and, across ranks,
  // all-reducing) work that is known to be dead.  Every rank reads the same bits.
  if ((int)s.pcg_guess.size() < (int)s.levels.size()) { s.pcg_guess.assign(s.levels.size(), std::min(it, 2)); s.pcg_cnt.assign(s.levels.size(), 0); }
  int& guess = s.pcg_guess[level];
  const MarchLaunch mlp = march_for(s, l.g);
  int i = 1;
  while (i <= it) {
    const int upto = std::min(it, std::max(guess, i));
    for (; i <= upto; ++i) {
      if (ghosts) {
        if ((rc = dist_halo(s, l.g, l.eps, 1, 1))) return rc;  // perBC!(ϵ) (:129), see pois_increment
        if (fine) {  // only level 1 has non-zero z ghosts (z ≡ σ)
          const int64_t cnt = global_ghost_count(l.g);
          k_shell_dot<<<(int)std::min<int64_t>((cnt + 255) / 256, (int64_t)s.nsm * 4), 256, 0, s.stream>>>(l, s.ws, fa, cnt);
          ++s.launches;  // multi-rank: its partial is parked next to the z⋅ϵ partial and shares that all-reduce
        }
      }
      if (fine) wl_prof_begin(s, WL_PROF_PCG_MULT);
      if (mlp.ok && mlp.staged && s.stage_on) { k_pcg_mult_s<<<mlp.grid, mlp.blk, mlp.smem, s.stream>>>(l, mlp.mg, s.ws, fa); ++s.launches; }
      else if (mlp.ok) { k_pcg_mult_m<<<mlp.grid, mlp.blk, 0, s.stream>>>(l, mlp.mg, s.ws, fa); ++s.launches; }
      else RLAUNCH(s, k_pcg_mult, l.g, l, s.ws, fa);
      if ((rc = post_reduce(s, FIN_PCG_MULT, fa))) return rc;
      if (fine) { wl_prof_end(s, WL_PROF_PCG_MULT, 0); wl_prof_begin(s, WL_PROF_PCG_UPDATE); }
      RLAUNCH(s, k_pcg_update, l.g, l, s.ws, fa, i == it ? 1 : 0);
      if ((rc = post_reduce(s, i == it ? FIN_PCG_UPDATE_LAST : FIN_PCG_UPDATE, fa))) return rc;
      if (fine) wl_prof_end(s, WL_PROF_PCG_UPDATE, 0);
      if (i == it) { i = it + 1; break; }
      if (fine) wl_prof_begin(s, WL_PROF_PCG_DIR);
      RLAUNCH(s, k_pcg_dir, l.g, l);
      if (fine) wl_prof_end(s, WL_PROF_PCG_DIR, 0);
    }
    if (i > it) break;  // all sweeps enqueued
    if ((rc = wl_read_scalars(s, l.sc, sizeof(LevelScalars)))) return rc;
    LevelScalars h;
    memcpy(&h, s.h_pinned, sizeof(h));
    const int ran = h.cnt_mult - s.pcg_cnt[level];  // sweeps that really ran since the last check
    s.pcg_cnt[level] = h.cnt_mult;
    if (!h.active) { guess = std::max(1, std::min(it, ran)); break; }
    guess = i;  // still iterating: one more sweep per check from here on
  }
  return wl_cuda_check(s, cudaGetLastError(), "pcg");
}
// Vcycle!(ml;l) (MultiLevelPoisson.jl:70-82)
int pois_vcycle(Sim& s, int level) {
  Lvl& fine = s.levels[level];
  Lvl& coarse = s.levels[level + 1];
  int rc = 0;
  const bool gather = s.dist.nranks > 1 && s.lvl_dist[level] && !s.lvl_dist[level + 1];
  const bool head = level == 0 && s.head_fused;  // Jacobi! + restrict! of this cycle ran inside the fused head of solver! (wl_resjac.cu)
  if (head) {
    s.head_fused = false;  // (s.x_parked, set by the head: the finest x is in fine.eps until the prolongation pass below)
  } else {
    if ((rc = pois_jacobi(s, fine, 1))) return rc;
    // restrict! + fill!(coarse.x,0); a distributed→replicated transition gathers the planes
    int ck0 = 1, ck1 = coarse.g.n[2] - 2, fshift = 0;
    if (gather) {
      const int cp = (coarse.g.n[2] - 2) / s.dist.nranks;
      ck0 = 1 + s.dist.rank * cp; ck1 = ck0 + cp - 1; fshift = -fine.g.zoff;
    }
    RLAUNCH(s, k_restrict, coarse.g, coarse, fine, ck0, ck1, fshift);
    if (gather && (rc = dist_allgather(s, coarse.r + coarse.g.s[2], (size_t)((coarse.g.n[2] - 2) / s.dist.nranks) * coarse.g.s[2]))) return rc;
  }
  if (s.vc_first == level + 1) {
    // the rest of the cycle below this level is one persistent kernel (wl_vcycle.cu)
    if ((rc = vcycle_fused_run(s))) return rc;
  } else {
    if (level + 2 < (int)s.levels.size() && (rc = pois_vcycle(s, level + 1))) return rc;
    if ((rc = pois_pcg(s, level + 1, 6))) return rc;
  }
  if (s.fuse_prolong) {
    // inside solver!: prolongate! + increment! + the head of the pcg! that always follows, in one pass
    const bool fdist = s.dist.nranks > 1 && s.lvl_dist[level];
    if (fdist && !gather && (rc = dist_halo(s, coarse.g, coarse.x, 1, 1))) return rc;  // ϵ across slab faces = the neighbour's coarse x
    ProMap m{};
    for (int d = 0; d < 3; ++d) { m.per[d] = fine.g.per[d]; m.n[d] = fine.g.n[d]; }
    m.zmode = gather ? 1 : 0;
    m.zoff = fine.g.zoff;
    m.nzg = s.Nz_global >> level;
    m.perz_global = s.per_z_global;
    const FinArgs fa = fin_level(s, level);  // ρ steers the pcg! of this level: all-reduced on every distributed level
    const bool top = level == 0;
    if (top) wl_prof_begin(s, WL_PROF_INCREMENT);
    const MarchLaunch ml = march_for(s, fine.g);
    const float* xsrc = (level == 0 && s.x_parked) ? fine.eps : fine.x;  // same cell read and (as the new ϵ) written by one thread: no hazard
    if (level == 0) s.x_parked = false;
    if (ml.ok && (coarse.g.n[0] - 2) % 2 == 0) { k_prolong_increment_init_m<<<ml.grid, ml.blk, 0, s.stream>>>(fine, coarse, m, ml.mg, xsrc, s.ws, fa); ++s.launches; }
    else RLAUNCH(s, k_prolong_increment_init, fine.g, fine, coarse, m, xsrc, s.ws, fa);
    if ((rc = post_reduce(s, FIN_INC_INIT, fa))) return rc;
    if (top) wl_prof_end(s, WL_PROF_INCREMENT, 4.0 * (double)fine.g.interior * wl_words(s, s.D + 7.125, s.D + 2));
    s.pcg_init_done = level;
    return wl_cuda_check(s, cudaGetLastError(), "k_prolong_increment_init");
  }
  RLAUNCH(s, k_prolongate, fine.g, fine, coarse, gather ? fine.g.zoff : 0);
  return pois_increment(s, fine);
}
// residual!(fine) and the opening of the first Vcycle! — Jacobi!(fine), restrict!, fill!(coarse.x,0)
// (MultiLevelPoisson.jl:88,72-75) — as their own entry: the fused pass where the grid allows, else the separate kernels
int pois_head(Sim& s) {
  Lvl& fine = s.levels[0];
  Lvl& coarse = s.levels[1];
  int rc;
  if (resjac_eligible(s)) {
    if ((rc = resjac_run(s))) return rc;
    s.eps_is_jacobi = false;
    s.x_parked = false;
    return resjac_unpark(s);
  }
  if ((rc = pois_residual(s, fine, false, false, true, nullptr))) return rc;
  if ((rc = pois_jacobi(s, fine, 1))) return rc;
  if (s.dist.nranks > 1 && s.lvl_dist[0] && !s.lvl_dist[1]) return wl_fail(WL_ERR_ARG, "wl_pois_head: not available across a distributed/replicated level transition");
  RLAUNCH(s, k_restrict, coarse.g, coarse, fine, 1, coarse.g.n[2] - 2, 0);
  return wl_cuda_check(s, cudaGetLastError(), "k_restrict");
}
static int fuse_prolong_env() {
  static int fuse_env = -1;
  if (fuse_env < 0) { const char* e = getenv("WL_FUSE_PROLONG"); fuse_env = e ? atoi(e) : 1; }
  return fuse_env;
}
bool pois_head_fusable(const Sim& s) {
  return s.pois_kind == WL_POIS_MULTILEVEL && !s.vc_solve && fuse_prolong_env() != 0 && resjac_eligible(s);
}
// solver!(ml;tol,itmx) (MultiLevelPoisson.jl:87-99) / solver!(p::Poisson;tol,itmx) (Poisson.jl:162-172)
int pois_solve(Sim& s, double tol, int itmx, int* n_out, bool ghosts_valid, int slot, bool defer_count, const float* divu) {
  Lvl& p = s.levels[0];
  wl_prof_begin(s, WL_PROF_SOLVER);
  if (s.vc_solve && s.pois_kind == WL_POIS_MULTILEVEL) {
    // the whole solver! is one persistent kernel (wl_vcycle.cu); its V-cycle count stays on the device
    // until the caller reads it (mom_step! reads both counts together with CFL's maximum)
    int rc = vcycle_fused_solve(s, tol, itmx, ghosts_valid, slot, divu);
    wl_prof_end(s, WL_PROF_SOLVER, 0.0);
    if (rc || defer_count) return rc;
    if ((rc = wl_read_scalars(s, &s.d_misc->solve_n[slot], sizeof(int)))) return rc;
    int n;
    memcpy(&n, s.h_pinned, sizeof(int));
    s.n_log.push_back(n);
    if (n_out) *n_out = n;
    return 0;
  }
  const int fuse_env = fuse_prolong_env();
  int rc;
  s.head_fused = false;
  s.x_parked = false;
  if (itmx > 0 && pois_head_fusable(s)) {
    // residual! and the Jacobi! + restrict! that open the first Vcycle! as one pass (wl_resjac.cu); project!'s z = div u is
    // evaluated inside it (x needs no perBC!: the pass wraps its own periodic reads)
    if (divu && !s.resjac_div) { if ((rc = mom_div(s, divu, p.z))) return rc; divu = nullptr; }
    if ((rc = resjac_run(s, divu))) return rc;
    s.eps_is_jacobi = false;
    s.head_fused = true;
  } else {
    rc = pois_residual(s, p, false, ghosts_valid, s.pois_kind == WL_POIS_MULTILEVEL && itmx > 0, divu);
    if (rc) return rc;
  }
  int n = 0;
  s.fuse_prolong = fuse_env != 0;
  s.pcg_init_done = -1;
  while (n < itmx) {
    if (s.pois_kind == WL_POIS_MULTILEVEL && (rc = pois_vcycle(s, 0))) { s.fuse_prolong = false; return rc; }
    if ((rc = pois_pcg(s, 0, 6))) { s.fuse_prolong = false; return rc; }
    ++n;
    // r₂ = L₂(p): one host read per solver iteration (the reference needs ~100); every rank reads
    // the same all-reduced value and therefore leaves the loop together
    if ((rc = wl_read_scalars(s, &p.sc->r2, sizeof(float)))) return rc;
    if ((double)s.h_pinned[0] < tol) break;
  }
  s.fuse_prolong = false;
  rc = bc_per(s, p.g, p.x, p.g.per);
  wl_prof_end(s, WL_PROF_SOLVER, 0.0);
  s.n_log.push_back(n);
  if (n_out) *n_out = n;
  return rc;
}
// perBC!(ϵ) made visible: the periodic ghost cells of ϵ on levels ≥ `level` (API inspection only)
int pois_materialize_eps(Sim& s, int level) {
  for (size_t l = (size_t)level; l < s.levels.size(); ++l) {
    const int rc = bc_per(s, s.levels[l].g, s.levels[l].eps, s.levels[l].g.per);
    if (rc) return rc;
  }
  return 0;
}
int pois_dot_rr(Sim& s, Lvl& l, double* out) {
  FinArgs fa = fin_level(s, level_of(s, l));
  RLAUNCH(s, k_dot_rr, l.g, l, s.ws, fa);
  int rc = post_reduce(s, FIN_DOT, fa);
  if (rc) return rc;
  if ((rc = wl_read_scalars(s, fa.out, sizeof(double)))) return rc;
  double v;
  memcpy(&v, s.h_pinned, sizeof(double));
  *out = v;
  return 0;
}
