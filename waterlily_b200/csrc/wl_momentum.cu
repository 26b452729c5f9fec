// wl_momentum.cu — momentum-step kernels (reference src/Flow.jl): conv_diff! (QUICK/central flux
// stencil), accelerate!, BDIM!, scale_u!, project! (div / pressure correction), CFL, mom_step!.
//
// Fusion relative to the reference's pass structure:
//  * conv_diff!'s 1+4D² single-statement launches are ONE kernel that, per cell and component,
//    accumulates the 2D face fluxes in the reference's order; `accelerate!` and BDIM! stage 1
//    (f = u⁰ + Δt·f − V) are applied in the same kernel before the only store of f;
//  * `u⁰ .= u` is a pointer swap and `scale_u!` is folded into BDIM! stage 2;
//  * the reference's ghost-layer side effects are reproduced: r/f on the upper ghost layer (Q0)
//    and the stale flux Φ left in σ's ghosts (Q1/Q2) — the kernel covers the padded box and
//    evaluates exactly the passes that reach each ghost cell.
// Float64 sub-expressions follow Julia's promotion rules.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "wl_internal.h"
#include "wl_vec.cuh"
#include "wl_flux.cuh"

namespace {

// Φ[I] of the inner loop (Flow.jl:45) / periodic lower boundary (:58-59); `upup` is the offset of the
// upwind-upwind sample (−2sj, or the periodic wrap to N[j]-2).
__device__ __forceinline__ float phi_face(const float* __restrict__ ui, const float* __restrict__ uj, int64_t I, int64_t si, int64_t sj,
                                          int64_t upup, float nu) {
  const float uf = (uj[I] + uj[I - si]) * 0.5f;  // ϕ(i,CI(I,j),u): exact in Float32
  const float c = ui[I], m = ui[I - sj];
  const float q = uf > 0 ? quick(ui[I + upup], m, c) : quick(ui[I + sj], c, m);
  return (float)((double)uf * (double)q - (double)(nu * (c - m)));
}

template <int D>
__device__ __forceinline__ void convdiff_cell(const CDArgs& a, const int (&c)[3], int64_t I, float (&r)[D]) {
  const GridD& g = a.g;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float* ui = a.u + i * g.comp;
    const int64_t si = g.s[i];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const int n = g.n[j], cj = c[j];
      if (cj < 1 || cj > n - 2) continue;  // no pass of direction j reaches a j-ghost cell
      const float* uj = a.u + j * g.comp;
      const int64_t sj = g.s[j];
      // flux through the lower j-face: lowerBoundary! (Flow.jl:54,58) or inner loop (:45-46)
      if (cj == 1) {
        if (g.per[j]) {
          acc += phi_face(ui, uj, I, si, sj, (int64_t)(n - 4) * sj, a.nu);
        } else if (j == 2 && g.zlo_halo) {  // slab face: a plain inner cell, I-2δ is in the depth-2 halo
          acc += phi_face(ui, uj, I, si, sj, -2 * sj, a.nu);
        } else {  // ϕuL, accumulated in Float64
          const double uf = (double)(uj[I] + uj[I - si]) * 0.5;
          const double fl = uf > 0 ? uf * ((double)(ui[I] + ui[I - sj]) * 0.5) : uf * (double)quick(ui[I + sj], ui[I], ui[I - sj]);
          acc = (float)((double)acc + (fl - (double)(a.nu * (ui[I] - ui[I - sj]))));
        }
      } else {
        acc += phi_face(ui, uj, I, si, sj, -2 * sj, a.nu);
      }
      // flux through the upper j-face: inner loop (:47) or upperBoundary! (:55,60)
      if (cj == n - 2) {
        if (g.per[j]) {
          acc -= phi_face(ui, uj, I + (int64_t)(1 - cj) * sj, si, sj, (int64_t)(n - 4) * sj, a.nu);
        } else if (j == 2 && g.zhi_halo) {
          acc -= phi_face(ui, uj, I + sj, si, sj, -2 * sj, a.nu);
        } else {  // ϕuR
          const int64_t J = I + sj;
          const double uf = (double)(uj[J] + uj[J - si]) * 0.5;
          const double fl = uf < 0 ? uf * ((double)(ui[J] + ui[J - sj]) * 0.5) : uf * (double)quick(ui[J - 2 * sj], ui[J - sj], ui[J]);
          acc = (float)((double)acc + (-fl + (double)(a.nu * (ui[J] - ui[J - sj]))));
        }
      } else {
        acc -= phi_face(ui, uj, I + sj, si, sj, -2 * sj, a.nu);
      }
    }
    r[i] = acc;
  }
}

// one cell of conv_diff! (+ accelerate! + BDIM! stage 1), any position in the padded box
template <int D>
__device__ __forceinline__ void convdiff_store_cell(const CDArgs& a, const int (&c)[3], int64_t I) {
  const GridD& g = a.g;
  float r[D];
  bool lowghost = false, ghost = false;
#pragma unroll
  for (int k = 0; k < D; ++k) { lowghost |= c[k] == 0; ghost |= c[k] == g.n[k] - 1; }
  // a z-plane owned by the neighbouring slab: its f arrives by halo exchange, nothing to compute
  if (D > 2 && c[2] == g.n[2] - 1 && g.zhi_halo) lowghost = true;
  if (lowghost) {
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = 0.f;  // index-1 ghosts are never written (r .= 0)
  } else {
    convdiff_cell<D>(a, c, I, r);
    if (ghost) {
      // stale Φ: last pass (i=D, largest j) that covers this upper-ghost cell
      for (int j = D - 1; j >= 0; --j) {
        const int lo = (g.per[j] || (j == 2 && g.zlo_halo)) ? 1 : 2;
        if (c[j] >= lo && c[j] <= g.n[j] - 2) {
          const int64_t sj = g.s[j];
          a.sigma[I] = phi_face(a.u + (D - 1) * g.comp, a.u + j * g.comp, I, g.s[D - 1], sj,
                                (c[j] == 1 && g.per[j]) ? (int64_t)(g.n[j] - 4) * sj : -2 * sj, a.nu);
          break;
        }
      }
    }
  }
  // BDIM! stage 2 on the spot (fused step, interior cell whose own μ₁ is zero: no neighbour of f is needed)
  bool stage2 = false;
  unsigned char mk = 0;
  if (a.ufused) {
    stage2 = true;
#pragma unroll
    for (int k = 0; k < D; ++k) stage2 &= c[k] >= 1 && c[k] <= g.n[k] - 2;
    if (stage2) { mk = a.mask[I]; stage2 = (mk & 2) == 0; }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float v = r[i];
    if (a.has_g) v = (float)((double)v + a.gacc[i]);          // accelerate! (Flow.jl:68-70)
    if (a.fuse_bdim1) v = a.u0[I + i * g.comp] + a.dt * v - a.V[I + i * g.comp];  // BDIM! stage 1 (:133)
    a.f[I + i * g.comp] = v;
    if (stage2) {  // the expressions of k_bdim2 with μ₁ = 0 (s = 0.f + 0·(…) = +0)
      const int64_t Ii = I + i * g.comp;
      const double acc = a.uacc ? (double)a.uacc[Ii] : 0.0;
      if (mk == 0) {
        a.ufused[Ii] = (float)(acc + (double)v) * a.bscale;
      } else {
        const double add = (0.5 * (double)0.f + (double)a.V[Ii]) + (double)(a.mu0[Ii] * v);
        a.ufused[Ii] = (float)(acc + add) * a.bscale;
      }
    }
  }
}

// generic path: the whole padded box (2-D and small grids)
template <int D>
__global__ void __launch_bounds__(256) k_conv_diff(CDArgs a) {
  const Box b = box_all(a.g);
  for_cells(a.g, b, [&](int i0, int j0, int k0, int64_t I) {
    const int c[3] = {i0, j0, k0};
    convdiff_store_cell<D>(a, c, I);
  });
}
// generic path on the boundary shell of thickness 2 only (the marching kernel does the core).
// The shell is walked as three groups with 32-bit index arithmetic and x fastest wherever a group has
// an x extent: full z-planes outside the core's z-range, then the 4 outer y-rows of every core plane,
// then the 4 outer x-cells of every remaining row.
#ifndef WL_SHELL_MINB
#define WL_SHELL_MINB 5  // 48 registers (a few spills): the shell is latency-bound, 40 resident warps instead of 32 (512³: 0.33 → 0.17 ms per launch)
#endif
#ifndef WL_PCC_MINB
#define WL_PCC_MINB 5    // k_project_correct_cfl: 48 instead of 54 registers, five resident blocks (512³: 1.03 → 0.95 ms)
#endif
template <int D>
__global__ void __launch_bounds__(256, WL_SHELL_MINB) k_conv_diff_shell(CDArgs a, unsigned int nZ, unsigned int nY, unsigned int nX, int xlow) {
  const GridD& g = a.g;
  const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const int zc0 = g.zlo_halo ? 1 : 2, zc1 = g.zhi_halo ? n2 - 2 : n2 - 3;  // core z-range
  const unsigned int total = nZ + nY + nX;
  for (unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int c[3];
    if (idx < nZ) {
      const unsigned int row = idx / (unsigned int)n0, kk = row / (unsigned int)n1;
      c[0] = (int)(idx - row * n0);
      c[1] = (int)(row - kk * n1);
      c[2] = (int)kk < zc0 ? (int)kk : zc1 + 1 + ((int)kk - zc0);
    } else if (idx < nZ + nY) {
      const unsigned int t = idx - nZ, row = t / (unsigned int)n0, kk = row >> 2, jj = row & 3u;
      c[0] = (int)(t - row * n0);
      c[1] = jj < 2 ? (int)jj : n1 - 4 + (int)jj;
      c[2] = zc0 + (int)kk;
    } else {
      // the xlow low columns (the core starts at x = xlow) and the 2 high ones of every remaining row
      const unsigned int t = idx - nZ - nY, nxs = (unsigned int)xlow + 2u, row = t / nxs, ii = t - row * nxs, kk = row / (unsigned int)(n1 - 4);
      c[0] = (int)ii < xlow ? (int)ii : n0 - 2 + ((int)ii - xlow);
      c[1] = 2 + (int)(row - kk * (n1 - 4));
      c[2] = zc0 + (int)kk;
    }
    convdiff_store_cell<D>(a, c, c[0] + c[1] * g.s[1] + c[2] * g.s[2]);
  }
}

// ---- 3-D core kernel --------------------------------------------------------------------------
// Cells with every coordinate in [2,n-3] see only the plain inner-loop stencil (Flow.jl:45-47).
// A block owns a 128×4 (x,y) tile and marches along z:
//  * each thread keeps its column's five z-samples of the three components in registers, so one
//    new plane (3 loads) per step serves the whole z-stencil;
//  * every face flux is computed once: the z-flux of the upper face is carried to the next step in
//    a register, x-fluxes travel between lanes by shuffle (between warps through shared memory),
//    y-fluxes between warps through shared memory (double-buffered: one barrier per step);
//  * Φ = uf·λ − ν∂ is one fmaf: the reference rounds the Float64 expression once (Appendix A).
template <int CD_TX, int CD_TY, int MINB, bool HASG, bool CORR>
__global__ void __launch_bounds__(CD_TX * CD_TY, MINB) k_conv_diff3_march(CDArgs a, int zchunk, int ntx, int nty) {
  __shared__ float Fx[2][3][CD_TY][CD_TX / 32];
  __shared__ float Fy[2][3][CD_TY][CD_TX];
  const GridD& g = a.g;
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  const int64_t s1 = g.s[1], s2 = g.s[2];
  const int tix = threadIdx.x, tiy = threadIdx.y, lane = tix & 31, wx = tix >> 5;
  const int bt = blockIdx.x;
  const int tx = bt % ntx, ty = (bt / ntx) % nty, tz = bt / (ntx * nty);
  const int x = 2 + tx * CD_TX + tix, y = 2 + ty * CD_TY + tiy;
  const int zlo = g.zlo_halo ? 1 : 2, zhi = g.zhi_halo ? nz - 2 : nz - 3;  // slab faces are plain inner cells
  const int z0 = zlo + tz * zchunk, z1 = min(z0 + zchunk, zhi + 1);
  const bool valid = x <= nx - 3 && y <= ny - 3;
  const int xc = min(x, nx - 3), yc = min(y, ny - 3);  // clamped: out-of-range threads shadow the last cell
  const bool extra_x = (xc == nx - 3) || (tix == CD_TX - 1);
  const bool extra_y = (yc == ny - 3) || (tiy == CD_TY - 1);
  const float nu = a.nu;
  int64_t I = xc + yc * s1 + (int64_t)(z0 - 1) * s2;
  // per-component pointers to the current cell; all neighbour offsets are 32-bit
  const float* __restrict__ U[3] = {a.u + I, a.u + g.comp + I, a.u + 2 * g.comp + I};
  const int o1 = (int)s1, o2 = (int)s2;
  float w[3][5];  // w[c][k] = u_c at z-2+k
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    w[c][0] = 0.f;
    w[c][1] = U[c][-o2];
    w[c][2] = U[c][0];
    w[c][3] = U[c][o2];
  }
  float fz_lo[3] = {0.f, 0.f, 0.f};
  int buf = 0;
  for (int z = z0 - 1; z < z1; ++z, I += s2, U[0] += o2, U[1] += o2, U[2] += o2) {
#pragma unroll
    for (int c = 0; c < 3; ++c) w[c][4] = U[c][2 * o2];
    // upper z-face = lower z-face of cell z+1: uf = (u_z[J] + u_z[J-δi])/2, J = I+δz
    float fz_hi[3];
    {
      const float uz = w[2][3];
      const float ufx = (uz + U[2][o2 - 1]) * 0.5f, ufy = (uz + U[2][o2 - o1]) * 0.5f, ufz = (uz + w[2][2]) * 0.5f;
      fz_hi[0] = flux1(ufx, w[0][1], w[0][2], w[0][3], w[0][4], nu);
      fz_hi[1] = flux1(ufy, w[1][1], w[1][2], w[1][3], w[1][4], nu);
      fz_hi[2] = flux1(ufz, w[2][1], w[2][2], w[2][3], w[2][4], nu);
    }
    if (z >= z0) {  // uniform per block
      float xm2[3], xm1[3], xp1[3], ym2[3], ym1[3], yp1[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        xm2[c] = U[c][-2]; xm1[c] = U[c][-1]; xp1[c] = U[c][1];
        ym2[c] = U[c][-2 * o1]; ym1[c] = U[c][-o1]; yp1[c] = U[c][o1];
      }
      float fx_lo[3], fy_lo[3], fx_hi[3], fy_hi[3];
      {
        const float ux = w[0][2];
        fx_lo[0] = flux1((ux + xm1[0]) * 0.5f, xm2[0], xm1[0], w[0][2], xp1[0], nu);
        fx_lo[1] = flux1((ux + ym1[0]) * 0.5f, xm2[1], xm1[1], w[1][2], xp1[1], nu);
        fx_lo[2] = flux1((ux + w[0][1]) * 0.5f, xm2[2], xm1[2], w[2][2], xp1[2], nu);
        const float uy = w[1][2];
        fy_lo[0] = flux1((uy + xm1[1]) * 0.5f, ym2[0], ym1[0], w[0][2], yp1[0], nu);
        fy_lo[1] = flux1((uy + ym1[1]) * 0.5f, ym2[1], ym1[1], w[1][2], yp1[1], nu);
        fy_lo[2] = flux1((uy + w[1][1]) * 0.5f, ym2[2], ym1[2], w[2][2], yp1[2], nu);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (lane == 0) Fx[buf][i][tiy][wx] = fx_lo[i];
        Fy[buf][i][tiy][tix] = fy_lo[i];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        fx_hi[i] = __shfl_down_sync(0xffffffffu, fx_lo[i], 1);
        if (lane == 31 && wx < CD_TX / 32 - 1) fx_hi[i] = Fx[buf][i][tiy][wx + 1];
        fy_hi[i] = tiy < CD_TY - 1 ? Fy[buf][i][tiy + 1][tix] : 0.f;
      }
      if (extra_x) {  // lower x-face of cell x+1 that no thread of the block owns
        const float ux = xp1[0];
        fx_hi[0] = flux1((ux + w[0][2]) * 0.5f, xm1[0], w[0][2], xp1[0], U[0][2], nu);
        fx_hi[1] = flux1((ux + U[0][1 - o1]) * 0.5f, xm1[1], w[1][2], xp1[1], U[1][2], nu);
        fx_hi[2] = flux1((ux + U[0][1 - o2]) * 0.5f, xm1[2], w[2][2], xp1[2], U[2][2], nu);
      }
      if (extra_y) {
        const float uy = yp1[1];
        fy_hi[0] = flux1((uy + U[1][o1 - 1]) * 0.5f, ym1[0], w[0][2], yp1[0], U[0][2 * o1], nu);
        fy_hi[1] = flux1((uy + w[1][2]) * 0.5f, ym1[1], w[1][2], yp1[1], U[1][2 * o1], nu);
        fy_hi[2] = flux1((uy + U[1][o1 - o2]) * 0.5f, ym1[2], w[2][2], yp1[2], U[2][2 * o1], nu);
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float v = 0.f;
          v += fx_lo[i]; v -= fx_hi[i];
          v += fy_lo[i]; v -= fy_hi[i];
          v += fz_lo[i]; v -= fz_hi[i];
          if (HASG) v = (float)((double)v + a.gacc[i]);
          if (a.fuse_bdim1) {
            const float u0v = CORR ? a.u0[I + i * g.comp] : w[i][2];
            v = u0v + a.dt * v - a.V[I + i * g.comp];
          }
          a.f[I + i * g.comp] = v;
        }
      }
      buf ^= 1;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      w[c][0] = w[c][1]; w[c][1] = w[c][2]; w[c][2] = w[c][3]; w[c][3] = w[c][4];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) fz_lo[i] = fz_hi[i];
  }
}

// accelerate! alone (API parity hook)
__global__ void __launch_bounds__(256) k_accelerate(GridD g, float* f, double g0, double g1, double g2) {
  const double ga[3] = {g0, g1, g2};
  const Box b = box_all(g);
  for_cells(g, b, [&](int, int, int, int64_t I) {
    for (int i = 0; i < g.D; ++i) f[I + i * g.comp] = (float)((double)f[I + i * g.comp] + ga[i]);
  });
}
// BDIM! stage 1 alone (API parity hook)
__global__ void __launch_bounds__(256) k_bdim1(GridD g, float* f, const float* u0, const float* V, float dt) {
  const Box b = box_all(g);
  for_cells(g, b, [&](int, int, int, int64_t I) {
    for (int i = 0; i < g.D; ++i) { const int64_t Ii = I + i * g.comp; f[Ii] = u0[Ii] + dt * f[Ii] - V[Ii]; }
  });
}

// "Body band" mask: 0 where BDIM! stage 2 degenerates to u += f (μ₁ = 0, V = 0, μ₀ = 1 for every
// component — everywhere except within 2+ϵ cells of a body and on the domain-boundary faces), non-zero
// elsewhere (bit 1 set where some μ₁ ≠ 0, i.e. where stage 2 also reads f of the six neighbours).  One
// byte per cell replaces 15 words of μ₁/V/μ₀ reads; rebuilt whenever one of the three fields changes
// (upload, BC!).  The result is bit-identical: with those values the general expression reduces to
// (Float64)u + (Float64)f exactly.
template <int D, int V>
__global__ void __launch_bounds__(256) k_body_mask(GridD g, const float* __restrict__ Vb, const float* __restrict__ mu0,
                                                   const float* __restrict__ mu1, unsigned char* __restrict__ mask) {
  for_rows<V>(g, [&](int, int, int, int64_t I) {
    unsigned char m[V];
#pragma unroll
    for (int q = 0; q < V; ++q) m[q] = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const Pack<V> vb = ldp<V>(Vb + I + i * g.comp), m0 = ldp<V>(mu0 + I + i * g.comp);
#pragma unroll
      for (int q = 0; q < V; ++q) m[q] |= (vb.v[q] != 0.f) | (m0.v[q] != 1.f);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const Pack<V> m1 = ldp<V>(mu1 + I + (int64_t)(i + D * j) * g.comp);
#pragma unroll
        for (int q = 0; q < V; ++q) m[q] |= (m1.v[q] != 0.f) ? 3 : 0;
      }
    }
#pragma unroll
    for (int q = 0; q < V; ++q) mask[I + q] = m[q];
  });
}

// slab decomposition: the same for the interior cells of one z-plane (the neighbour's cells in a halo plane —
// their μ₁ decides whether this rank's adjacent tiles may skip f, see k_tile_clean)
template <int D>
__global__ void __launch_bounds__(256) k_body_mask_plane(GridD g, const float* __restrict__ Vb, const float* __restrict__ mu0,
                                                         const float* __restrict__ mu1, unsigned char* __restrict__ mask, int k) {
  Box b = box_interior(g);
  b.lo[2] = b.hi[2] = k;
  for_cells(g, b, [&](int, int, int, int64_t I) {
    unsigned char m = 0;
    for (int i = 0; i < D; ++i) {
      m |= (Vb[I + i * g.comp] != 0.f) | (mu0[I + i * g.comp] != 1.f);
      for (int j = 0; j < D; ++j) m |= (mu1[I + (int64_t)(i + D * j) * g.comp] != 0.f) ? 3 : 0;
    }
    mask[I] = m;
  });
}
// Which (tile, plane) units of the conv_diff! core kernel may fold BDIM! stage 2 into their store: every cell of
// the tile in that plane is plain (mask = 0), and no cell within one cell of the tile — in x, y or z — has μ₁ ≠ 0
// (such a cell reads f of its six neighbours, so the tile must deliver f).  One warp per unit.
__global__ void __launch_bounds__(256) k_tile_clean(GridD g, const unsigned char* __restrict__ mask, unsigned char* __restrict__ clean, int ntx,
                                                    int nty, int zlo, int zhi) {
  const int lane = threadIdx.x & 31;
  const int64_t nunits = (int64_t)ntx * nty * g.n[2];
  for (int64_t unit = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < nunits; unit += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int tx = (int)(unit % ntx), ty = (int)((unit / ntx) % nty), z = (int)(unit / ((int64_t)ntx * nty));
    bool ok = z >= zlo && z <= zhi;
    if (ok) {
      const int x0 = WL_CD_X0 + tx * WL_CD_TX, x1 = min(x0 + WL_CD_TX - 1, g.n[0] - 3);
      const int y0 = WL_CD_Y0 + ty * WL_CD_TY, y1 = min(y0 + WL_CD_TY - 1, g.n[1] - 3);
      const int wx = x1 - x0 + 3, wy = y1 - y0 + 3;
      bool bad = false;
      for (int t = lane; t < wx * wy * 3; t += 32) {
        const int dx = t % wx, q = t / wx, dy = q % wy, dz = q / wy;
        const int x = x0 - 1 + dx, y = y0 - 1 + dy, zz = z - 1 + dz;
        const unsigned char m = mask[x + y * g.s[1] + zz * g.s[2]];
        const bool inside = dx >= 1 && dx <= wx - 2 && dy >= 1 && dy <= wy - 2 && dz == 1;
        bad |= inside ? (m != 0) : ((m & 2) != 0);
      }
      ok = !__any_sync(0xffffffffu, bad);
    }
    if (lane == 0) clean[unit] = ok ? 1 : 0;
  }
}

// BDIM! stage 2 (Flow.jl:134) with μddn (:18-24); `accumulate`=0 folds the predictor's
// scale_u!(a,0), `scale` folds the corrector's scale_u!(a,0.5)
// `u` receives the result, `ua` is the u that is accumulated (the same buffer unless the corrector writes the new
// u elsewhere, see mom_step).  Cells of a clean tile-plane (k_tile_clean) were already finished by the conv_diff! core
// kernel: they are skipped (`clean` null = none).
struct FusedTiles { const unsigned char* clean; int ntx, nty, zlo, zhi; };
template <int D, int V>
__global__ void __launch_bounds__(256) k_bdim2(GridD g, float* __restrict__ u, const float* __restrict__ ua, const float* __restrict__ f,
                                               const float* __restrict__ Vb, const float* __restrict__ mu0, const float* __restrict__ mu1,
                                               const unsigned char* __restrict__ mask, FusedTiles ft, int accumulate, float scale) {
  for_rows<V>(g, [&](int i0, int j0, int k0, int64_t I) {
    // bit q set: cell i0+q is done already
    unsigned int done = 0;
    if (D == 3 && ft.clean && k0 >= ft.zlo && k0 <= ft.zhi && j0 >= WL_CD_Y0 && j0 <= g.n[1] - 3) {
      const unsigned char* row = ft.clean + ((size_t)k0 * ft.nty + ((j0 - WL_CD_Y0) / WL_CD_TY)) * ft.ntx;
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const int x = i0 + q;
        if (x >= WL_CD_X0 && x <= g.n[0] - 3 && row[(x - WL_CD_X0) / WL_CD_TX]) done |= 1u << q;
      }
      if (done == (1u << V) - 1u) return;
    }
    auto store = [&](float* dst, const Pack<V>& a) {
      if (done == 0) { stp<V>(dst, a); return; }
#pragma unroll
      for (int q = 0; q < V; ++q) if (!((done >> q) & 1u)) dst[q] = a.v[q];
    };
    // one 32-bit load covers the 4 cells of a pack (mask + I is 4-byte aligned like the fields are 16-byte aligned)
    const bool plain = V == 4 ? (*reinterpret_cast<const unsigned int*>(mask + I) == 0u) : (mask[I] == 0);
    if (plain) {  // away from any body and boundary face: u = (u + f)·scale
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const int64_t Ii = I + i * g.comp;
        const Pack<V> fc = ldp<V>(f + Ii);
        Pack<V> un;
#pragma unroll
        for (int q = 0; q < V; ++q) un.v[q] = 0.f;
        if (accumulate) un = ldp<V>(ua + Ii);
#pragma unroll
        for (int q = 0; q < V; ++q) un.v[q] = (float)((accumulate ? (double)un.v[q] : 0.0) + (double)fc.v[q]) * scale;
        store(u + Ii, un);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int64_t Ii = I + i * g.comp;
      const Pack<V> fc = ldp<V>(f + Ii);
      Pack<V> s;
      {
        const Pack<V> m = ldp<V>(mu1 + I + (int64_t)i * g.comp), fm = ld_xm<V>(f + Ii, fc), fp = ld_xp<V>(f + Ii, fc);
#pragma unroll
        for (int q = 0; q < V; ++q) s.v[q] = 0.f + m.v[q] * (fp.v[q] - fm.v[q]);
      }
#pragma unroll
      for (int j = 1; j < D; ++j) {
        const Pack<V> m = ldp<V>(mu1 + I + (int64_t)(i + D * j) * g.comp), fm = ldp<V>(f + Ii - g.s[j]), fp = ldp<V>(f + Ii + g.s[j]);
#pragma unroll
        for (int q = 0; q < V; ++q) s.v[q] += m.v[q] * (fp.v[q] - fm.v[q]);
      }
      const Pack<V> vb = ldp<V>(Vb + Ii), m0 = ldp<V>(mu0 + Ii);
      Pack<V> un;
#pragma unroll
      for (int q = 0; q < V; ++q) un.v[q] = 0.f;
      if (accumulate) un = ldp<V>(ua + Ii);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const double add = (0.5 * (double)s.v[q] + (double)vb.v[q]) + (double)(m0.v[q] * fc.v[q]);
        un.v[q] = (float)((accumulate ? (double)un.v[q] : 0.0) + add) * scale;
      }
      store(u + Ii, un);
    }
  });
}
// ---- BDIM! stage 2 on what the fused conv_diff! left over (3-D) -----------------------------------
// one cell, all components: the expressions of k_bdim2
__device__ __forceinline__ void bdim2_cell(const GridD& g, float* __restrict__ u, const float* __restrict__ ua, const float* __restrict__ f,
                                           const float* __restrict__ Vb, const float* __restrict__ mu0, const float* __restrict__ mu1,
                                           const unsigned char* __restrict__ mask, int64_t I, int accumulate, float scale) {
  const bool plain = mask[I] == 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int64_t Ii = I + i * g.comp;
    const float fc = f[Ii];
    const double acc = accumulate ? (double)ua[Ii] : 0.0;
    if (plain) {
      u[Ii] = (float)(acc + (double)fc) * scale;
    } else {
      float s = 0.f + mu1[I + (int64_t)i * g.comp] * (f[Ii + 1] - f[Ii - 1]);
#pragma unroll
      for (int j = 1; j < 3; ++j) s += mu1[I + (int64_t)(i + 3 * j) * g.comp] * (f[Ii + g.s[j]] - f[Ii - g.s[j]]);
      const double add = (0.5 * (double)s + (double)Vb[Ii]) + (double)(mu0[Ii] * fc);
      u[Ii] = (float)(acc + add) * scale;
    }
  }
}
// the interior cells outside the core of the conv_diff! kernel (the enumeration of k_conv_diff_shell, ghost cells skipped)
__global__ void __launch_bounds__(256) k_bdim2_shell(GridD g, float* __restrict__ u, const float* __restrict__ ua, const float* __restrict__ f,
                                                     const float* __restrict__ Vb, const float* __restrict__ mu0, const float* __restrict__ mu1,
                                                     const unsigned char* __restrict__ mask, unsigned int nZ, unsigned int nY, unsigned int nX,
                                                     int xlow, int accumulate, float scale) {
  const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const int zc0 = g.zlo_halo ? 1 : 2, zc1 = g.zhi_halo ? n2 - 2 : n2 - 3;
  const unsigned int total = nZ + nY + nX;
  for (unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int c[3];
    if (idx < nZ) {
      const unsigned int row = idx / (unsigned int)n0, kk = row / (unsigned int)n1;
      c[0] = (int)(idx - row * n0);
      c[1] = (int)(row - kk * n1);
      c[2] = (int)kk < zc0 ? (int)kk : zc1 + 1 + ((int)kk - zc0);
    } else if (idx < nZ + nY) {
      const unsigned int t = idx - nZ, row = t / (unsigned int)n0, kk = row >> 2, jj = row & 3u;
      c[0] = (int)(t - row * n0);
      c[1] = jj < 2 ? (int)jj : n1 - 4 + (int)jj;
      c[2] = zc0 + (int)kk;
    } else {
      const unsigned int t = idx - nZ - nY, nxs = (unsigned int)xlow + 2u, row = t / nxs, ii = t - row * nxs, kk = row / (unsigned int)(n1 - 4);
      c[0] = (int)ii < xlow ? (int)ii : n0 - 2 + ((int)ii - xlow);
      c[1] = 2 + (int)(row - kk * (n1 - 4));
      c[2] = zc0 + (int)kk;
    }
    if (c[0] < 1 || c[0] > n0 - 2 || c[1] < 1 || c[1] > n1 - 2 || c[2] < 1 || c[2] > n2 - 2) continue;  // BDIM! stage 2 is @inside
    const int64_t I = c[0] + c[1] * g.s[1] + c[2] * g.s[2];
    if ((mask[I] & 2) == 0) continue;  // no μ₁ here: k_conv_diff_shell applied stage 2 itself
    bdim2_cell(g, u, ua, f, Vb, mu0, mu1, mask, I, accumulate, scale);
  }
}
// the core tile-planes that are not clean: one block per (tile, plane) unit, two cells per thread
__global__ void __launch_bounds__(256) k_bdim2_tiles(GridD g, float* __restrict__ u, const float* __restrict__ ua, const float* __restrict__ f,
                                                     const float* __restrict__ Vb, const float* __restrict__ mu0, const float* __restrict__ mu1,
                                                     const unsigned char* __restrict__ mask, FusedTiles ft, int accumulate, float scale) {
  // the flags of 256 consecutive units are read with one coalesced load; the (few) units that are not clean are
  // collected in shared memory and worked off by the whole block
  __shared__ int todo[256];
  __shared__ int ntodo;
  const int per_plane = ft.ntx * ft.nty;
  const int nunits = per_plane * (ft.zhi - ft.zlo + 1);
  const int base0 = per_plane * ft.zlo;  // the map starts at plane 0
  for (int chunk = blockIdx.x * 256; chunk < nunits; chunk += gridDim.x * 256) {
    if (threadIdx.x == 0) ntodo = 0;
    __syncthreads();
    const int mine = chunk + (int)threadIdx.x;
    if (mine < nunits && ft.clean[base0 + mine] == 0) todo[atomicAdd(&ntodo, 1)] = mine;
    __syncthreads();
    const int n = ntodo;
    for (int t = 0; t < n; ++t) {
      const int unit = todo[t];
      const int tx = unit % ft.ntx, q = unit / ft.ntx, ty = q % ft.nty, z = ft.zlo + q / ft.nty;
      const int y = WL_CD_Y0 + ty * WL_CD_TY + (int)(threadIdx.x >> 5);
      if (y > g.n[1] - 3) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = WL_CD_X0 + tx * WL_CD_TX + (int)(threadIdx.x & 31) + 32 * h;
        if (x <= g.n[0] - 3) bdim2_cell(g, u, ua, f, Vb, mu0, mu1, mask, x + y * g.s[1] + (int64_t)z * g.s[2], accumulate, scale);
      }
    }
    __syncthreads();
  }
}

template <int D, int V>
__global__ void __launch_bounds__(256) k_scale_u(GridD g, float* u, double scale) {
  for_rows<V>(g, [&](int, int, int, int64_t I) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      Pack<V> a = ldp<V>(u + I + i * g.comp);
#pragma unroll
      for (int q = 0; q < V; ++q) a.v[q] = (float)((double)a.v[q] * scale);
      stp<V>(u + I + i * g.comp, a);
    }
  });
}

// project!: z = div(u) (Flow.jl:139, :11-17)
template <int D, int V>
__global__ void __launch_bounds__(256) k_project_pre(GridD g, const float* __restrict__ u, float* __restrict__ z) {
  for_rows<V>(g, [&](int, int, int, int64_t I) { stp<V>(z + I, div_pack<D, V>(g, u, I)); });
}
// project!: x .*= dt (Flow.jl:139) / x ./= dt (:144) over the WHOLE allocation of x — interior, ghost
// cells and slab halo planes alike.  A ghost cell holds a copy of an interior value, and scaling a
// copy gives the bits of the scaled original, so the ghosts stay consistent and neither the perBC!
// of residual! (Poisson.jl:92) nor a halo exchange is needed afterwards.
// Inside mom_step! a project! leaves x scaled and records the pending x ./= dt (Sim::p_pend); the next scaling pass
// applies it first — the reference's roundings, in its order, in one pass instead of two: `pre` 1 divides by the
// Float32 pre32 (w = 1: dt is Float32), 2 by the Float64 pre64 (w = 0.5: dt = 0.5Δt is Float64), then `op` 1 / 2
// multiplies by f32v / f64v (0: nothing, i.e. only the pending division is materialised).
__global__ void __launch_bounds__(256) k_scale_alloc(float4* __restrict__ x, int64_t n4, int op, float f32v, double f64v, int pre, float pre32, double pre64) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[t];
    float* p = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (pre == 1) p[q] = p[q] / pre32;
      else if (pre == 2) p[q] = (float)((double)p[q] / pre64);
      if (op == 1) p[q] = p[q] * f32v;
      else if (op == 2) p[q] = (float)((double)p[q] * f64v);
    }
    x[t] = v;
  }
}
// project!: u[I,i] -= L[I,i]*∂(i,I,x) (Flow.jl:141-143)
// On a uniform pack (Lvl::umask) L[I,i] is the constant uL = 1 and is not loaded.
// `xdiv` (inside mom_step!, when the next solver! opens with the fused head of wl_resjac.cu): the closing x ./= dt of
// project! (Flow.jl:144) is stored for the interior cells into that spare array — one extra word written here instead of a
// read-modify-write pass of its own; x itself keeps the scaled values with the division pending (Sim::p_pend).
template <int V>
__device__ __forceinline__ void store_xdiv(float* __restrict__ xdiv, int64_t I, const Pack<V>& xc, int pre, float pre32, double pre64) {
  Pack<V> xd;
#pragma unroll
  for (int q = 0; q < V; ++q) xd.v[q] = pre == 1 ? xc.v[q] / pre32 : (float)((double)xc.v[q] / pre64);
  stp<V>(xdiv + I, xd);
}
template <int D, int V>
__global__ void __launch_bounds__(256) k_project_correct(Lvl l, float* __restrict__ u, float* __restrict__ xdiv, int pre, float pre32, double pre64) {
  const GridD& g = l.g;
  const float* __restrict__ L = l.L;
  const float* __restrict__ x = l.x;
  for_rows<V>(g, [&](int, int, int, int64_t I) {
    const Pack<V> xc = ldp<V>(x + I);
    if (xdiv) store_xdiv<V>(xdiv, I, xc, pre, pre32, pre64);
    const bool uni = pack_uniform<V>(l, I);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const Pack<V> xm = i == 0 ? ld_xm<V>(x + I, xc) : ldp<V>(x + I - g.s[i]);
      const Pack<V> Lc = uni ? splat<V>(l.uL) : ldp<V>(L + I + i * g.comp);
      Pack<V> uc = ldp<V>(u + I + i * g.comp);
#pragma unroll
      for (int q = 0; q < V; ++q) uc.v[q] -= Lc.v[q] * (xc.v[q] - xm.v[q]);
      stp<V>(u + I + i * g.comp, uc);
    }
  });
}
// The corrector's project! correction and CFL's σ = flux_out(I,u) (Flow.jl:141-143, 173, 176-182) in one pass: CFL used
// to re-read the three components the correction had just written (5 words per cell, 0.53 ms at 512³).  flux_out needs the
// corrected u_i on the UPPER faces too, i.e. the neighbours' correction: their operands (u, x, L one cell up) are read
// as well — which rules out the in-place update, so the corrected field goes to a spare velocity buffer and the caller
// swaps the pointers.  Cells whose upper neighbour is a ghost cell (any coordinate n-2) need the values BC! sets
// afterwards: they are left to k_cfl_upper and stay out of the maximum here.
__global__ void __launch_bounds__(256, WL_PCC_MINB) k_project_correct_cfl(Lvl l, const float* __restrict__ u, float* __restrict__ uo, float* __restrict__ sigma, ReduceWS ws,
                                                             FinArgs fa, float* __restrict__ xdiv, int pre, float pre32, double pre64) {
  const GridD& g = l.g;
  const float* __restrict__ L = l.L;
  const float* __restrict__ x = l.x;
  const int64_t s1 = g.s[1], s2 = g.s[2], cp = g.comp;
  double acc[1] = {0.0};
  for_rows<4>(g, [&](int i, int j, int k, int64_t I) {
    const Pack<4> xc = ldp<4>(x + I);
    if (xdiv) store_xdiv<4>(xdiv, I, xc, pre, pre32, pre64);
    const bool uni = pack_uniform<4>(l, I);
    const Pack<4> one = splat<4>(l.uL);
    // lower faces: the correction itself (k_project_correct)
    Pack<4> un[3];
    {
      const Pack<4> xm = ld_xm<4>(x + I, xc), Lc = uni ? one : ldp<4>(L + I), uc = ldp<4>(u + I);
#pragma unroll
      for (int q = 0; q < 4; ++q) un[0].v[q] = uc.v[q] - Lc.v[q] * (xc.v[q] - xm.v[q]);
    }
    {
      const Pack<4> xm = ldp<4>(x + I - s1), Lc = uni ? one : ldp<4>(L + I + cp), uc = ldp<4>(u + I + cp);
#pragma unroll
      for (int q = 0; q < 4; ++q) un[1].v[q] = uc.v[q] - Lc.v[q] * (xc.v[q] - xm.v[q]);
    }
    {
      const Pack<4> xm = ldp<4>(x + I - s2), Lc = uni ? one : ldp<4>(L + I + 2 * cp), uc = ldp<4>(u + I + 2 * cp);
#pragma unroll
      for (int q = 0; q < 4; ++q) un[2].v[q] = uc.v[q] - Lc.v[q] * (xc.v[q] - xm.v[q]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) stp<4>(uo + I + c * cp, un[c]);
    // upper faces: the corrected value of the neighbour cell, evaluated with its operands in its operation order
    Pack<4> up[3];
    {
      const float x4 = x[I + 4], u4 = u[I + 4], L4 = uni ? l.uL : L[I + 4];
#pragma unroll
      for (int q = 0; q < 3; ++q) up[0].v[q] = un[0].v[q + 1];
      up[0].v[3] = u4 - L4 * (x4 - xc.v[3]);
    }
    {
      const Pack<4> xp = ldp<4>(x + I + s1), Lp = uni ? one : ldp<4>(L + I + cp + s1), uq = ldp<4>(u + I + cp + s1);
#pragma unroll
      for (int q = 0; q < 4; ++q) up[1].v[q] = uq.v[q] - Lp.v[q] * (xp.v[q] - xc.v[q]);
    }
    {
      const Pack<4> xp = ldp<4>(x + I + s2), Lp = uni ? one : ldp<4>(L + I + 2 * cp + s2), uq = ldp<4>(u + I + 2 * cp + s2);
#pragma unroll
      for (int q = 0; q < 4; ++q) up[2].v[q] = uq.v[q] - Lp.v[q] * (xp.v[q] - xc.v[q]);
    }
    Pack<4> o;
    const bool edge = j == g.n[1] - 2 || k == g.n[2] - 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // max(0., u) on a Float32 u: the Float32 maximum converted is the same Float64 number
      double sg = 0.0 + ((double)fmaxf(0.f, up[0].v[q]) + (double)fmaxf(0.f, -un[0].v[q]));
      sg += (double)fmaxf(0.f, up[1].v[q]) + (double)fmaxf(0.f, -un[1].v[q]);
      sg += (double)fmaxf(0.f, up[2].v[q]) + (double)fmaxf(0.f, -un[2].v[q]);
      o.v[q] = (float)sg;
      if (!edge && i + q != g.n[0] - 2) acc[0] = fmax(acc[0], (double)o.v[q]);
    }
    stp<4>(sigma + I, o);
  });
  double tot[1];
  if (grid_reduce<1, true>(acc, ws, tot)) finish_reduction(FIN_CFL, fa, tot[0], 0.0);
}
// σ = flux_out(I,u) on the interior cells with a coordinate n-2 (their upper neighbour is a ghost cell, final after BC!),
// folded into the maximum of k_project_correct_cfl.  Edge cells are visited by more than one plane: same value.
__global__ void __launch_bounds__(256) k_cfl_upper(GridD g, const float* __restrict__ u, float* __restrict__ sigma, ReduceWS ws, FinArgs fa, int64_t shell) {
  const int m0 = g.n[0] - 2, m1 = g.n[1] - 2, m2 = g.n[2] - 2;
  const int cA = m1 * m2, cB = m0 * m2, cC = m0 * m1;
  double acc[1] = {0.0};
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < cA + cB + cC; idx += gridDim.x * blockDim.x) {
    int c0, c1, c2;
    if (idx < cA) { c0 = m0; c1 = 1 + idx % m1; c2 = 1 + idx / m1; }
    else if (idx < cA + cB) { const int t = idx - cA; c0 = 1 + t % m0; c1 = m1; c2 = 1 + t / m0; }
    else { const int t = idx - cA - cB; c0 = 1 + t % m0; c1 = 1 + t / m0; c2 = m2; }
    const int64_t I = c0 + c1 * g.s[1] + c2 * g.s[2];
    double sg = 0.0 + (fmax(0., (double)u[I + 1]) + fmax(0., (double)-u[I]));
#pragma unroll
    for (int d = 1; d < 3; ++d) sg += fmax(0., (double)u[I + g.s[d] + d * g.comp]) + fmax(0., (double)-u[I + d * g.comp]);
    const float o = (float)sg;
    sigma[I] = o;
    acc[0] = fmax(acc[0], (double)o);
  }
  // … and the ghost shell of σ, which the reference's maximum over the padded array includes (k_shell_max): one reduction for both
  for_global_ghosts(g, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x,
                    [&](const int (&c)[3]) { acc[0] = fmax(acc[0], (double)sigma[c[0] + c[1] * g.s[1] + c[2] * g.s[2]]); });
  double tot[1];
  if (grid_reduce<1, true>(acc, ws, tot)) finish_reduction(FIN_CFL_SHELL, fa, tot[0], 0.0);
}
// CFL: σ = flux_out (Flow.jl:173,176-182) + interior max
template <int D, int V>
__global__ void __launch_bounds__(256) k_cfl(GridD g, const float* __restrict__ u, float* __restrict__ sigma, ReduceWS ws, FinArgs fa) {
  double acc[1] = {0.0};
  for_rows<V>(g, [&](int, int, int, int64_t I) {
    double s[V];
    {
      const Pack<V> uc = ldp<V>(u + I), up = ld_xp<V>(u + I, uc);
#pragma unroll
      for (int q = 0; q < V; ++q) s[q] = 0.0 + (fmax(0., (double)up.v[q]) + fmax(0., (double)-uc.v[q]));
    }
#pragma unroll
    for (int i = 1; i < D; ++i) {
      const Pack<V> uc = ldp<V>(u + I + i * g.comp), up = ldp<V>(u + I + g.s[i] + i * g.comp);
#pragma unroll
      for (int q = 0; q < V; ++q) s[q] += fmax(0., (double)up.v[q]) + fmax(0., (double)-uc.v[q]);
    }
    Pack<V> o;
#pragma unroll
    for (int q = 0; q < V; ++q) { o.v[q] = (float)s[q]; acc[0] = fmax(acc[0], (double)o.v[q]); }
    stp<V>(sigma + I, o);
  });
  double tot[1];
  if (grid_reduce<1, true>(acc, ws, tot)) finish_reduction(FIN_CFL, fa, tot[0], 0.0);
}
// the reference takes maximum(σ) over the PADDED array (Flow.jl:174): fold in the ghost shell, which
// holds the corrector's stale Φ.  Shell = cells with any coordinate on a boundary.
__global__ void __launch_bounds__(256) k_shell_max(GridD g, const float* __restrict__ sigma, ReduceWS ws, FinArgs fa, int64_t count) {
  double acc[1] = {-1e300};
  for_global_ghosts(g, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x,
                    [&](const int (&c)[3]) { acc[0] = fmax(acc[0], (double)sigma[c[0] + c[1] * g.s[1] + c[2] * g.s[2]]); });
  double tot[1];
  if (grid_reduce<1, true>(acc, ws, tot)) finish_reduction(FIN_CFL_SHELL, fa, tot[0], 0.0);
}
// with exitBC the exit plane of u₁ survives from step to step (saveexit); after the u/u⁰ pointer swap
// it has to be carried over explicitly
__global__ void __launch_bounds__(256) k_copy_exit_plane(GridD g, float* __restrict__ u, const float* __restrict__ u0) {
  const int ny = g.n[1], nz = g.n[2];
  const int64_t count = (int64_t)ny * nz;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = (g.n[0] - 1) + (idx % ny) * g.s[1] + (idx / ny) * g.s[2];
    u[I] = u0[I];
  }
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

template <int TX, int TY, int MINB>
static void launch_march_cfg(Sim& s, const CDArgs& a, int zchunk, bool corr) {
  const GridD& g = s.g;
  const int nzc = g.n[2] - 4 + (g.zlo_halo ? 1 : 0) + (g.zhi_halo ? 1 : 0);
  const int ntx = (g.n[0] - 4 + TX - 1) / TX, nty = (g.n[1] - 4 + TY - 1) / TY, ntz = (nzc + zchunk - 1) / zchunk;
  const dim3 blk(TX, TY, 1);
  const int nblk = ntx * nty * ntz;
  if (a.has_g) {
    if (corr) k_conv_diff3_march<TX, TY, MINB, true, true><<<nblk, blk, 0, s.stream>>>(a, zchunk, ntx, nty);
    else k_conv_diff3_march<TX, TY, MINB, true, false><<<nblk, blk, 0, s.stream>>>(a, zchunk, ntx, nty);
  } else {
    if (corr) k_conv_diff3_march<TX, TY, MINB, false, true><<<nblk, blk, 0, s.stream>>>(a, zchunk, ntx, nty);
    else k_conv_diff3_march<TX, TY, MINB, false, false><<<nblk, blk, 0, s.stream>>>(a, zchunk, ntx, nty);
  }
  ++s.launches;
}
static void launch_march(Sim& s, const CDArgs& a, int variant, int zchunk, bool corr) {
  switch (variant) {
    case 1: launch_march_cfg<128, 4, 1>(s, a, zchunk, corr); break;
    case 2: launch_march_cfg<64, 4, 3>(s, a, zchunk, corr); break;
    case 3: launch_march_cfg<128, 2, 3>(s, a, zchunk, corr); break;
    case 4: launch_march_cfg<64, 4, 4>(s, a, zchunk, corr); break;
    case 5: launch_march_cfg<32, 8, 3>(s, a, zchunk, corr); break;
    default: launch_march_cfg<128, 4, 2>(s, a, zchunk, corr); break;
  }
}

// the body-band mask and, from it, the clean map of the core tiles (when one of V, μ₀, μ₁ changed)
static int refresh_masks(Sim& s) {
  if (!s.mask_dirty) return 0;
  RLAUNCH(s, k_body_mask, s.g, s.g, s.V, s.mu0, s.mu1, s.body_mask);
  int ntx = 0, nty = 0;
  if (s.D == 3 && s.bdim_fuse && convdiff_tma_tiles(s, &ntx, &nty)) {
    const GridD& g = s.g;
    const dim3 blk = wl_pick_block(g);
    Box b = box_interior(g);
    b.lo[2] = b.hi[2] = 0;
    if (g.zlo_halo) LAUNCH(s, k_body_mask_plane<3>, box_tiles(b, blk), blk, g, s.V, s.mu0, s.mu1, s.body_mask, 0);
    if (g.zhi_halo) LAUNCH(s, k_body_mask_plane<3>, box_tiles(b, blk), blk, g, s.V, s.mu0, s.mu1, s.body_mask, g.n[2] - 1);
    const int zlo = g.zlo_halo ? 1 : 2, zhi = g.zhi_halo ? g.n[2] - 2 : g.n[2] - 3;
    const int64_t nunits = (int64_t)ntx * nty * g.n[2];
    k_tile_clean<<<(int)std::min<int64_t>((nunits + 7) / 8, (int64_t)s.nsm * 16), 256, 0, s.stream>>>(g, s.body_mask, s.tile_clean, ntx, nty, zlo, zhi);
    ++s.launches;
  }
  s.mask_dirty = false;
  return wl_cuda_check(s, cudaGetLastError(), "k_body_mask");
}

// bdim2_mode (3-D core kernel, with fuse_bdim1): 1 = the predictor's BDIM! stage 2 (u = f·1), 2 = the corrector's
// (w = (u + f)·½) is folded into the store on clean tile-planes; 0 = f is written everywhere
int mom_conv_diff(Sim& s, const float* src, bool fuse_bdim1, const double* gacc, int bdim2_mode, float dt_override) {
  CDArgs a;
  a.g = s.g; a.u = src; a.u0 = s.u0; a.V = s.V; a.f = s.f; a.sigma = s.sigma;
  a.nu = s.nu; a.dt = dt_override > 0.f ? dt_override : s.dt.back();
  a.ufused = nullptr; a.clean = nullptr; a.bscale = 1.f; a.mask = s.body_mask; a.mu0 = s.mu0; a.uacc = nullptr;
  {
    int ntx, nty;
    if (bdim2_mode && fuse_bdim1 && s.D == 3 && s.bdim_fuse && convdiff_tma_tiles(s, &ntx, &nty)) {
      const int rc0 = refresh_masks(s);
      if (rc0) return rc0;
      a.ufused = bdim2_mode == 1 ? s.u : s.w;
      a.clean = s.tile_clean;
      a.bscale = bdim2_mode == 1 ? 1.f : 0.5f;
      a.uacc = bdim2_mode == 1 ? nullptr : s.u;
    }
  }
  a.has_g = gacc ? 1 : 0;
  for (int k = 0; k < 3; ++k) a.gacc[k] = (gacc && k < s.D) ? gacc[k] : 0.0;
  a.fuse_bdim1 = fuse_bdim1 ? 1 : 0;
  wl_prof_begin(s, WL_PROF_CONVDIFF);
  const GridD& g = s.g;
  if (s.D == 3 && g.n[0] >= 8 && g.n[1] >= 8 && g.n[2] >= 8) {
    // core [2,n-3]³ by the marching kernel, boundary shell of thickness 2 by the generic one
    const bool corr = src != s.u0;
    s.last_cd_fused = false;
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("WL_CD_VARIANT"); variant = e ? atoi(e) : 4; }
    static int zchunk = 0;
    if (!zchunk) { const char* e = getenv("WL_CD_ZCHUNK"); zchunk = e ? atoi(e) : 32; }
    static int zchunk_tma = 0;
    if (!zchunk_tma) { const char* e = getenv("WL_CD_TMA_ZCHUNK"); zchunk_tma = std::min(128, std::max(4, e ? atoi(e) : 64)); }
    const bool used_tma = convdiff_tma_launch(s, a, zchunk_tma, corr);
    if (!used_tma) { a.ufused = nullptr; launch_march(s, a, variant, zchunk, corr); }
    s.last_cd_fused = a.ufused != nullptr;
    // the profiled "dominant kernel" is the core march alone: algorithmic words/cell = read u (D) [+ u⁰, V (2D)], write f (D)
    // with stage 2 folded in (clean tile-planes): read u [+ u⁰], write u — neither V nor f
    wl_prof_end(s, WL_PROF_CONVDIFF, 4.0 * (double)s.g.interior * (fuse_bdim1 ? ((src == s.u0 ? 3 : 4) - (a.ufused ? 1 : 0)) * s.D : 2 * s.D));
    {
      const int zc0 = g.zlo_halo ? 1 : 2, zc1 = g.zhi_halo ? g.n[2] - 2 : g.n[2] - 3, nzc = zc1 - zc0 + 1;
      const unsigned int nZ = (unsigned int)g.n[0] * g.n[1] * (unsigned int)(g.n[2] - nzc);
      const unsigned int nY = (unsigned int)g.n[0] * 4u * (unsigned int)nzc;
      const int xlow = used_tma ? convdiff_tma_core_x0(s) : 2;
      const unsigned int nX = (unsigned int)(xlow + 2) * (unsigned int)(g.n[1] - 4) * (unsigned int)nzc;
      const int64_t cnt = (int64_t)nZ + nY + nX;
      static int bps = 0;
      if (!bps) { const char* e = getenv("WL_SHELL_BPS"); bps = std::max(1, e ? atoi(e) : 16); }
      k_conv_diff_shell<3><<<(int)std::min<int64_t>((cnt + 255) / 256, (int64_t)s.nsm * bps), 256, 0, s.stream>>>(a, nZ, nY, nX, xlow);
      ++s.launches;
    }
  } else {
    s.last_cd_fused = false;
    const dim3 blk = wl_pick_block(s.g);
    LAUNCH_D(s, k_conv_diff, s.D, box_tiles(box_all(s.g), blk), blk, a);
    wl_prof_end(s, WL_PROF_CONVDIFF, 4.0 * (double)s.g.interior * (fuse_bdim1 ? (src == s.u0 ? 3 : 4) * s.D : 2 * s.D));
  }
  int rc = wl_cuda_check(s, cudaGetLastError(), "k_conv_diff");
  if (rc) return rc;
  return dist_halo(s, s.g, s.f, s.D, 1);  // BDIM! stage 2 reads f[I±δz] across slab faces
}
int mom_accelerate(Sim& s, const double* g) {
  if (!g) return 0;
  const dim3 blk = wl_pick_block(s.g);
  LAUNCH(s, k_accelerate, box_tiles(box_all(s.g), blk), blk, s.g, s.f, g[0], s.D > 1 ? g[1] : 0.0, s.D > 2 ? g[2] : 0.0);
  return wl_cuda_check(s, cudaGetLastError(), "k_accelerate");
}
int mom_bdim1(Sim& s) {
  const dim3 blk = wl_pick_block(s.g);
  LAUNCH(s, k_bdim1, box_tiles(box_all(s.g), blk), blk, s.g, s.f, s.u0, s.V, s.dt.back());
  return wl_cuda_check(s, cudaGetLastError(), "k_bdim1");
}
// `fused`: called by mom_step! after a conv_diff! that may have folded stage 2 into its store on the clean
// tile-planes: those cells are skipped, and the corrector's result goes to the spare buffer w (the kernel that
// wrote the clean cells could not update u in place while other blocks still read it)
int mom_bdim2(Sim& s, bool accumulate, float scale, bool fused) {
  int rc = refresh_masks(s);
  if (rc) return rc;
  FusedTiles ft{};
  float* dst = s.u;
  if (fused && s.last_cd_fused) {
    ft.clean = s.tile_clean;
    convdiff_tma_tiles(s, &ft.ntx, &ft.nty);
    ft.zlo = s.g.zlo_halo ? 1 : 2; ft.zhi = s.g.zhi_halo ? s.g.n[2] - 2 : s.g.n[2] - 3;
    if (accumulate) dst = s.w;
  }
  wl_prof_begin(s, WL_PROF_BDIM2);
  if (ft.clean) {
    // what conv_diff! did not finish: the interior cells of its boundary shell, and the core tile-planes near a body
    const GridD& g = s.g;
    const int nzc = ft.zhi - ft.zlo + 1, xlow = convdiff_tma_core_x0(s);
    const unsigned int nZ = (unsigned int)g.n[0] * g.n[1] * (unsigned int)(g.n[2] - nzc);
    const unsigned int nY = (unsigned int)g.n[0] * 4u * (unsigned int)nzc;
    const unsigned int nX = (unsigned int)(xlow + 2) * (unsigned int)(g.n[1] - 4) * (unsigned int)nzc;
    const int64_t cnt = (int64_t)nZ + nY + nX;
    k_bdim2_shell<<<(int)std::min<int64_t>((cnt + 255) / 256, s.nsm * 16), 256, 0, s.stream>>>(g, dst, s.u, s.f, s.V, s.mu0, s.mu1, s.body_mask, nZ, nY, nX,
                                                                                         xlow, accumulate ? 1 : 0, scale);
    const int64_t nunits = (int64_t)ft.ntx * ft.nty * nzc;
    k_bdim2_tiles<<<(int)std::min<int64_t>((nunits + 255) / 256, (int64_t)s.nsm * 8), 256, 0, s.stream>>>(g, dst, s.u, s.f, s.V, s.mu0, s.mu1, s.body_mask, ft,
                                                                                           accumulate ? 1 : 0, scale);
    s.launches += 2;
  } else {
    RLAUNCH(s, k_bdim2, s.g, s.g, dst, s.u, s.f, s.V, s.mu0, s.mu1, s.body_mask, ft, accumulate ? 1 : 0, scale);
  }
  // algorithmic traffic away from a body: read f (D) [+ u (D)] + mask (¼), write u (D); nothing but the tile flag on fused cells
  wl_prof_end(s, WL_PROF_BDIM2, 4.0 * (double)s.g.interior * (ft.clean ? 0.02 : (accumulate ? 3 : 2) * s.D + 0.25));
  return wl_cuda_check(s, cudaGetLastError(), "k_bdim2");
}
// flow.f as the reference leaves it after mom_step! (the corrector's BDIM! stage-1 f on every cell): after a fused
// step the clean tiles never received it, so it is recomputed here, on demand, from the inputs the corrector used
int mom_materialize_f(Sim& s) {
  if (!s.f_stale) return 0;
  s.f_stale = false;
  return mom_conv_diff(s, s.w, true, s.f_has_g ? s.f_g : nullptr, 0, s.f_dt);
}
int mom_scale_u(Sim& s, double scale) {
  RLAUNCH(s, k_scale_u, s.g, s.g, s.u, scale);
  return wl_cuda_check(s, cudaGetLastError(), "k_scale_u");
}
static int scale_pressure(Sim& s, int op, float f32v, double f64v) {
  Lvl& b = s.levels[0];
  float4* xall = reinterpret_cast<float4*>(b.x - s.g.lead);
  const int64_t n4 = s.g.comp / 4;
  const int sgrid = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)s.nsm * 8);
  k_scale_alloc<<<sgrid, 256, 0, s.stream>>>(xall, n4, op, f32v, f64v, s.p_pend, s.p_pend32, s.p_pend64);
  ++s.launches;
  s.p_pend = 0;
  return wl_cuda_check(s, cudaGetLastError(), "k_scale_alloc");
}
// the pressure as the reference leaves it after project! (x ./= dt applied): on demand, before anything outside
// mom_step! looks at it
int mom_materialize_p(Sim& s) {
  s.xdiv_valid = false;  // every entry point that may touch x or ϵ outside mom_step! comes through here first
  if (!s.p_pend) return 0;
  return scale_pressure(s, 0, 0.f, 0.0);
}
// project!(a,b,w) (Flow.jl:137-145)
// `in_step` (mom_step!): z = div u is evaluated inside residual! instead of by a pass of its own, and the closing
// x ./= dt is left pending for the next project! (or mom_materialize_p)
int mom_project(Sim& s, double w, int* n, int slot, bool defer_count, bool in_step, bool with_cfl) {
  const float dt32 = s.dt.back();
  // dt = w*Δt: Float32 for the Int w=1, Float64 for w=0.5 — the reference then evaluates Float32(Float64(x)·dt) and
  // Float32(Float64(x)/dt).  When that Float64 dt is itself a Float32 number (always for w = 0.5: halving is exact) these
  // equal the Float32 operations x·dt and x/dt bit for bit: the product of two Float32 numbers is exact in Float64, and a
  // Float64 quotient rounded to Float32 is the correctly rounded Float32 quotient (double rounding is innocuous for
  // division when the wider format has ≥ 2·24+2 bits).  So only a dt that is not a Float32 number takes the Float64 path.
  static const bool exact32 = [] { const char* e = getenv("WL_SCALE_F64"); return !(e && atoi(e) != 0); }();
  const double dtw_ref = (w == 1.0) ? (double)dt32 : w * (double)dt32;
  const bool f32 = (w == 1.0) || (exact32 && (double)(float)dtw_ref == dtw_ref && fabs(dtw_ref) > 1e-30);
  const double dtw = dtw_ref;
  Lvl& b = s.levels[0];
  int rc;
  // inside mom_step! with the fused solver head: the pressure correction of the previous project! left x ./ dt of the interior
  // cells in the ϵ array (xdiv_valid) and the head applies x .*= dt while it stages x from there — no scaling pass
  const bool head = in_step && s.xdiv_on && pois_head_fusable(s);
  const bool from_div = head && s.xdiv_valid;
  s.xdiv_valid = false;
  wl_prof_begin(s, WL_PROF_PROJECT);
  if (!in_step) RLAUNCH(s, k_project_pre, s.g, s.g, s.u, b.z);
  if (from_div) { s.xmul_op = f32 ? 1 : 2; s.xmul32 = (float)dtw; s.xmul64 = dtw; }
  else if ((rc = scale_pressure(s, f32 ? 1 : 2, (float)dtw, dtw))) return rc;
  wl_prof_end(s, WL_PROF_PROJECT, 4.0 * (double)s.g.interior * ((in_step ? 0 : s.D + 1) + (from_div ? 0 : 2)));
  rc = pois_solve(s, 1e-4, s.pois_kind == WL_POIS_MULTILEVEL ? 32 : 1000, n, true, slot, defer_count, in_step ? s.u : nullptr);
  if (rc) return rc;
  wl_prof_begin(s, WL_PROF_PROJECT);
  // the corrector's correction also evaluates CFL's σ (k_project_correct_cfl) when a velocity-sized buffer is free to take
  // the corrected field: after a fused step flow.f is not materialised (f_stale), its array is that buffer
  const RowLaunch rlc = row_launch(s.g);
  float* xdiv = head ? b.eps : nullptr;  // ϵ is free between two solver! calls
  s.cfl_fused = false;
  if (with_cfl && s.cflfuse_on && s.D == 3 && rlc.V == 4 && s.f_stale) {
    const FinArgs fa = fin_misc(s);
    k_project_correct_cfl<<<wl_grid_for(s, rlc.nunits), rlc.blk, 0, s.stream>>>(b, s.u, s.f, s.sigma, s.ws, fa, xdiv, f32 ? 1 : 2, (float)dtw, dtw);
    ++s.launches;
    if ((rc = post_reduce(s, FIN_CFL, fa, true))) return rc;
    if (s.exitBC) {  // BC! rewrites every ghost cell of the new array except the saved exit plane
      const int64_t cnt = (int64_t)s.g.n[1] * s.g.n[2];
      k_copy_exit_plane<<<(int)std::min<int64_t>((cnt + 255) / 256, 256), 256, 0, s.stream>>>(s.g, s.f, s.u);
      ++s.launches;
    }
    float* t = s.u; s.u = s.f; s.f = t;
    s.cfl_fused = true;
  } else {
    RLAUNCH(s, k_project_correct, s.g, b, s.u, xdiv, f32 ? 1 : 2, (float)dtw, dtw);
  }
  s.p_pend = f32 ? 1 : 2; s.p_pend32 = (float)dtw; s.p_pend64 = dtw;
  s.xdiv_valid = xdiv != nullptr;
  if (!in_step && (rc = mom_materialize_p(s))) return rc;
  wl_prof_end(s, WL_PROF_PROJECT, 4.0 * (double)s.g.interior * (wl_words(s, 3 * s.D + 1, s.D) + (in_step ? 0 : 2) + (xdiv ? 1 : 0)));
  return wl_cuda_check(s, cudaGetLastError(), "project");
}
int mom_div(Sim& s, const float* u, float* z) {
  RLAUNCH(s, k_project_pre, s.g, s.g, u, z);
  return wl_cuda_check(s, cudaGetLastError(), "k_project_pre");
}
// CFL(a) (Flow.jl:172-175)
int mom_cfl(Sim& s, double* dt) {
  const FinArgs fa = fin_misc(s);
  int rc;
  if (s.cfl_fused) {
    // σ and its maximum away from the upper boundaries came with the pressure correction; the rest needs BC!'s ghosts
    s.cfl_fused = false;
    const int64_t cnt = (int64_t)(s.g.n[1] - 2) * (s.g.n[2] - 2) + (int64_t)(s.g.n[0] - 2) * (s.g.n[2] - 2) + (int64_t)(s.g.n[0] - 2) * (s.g.n[1] - 2);
    const int64_t shell = global_ghost_count(s.g);
    k_cfl_upper<<<(int)std::min<int64_t>((std::max(cnt, shell) + 255) / 256, (int64_t)s.nsm * 8), 256, 0, s.stream>>>(s.g, s.u, s.sigma, s.ws, fa, shell);
    ++s.launches;
    if ((rc = post_reduce(s, FIN_CFL_SHELL, fa, true))) return rc;
  } else {
    RLAUNCH(s, k_cfl, s.g, s.g, s.u, s.sigma, s.ws, fa);
    if ((rc = post_reduce(s, FIN_CFL, fa, true))) return rc;
    const int64_t cnt = global_ghost_count(s.g);
    k_shell_max<<<(int)std::min<int64_t>((cnt + 255) / 256, (int64_t)s.nsm * 4), 256, 0, s.stream>>>(s.g, s.sigma, s.ws, fa, cnt);
    ++s.launches;
    if ((rc = post_reduce(s, FIN_CFL_SHELL, fa, true))) return rc;
  }
  // one readback for the step: CFL's maximum and the V-cycle counts of the two solver! kernels
  if ((rc = wl_read_scalars(s, s.d_misc, sizeof(MiscScalars)))) return rc;
  MiscScalars hm;
  memcpy(&hm, s.h_pinned, sizeof(hm));
  s.last_solve_n[0] = hm.solve_n[0];
  s.last_solve_n[1] = hm.solve_n[1];
  const float m = hm.cfl_max + 5 * s.nu;
  const float v = 1.f / m;
  *dt = (double)(v < 10.f ? v : 10.f);
  return 0;
}
// mom_step!(a,b) (Flow.jl:153-169)
int mom_step(Sim& s, const double* U, const double* gp, const double* gc, double* dt_next, int* n_pred, int* n_corr) {
  int rc;
  wl_prof_begin(s, WL_PROF_STEP);
  // a.u⁰ .= a.u; scale_u!(a,0): pointer swap; every cell of the new u is rewritten below before it is read
  float* t = s.u; s.u = s.u0; s.u0 = t;
  if (s.exitBC) {
    const int64_t cnt = (int64_t)s.g.n[1] * s.g.n[2];
    k_copy_exit_plane<<<(int)std::min<int64_t>((cnt + 255) / 256, 256), 256, 0, s.stream>>>(s.g, s.u, s.u0);
    ++s.launches;
  }
  // predictor u → u'
  s.f_stale = false;
  if ((rc = mom_conv_diff(s, s.u0, true, gp, 1))) return rc;
  if ((rc = mom_bdim2(s, false, 1.f, true))) return rc;
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  if (s.exitBC) {
    if ((rc = bc_exit(s, U, s.dt.back()))) return rc;
    if ((rc = dist_halo(s, s.g, s.u, 1, 2))) return rc;  // the exit plane changed after BC!'s exchange
  }
  const bool defer = s.vc_solve && s.pois_kind == WL_POIS_MULTILEVEL;
  // the pressure is not observable between the two project! calls of a step: the predictor leaves it scaled
  // and the corrector's first pass does x = (x / Δt) * (Δt/2) with the reference's two roundings
  if ((rc = mom_project(s, 1.0, n_pred, 0, defer, true))) return rc;
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  // corrector u → u¹
  if ((rc = mom_conv_diff(s, s.u, true, gc, 2))) return rc;
  if ((rc = mom_bdim2(s, true, 0.5f, true))) return rc;
  if (s.last_cd_fused) {
    // the new u was assembled in w (clean tile-planes by conv_diff!, the rest by BDIM! stage 2): u ↔ w.  BC! below rewrites
    // every ghost cell of it except the saved exit plane, which is carried over; w keeps u' — with u⁰, V and Δt that is all
    // an on-demand flow.f needs (mom_materialize_f)
    if (s.exitBC) {
      const int64_t cnt = (int64_t)s.g.n[1] * s.g.n[2];
      k_copy_exit_plane<<<(int)std::min<int64_t>((cnt + 255) / 256, 256), 256, 0, s.stream>>>(s.g, s.w, s.u);
      ++s.launches;
    }
    float* t2 = s.u; s.u = s.w; s.w = t2;
    s.f_stale = true;
    s.f_dt = s.dt.back();
    s.f_has_g = gc != nullptr;
    for (int k = 0; k < 3; ++k) s.f_g[k] = (gc && k < s.D) ? gc[k] : 0.0;
  }
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  if ((rc = mom_project(s, 0.5, n_corr, 1, defer, true, true))) return rc;
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  double dt;
  if ((rc = mom_cfl(s, &dt))) return rc;
  if (defer) {
    for (int k = 0; k < 2; ++k) s.n_log.push_back(s.last_solve_n[k]);
    if (n_pred) *n_pred = s.last_solve_n[0];
    if (n_corr) *n_corr = s.last_solve_n[1];
  }
  wl_prof_end(s, WL_PROF_STEP, 0.0);
  if ((rc = dist_check_p2p_error(s))) return rc;
  s.dt.push_back((float)dt);
  if (dt_next) *dt_next = dt;
  return 0;
}
