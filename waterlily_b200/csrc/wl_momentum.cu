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
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float v = r[i];
    if (a.has_g) v = (float)((double)v + a.gacc[i]);          // accelerate! (Flow.jl:68-70)
    if (a.fuse_bdim1) v = a.u0[I + i * g.comp] + a.dt * v - a.V[I + i * g.comp];  // BDIM! stage 1 (:133)
    a.f[I + i * g.comp] = v;
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
template <int D>
__global__ void __launch_bounds__(256) k_conv_diff_shell(CDArgs a, unsigned int nZ, unsigned int nY, unsigned int nX, int xlow) {
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
// component — everywhere except within 2+ϵ cells of a body and on the domain-boundary faces), 1
// elsewhere.  One byte per cell replaces 15 words of μ₁/V/μ₀ reads; rebuilt whenever one of the three
// fields changes (upload, BC!).  The result is bit-identical: with those values the general
// expression reduces to (Float64)u + (Float64)f exactly.
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
        for (int q = 0; q < V; ++q) m[q] |= (m1.v[q] != 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < V; ++q) mask[I + q] = m[q];
  });
}

// BDIM! stage 2 (Flow.jl:134) with μddn (:18-24); `accumulate`=0 folds the predictor's
// scale_u!(a,0), `scale` folds the corrector's scale_u!(a,0.5)
template <int D, int V>
__global__ void __launch_bounds__(256) k_bdim2(GridD g, float* __restrict__ u, const float* __restrict__ f, const float* __restrict__ Vb,
                                               const float* __restrict__ mu0, const float* __restrict__ mu1,
                                               const unsigned char* __restrict__ mask, int accumulate, float scale) {
  for_rows<V>(g, [&](int, int, int, int64_t I) {
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
        if (accumulate) un = ldp<V>(u + Ii);
#pragma unroll
        for (int q = 0; q < V; ++q) un.v[q] = (float)((accumulate ? (double)un.v[q] : 0.0) + (double)fc.v[q]) * scale;
        stp<V>(u + Ii, un);
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
      if (accumulate) un = ldp<V>(u + Ii);
#pragma unroll
      for (int q = 0; q < V; ++q) {
        const double add = (0.5 * (double)s.v[q] + (double)vb.v[q]) + (double)(m0.v[q] * fc.v[q]);
        un.v[q] = (float)((accumulate ? (double)un.v[q] : 0.0) + add) * scale;
      }
      stp<V>(u + Ii, un);
    }
  });
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
  for_rows<V>(g, [&](int, int, int, int64_t I) {
    Pack<V> s;
    {
      const Pack<V> uc = ldp<V>(u + I), up = ld_xp<V>(u + I, uc);
#pragma unroll
      for (int q = 0; q < V; ++q) s.v[q] = 0.f + (up.v[q] - uc.v[q]);
    }
#pragma unroll
    for (int i = 1; i < D; ++i) {
      const Pack<V> uc = ldp<V>(u + I + i * g.comp), up = ldp<V>(u + I + g.s[i] + i * g.comp);
#pragma unroll
      for (int q = 0; q < V; ++q) s.v[q] += up.v[q] - uc.v[q];
    }
    stp<V>(z + I, s);
  });
}
// project!: x .*= dt (Flow.jl:139) / x ./= dt (:144) over the WHOLE allocation of x — interior, ghost
// cells and slab halo planes alike.  A ghost cell holds a copy of an interior value, and scaling a
// copy gives the bits of the scaled original, so the ghosts stay consistent and neither the perBC!
// of residual! (Poisson.jl:92) nor a halo exchange is needed afterwards.
// `prediv` > 0: first undo the PREVIOUS project!'s scaling (x ./= dt with its Float32 dt, w = 1) — mom_step!
// merges the predictor's x ./= dt with the corrector's x .*= dt (two roundings in one pass, same bits)
__global__ void __launch_bounds__(256) k_scale_alloc(float4* __restrict__ x, int64_t n4, float f32v, double f64v, int f32, int divide, float prediv) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[t];
    float* p = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (prediv > 0.f) p[q] = p[q] / prediv;
      if (divide) p[q] = f32 ? p[q] / f32v : (float)((double)p[q] / f64v);
      else p[q] = f32 ? p[q] * f32v : (float)((double)p[q] * f64v);
    }
    x[t] = v;
  }
}
// project!: u[I,i] -= L[I,i]*∂(i,I,x) (Flow.jl:141-143)
// On a uniform pack (Lvl::umask) L[I,i] is the constant uL = 1 and is not loaded.
template <int D, int V>
__global__ void __launch_bounds__(256) k_project_correct(Lvl l, float* __restrict__ u) {
  const GridD& g = l.g;
  const float* __restrict__ L = l.L;
  const float* __restrict__ x = l.x;
  for_rows<V>(g, [&](int, int, int, int64_t I) {
    const Pack<V> xc = ldp<V>(x + I);
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
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    int c[3];
    if (shell_decode(g, idx, c) && is_global_ghost(g, c)) acc[0] = fmax(acc[0], (double)sigma[c[0] + c[1] * g.s[1] + c[2] * g.s[2]]);
  }
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

int mom_conv_diff(Sim& s, const float* src, bool fuse_bdim1, const double* gacc) {
  CDArgs a;
  a.g = s.g; a.u = src; a.u0 = s.u0; a.V = s.V; a.f = s.f; a.sigma = s.sigma;
  a.nu = s.nu; a.dt = s.dt.back();
  a.has_g = gacc ? 1 : 0;
  for (int k = 0; k < 3; ++k) a.gacc[k] = (gacc && k < s.D) ? gacc[k] : 0.0;
  a.fuse_bdim1 = fuse_bdim1 ? 1 : 0;
  wl_prof_begin(s, WL_PROF_CONVDIFF);
  const GridD& g = s.g;
  if (s.D == 3 && g.n[0] >= 8 && g.n[1] >= 8 && g.n[2] >= 8) {
    // core [2,n-3]³ by the marching kernel, boundary shell of thickness 2 by the generic one
    const bool corr = src != s.u0;
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("WL_CD_VARIANT"); variant = e ? atoi(e) : 4; }
    static int zchunk = 0;
    if (!zchunk) { const char* e = getenv("WL_CD_ZCHUNK"); zchunk = e ? atoi(e) : 32; }
    static int zchunk_tma = 0;
    if (!zchunk_tma) { const char* e = getenv("WL_CD_TMA_ZCHUNK"); zchunk_tma = e ? atoi(e) : 64; }
    const bool used_tma = convdiff_tma_launch(s, a, zchunk_tma, corr);
    if (!used_tma) launch_march(s, a, variant, zchunk, corr);
    // the profiled "dominant kernel" is the core march alone: algorithmic words/cell = read u (D) [+ u⁰, V (2D)], write f (D)
    wl_prof_end(s, WL_PROF_CONVDIFF, 4.0 * (double)s.g.interior * (fuse_bdim1 ? (src == s.u0 ? 3 : 4) * s.D : 2 * s.D));
    {
      const int zc0 = g.zlo_halo ? 1 : 2, zc1 = g.zhi_halo ? g.n[2] - 2 : g.n[2] - 3, nzc = zc1 - zc0 + 1;
      const unsigned int nZ = (unsigned int)g.n[0] * g.n[1] * (unsigned int)(g.n[2] - nzc);
      const unsigned int nY = (unsigned int)g.n[0] * 4u * (unsigned int)nzc;
      const int xlow = used_tma ? convdiff_tma_core_x0(s) : 2;
      const unsigned int nX = (unsigned int)(xlow + 2) * (unsigned int)(g.n[1] - 4) * (unsigned int)nzc;
      const int64_t cnt = (int64_t)nZ + nY + nX;
      k_conv_diff_shell<3><<<(int)std::min<int64_t>((cnt + 255) / 256, s.nsm * 16), 256, 0, s.stream>>>(a, nZ, nY, nX, xlow);
      ++s.launches;
    }
  } else {
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
int mom_bdim2(Sim& s, bool accumulate, float scale) {
  if (s.mask_dirty) {
    RLAUNCH(s, k_body_mask, s.g, s.g, s.V, s.mu0, s.mu1, s.body_mask);
    s.mask_dirty = false;
  }
  wl_prof_begin(s, WL_PROF_BDIM2);
  RLAUNCH(s, k_bdim2, s.g, s.g, s.u, s.f, s.V, s.mu0, s.mu1, s.body_mask, accumulate ? 1 : 0, scale);
  // read f, V, μ₀ (3D), μ₁ (D²) [+ u (D)], write u (D)
  // algorithmic traffic away from a body: read f (D) [+ u (D)] + mask (¼), write u (D)
  wl_prof_end(s, WL_PROF_BDIM2, 4.0 * (double)s.g.interior * ((accumulate ? 3 : 2) * s.D + 0.25));
  return wl_cuda_check(s, cudaGetLastError(), "k_bdim2");
}
int mom_scale_u(Sim& s, double scale) {
  RLAUNCH(s, k_scale_u, s.g, s.g, s.u, scale);
  return wl_cuda_check(s, cudaGetLastError(), "k_scale_u");
}
// project!(a,b,w) (Flow.jl:137-145)
// `skip_unscale`: leave x scaled by dt (the caller's next project! undoes it in its own first pass);
// `prediv`: the Float32 dt such a previous project! left in x
int mom_project(Sim& s, double w, int* n, int slot, bool defer_count, bool skip_unscale, float prediv) {
  const float dt32 = s.dt.back();
  const bool f32 = (w == 1.0);                 // dt = w*Δt: Float32 for the Int w=1, Float64 for w=0.5
  const double dtw = f32 ? (double)dt32 : w * (double)dt32;
  Lvl& b = s.levels[0];
  float4* xall = reinterpret_cast<float4*>(b.x - s.g.lead);
  const int64_t n4 = s.g.comp / 4;
  const int sgrid = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)s.nsm * 8);
  wl_prof_begin(s, WL_PROF_PROJECT);
  RLAUNCH(s, k_project_pre, s.g, s.g, s.u, b.z);
  k_scale_alloc<<<sgrid, 256, 0, s.stream>>>(xall, n4, (float)dtw, dtw, f32 ? 1 : 0, 0, prediv);
  ++s.launches;
  wl_prof_end(s, WL_PROF_PROJECT, 4.0 * (double)s.g.interior * (s.D + 1 + 2));
  int rc = pois_solve(s, 1e-4, s.pois_kind == WL_POIS_MULTILEVEL ? 32 : 1000, n, true, slot, defer_count);
  if (rc) return rc;
  wl_prof_begin(s, WL_PROF_PROJECT);
  RLAUNCH(s, k_project_correct, s.g, b, s.u);
  if (!skip_unscale) {
    k_scale_alloc<<<sgrid, 256, 0, s.stream>>>(xall, n4, (float)dtw, dtw, f32 ? 1 : 0, 1, 0.f);
    ++s.launches;
  }
  wl_prof_end(s, WL_PROF_PROJECT, 4.0 * (double)s.g.interior * (wl_words(s, 3 * s.D + 1, s.D) + (skip_unscale ? 0 : 2)));
  return wl_cuda_check(s, cudaGetLastError(), "project");
}
// CFL(a) (Flow.jl:172-175)
int mom_cfl(Sim& s, double* dt) {
  const FinArgs fa = fin_misc(s);
  RLAUNCH(s, k_cfl, s.g, s.g, s.u, s.sigma, s.ws, fa);
  int rc = post_reduce(s, FIN_CFL, fa, true);
  if (rc) return rc;
  {
    const int64_t cnt = shell_count(s.g);
    k_shell_max<<<(int)std::min<int64_t>((cnt + 255) / 256, 256), 256, 0, s.stream>>>(s.g, s.sigma, s.ws, fa, cnt);
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
  if ((rc = mom_conv_diff(s, s.u0, true, gp))) return rc;
  if ((rc = mom_bdim2(s, false, 1.f))) return rc;
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  if (s.exitBC) {
    if ((rc = bc_exit(s, U, s.dt.back()))) return rc;
    if ((rc = dist_halo(s, s.g, s.u, 1, 2))) return rc;  // the exit plane changed after BC!'s exchange
  }
  const bool defer = s.vc_solve && s.pois_kind == WL_POIS_MULTILEVEL;
  // the pressure is not observable between the two project! calls of a step: the predictor leaves it scaled
  // and the corrector's first pass does x = (x / Δt) * (Δt/2) with the reference's two roundings
  if ((rc = mom_project(s, 1.0, n_pred, 0, defer, true, 0.f))) return rc;
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  // corrector u → u¹
  if ((rc = mom_conv_diff(s, s.u, true, gc))) return rc;
  if ((rc = mom_bdim2(s, true, 0.5f))) return rc;
  if ((rc = bc_vec(s, s.g, s.u, U, s.exitBC, s.g.per, 2))) return rc;
  if ((rc = mom_project(s, 0.5, n_corr, 1, defer, false, s.dt.back()))) return rc;
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
