// wl_metrics.cu — surface-integrated loads (reference src/Metrics.jl:23-31,94-141): pressure_force,
// viscous_force, pressure_moment.  SURVEY.md §8(f) N3: the README workflow calls these every output
// step, and in the reference each is a kernel launch plus a host-side `sum` over a D-component field.
//
// Here each is ONE kernel: the per-cell product (in Float64, as `T = promote_type(Float64,eltype(p))`
// makes it in the reference), its Float32 store into flow.f (the reference's df scratch) and the
// Float64 sums of the D components, finished by the last block.  The masked surface normal
// nds(body, loc(0,I), t) comes from the host as a Float64 array (the geometry stays in Julia).
#include <algorithm>
#include <cstring>

#include "wl_internal.h"

namespace {

// ∂(i,j,I,u) (Metrics.jl:27-31): ∂uᵢ/∂xⱼ at the centre of cell I
__device__ __forceinline__ float dudx(const GridD& g, const float* __restrict__ u, int i, int j, int64_t I) {
  const float* ui = u + i * g.comp;
  const int64_t si = g.s[i], sj = g.s[j];
  if (i == j) return ui[I + si] - ui[I];
  return (ui[I + sj] + ui[I + sj + si] - ui[I - sj] - ui[I - sj + si]) / 4;
}

struct MetricArgs {
  GridD g;
  const float* p;
  const float* u;
  float* f;           // df
  const double* nds;
  double x0[3];
  float mnu;          // -ν
  int which;          // 0 pressure_force, 1 viscous_force, 2 pressure_moment
  double* out;        // D totals of this rank
};

template <int D>
__global__ void __launch_bounds__(256) k_metric(MetricArgs a, ReduceWS ws) {
  const GridD& g = a.g;
  double acc[3] = {0.0, 0.0, 0.0};
  const Box b = box_interior(g);
  for_cells(g, b, [&](int i0, int j0, int k0, int64_t I) {
    float v[3] = {0.f, 0.f, 0.f};
    double n[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < D; ++i) n[i] = a.nds[I + i * g.comp];
    if (a.which == 0) {  // Metrics.jl:98
      const double p = (double)a.p[I];
#pragma unroll
      for (int i = 0; i < D; ++i) v[i] = (float)(p * n[i]);
    } else if (a.which == 1) {  // Metrics.jl:116, ∇²u (:107-108)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const float m = a.mnu * (dudx(g, a.u, i, j, I) + dudx(g, a.u, j, i, I));
          const double term = (double)m * n[j];
          s = j == 0 ? term : s + term;
        }
        v[i] = (float)s;
      }
    } else {  // Metrics.jl:139; loc(0,I) = I - 1.5 with 1-based I
      const int c[3] = {i0, j0, k0};
      double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int i = 0; i < D; ++i) r[i] = ((double)(c[i] + (i == 2 ? g.zoff : 0)) - 0.5) - a.x0[i];
      const double p = (double)a.p[I];
      if (D == 2) {
        v[0] = v[1] = (float)(p * (r[0] * n[1] - r[1] * n[0]));
      } else {
        v[0] = (float)(p * (r[1] * n[2] - r[2] * n[1]));
        v[1] = (float)(p * (r[2] * n[0] - r[0] * n[2]));
        v[2] = (float)(p * (r[0] * n[1] - r[1] * n[0]));
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      a.f[I + i * g.comp] = v[i];
      acc[i] += (double)v[i];
    }
  });
  double tot[3];
  if (grid_reduce<3>(acc, ws, tot)) {
    for (int i = 0; i < 3; ++i) a.out[i] = tot[i];
  }
}

// ---- derived fields: @inside σ[I] = f(I,u) (Metrics.jl:14-77) --------------------------------------
struct DerivedArgs {
  GridD g;
  const float* u;
  float* sigma;
  int kind, comp, hasU;
  double U[3], axis[3], center[3];
};
// middle eigenvalue of a symmetric 3×3 matrix (closed form; the reference calls LAPACK's eigvals)
__device__ __forceinline__ double sym3_mid_eig(double a00, double a01, double a02, double a11, double a12, double a22) {
  const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
  const double q = (a00 + a11 + a22) / 3.0;
  if (p1 == 0.0) {  // diagonal: the median of the diagonal
    const double lo = fmin(a00, a11), hi = fmax(a00, a11);
    return fmax(lo, fmin(hi, a22));
  }
  const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
  const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
  const double pp = sqrt(p2 / 6.0);
  const double c00 = b00 / pp, c11 = b11 / pp, c22 = b22 / pp, c01 = a01 / pp, c02 = a02 / pp, c12 = a12 / pp;
  const double det = c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02);
  const double r = fmin(1.0, fmax(-1.0, det / 2.0));
  const double phi = acos(r) / 3.0;
  const double e1 = q + 2.0 * pp * cos(phi);                                   // largest
  const double e3 = q + 2.0 * pp * cos(phi + 2.0943951023931954923084289221863);  // smallest
  return 3.0 * q - e1 - e3;
}
template <int D>
__global__ void __launch_bounds__(256) k_derived(DerivedArgs a) {
  const GridD& g = a.g;
  const Box b = box_interior(g);
  for_cells(g, b, [&](int i0, int j0, int k0, int64_t I) {
    float out = 0.f;
    if (a.kind == WL_DF_KE) {  // Metrics.jl:20-22
      if (a.hasU) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) { const double t = (double)(a.u[I + i * g.comp] + a.u[I + g.s[i] + i * g.comp]) - 2.0 * a.U[i]; s += t * t; }
        out = (float)(0.125 * s);
      } else {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < D; ++i) { const float t = a.u[I + i * g.comp] + a.u[I + g.s[i] + i * g.comp]; s += t * t; }
        out = (float)(0.125 * (double)s);
      }
    } else if (a.kind == WL_DF_CURL) {  // Metrics.jl:55: permute((j,k)->∂(j,CI(I,k),u), i), ∂(a,I,f) = f[I]-f[I-δa] (Flow.jl:1)
      const int i = a.comp - 1, j = (i + 1) % 3, k = (i + 2) % 3;
      const float fjk = a.u[I + k * g.comp] - a.u[I - g.s[j] + k * g.comp];
      const float fkj = a.u[I + j * g.comp] - a.u[I - g.s[k] + j * g.comp];
      out = fjk - fkj;
    } else if (D == 3) {
      float J[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) J[i][j] = dudx(g, a.u, i, j, I);
      // ω(I,u) (Metrics.jl:61): ωᵢ = ∂(k,j) - ∂(j,k), (j,k) = (i%3+1, (i+1)%3+1)
      const float w[3] = {J[2][1] - J[1][2], J[0][2] - J[2][0], J[1][0] - J[0][1]};
      if (a.kind == WL_DF_OMEGA_MAG) {
        out = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      } else if (a.kind == WL_DF_LAMBDA2) {  // Metrics.jl:40-44
        float S[3][3], O[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) { S[i][j] = (J[i][j] + J[j][i]) / 2; O[i][j] = (J[i][j] - J[j][i]) / 2; }
        float M[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            float s2 = 0.f, o2 = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) { s2 += S[i][k] * S[k][j]; o2 += O[i][k] * O[k][j]; }
            M[i][j] = s2 + o2;
          }
        out = (float)sym3_mid_eig(M[0][0], M[0][1], M[0][2], M[1][1], M[1][2], M[2][2]);
      } else {  // ω_θ (Metrics.jl:74-78): θ = z × (loc(0,I) - center)
        const int c[3] = {i0, j0, k0 + g.zoff};
        double d[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) d[i] = (double)((float)c[i] - 0.5f) - a.center[i];
        const double th[3] = {a.axis[1] * d[2] - a.axis[2] * d[1], a.axis[2] * d[0] - a.axis[0] * d[2], a.axis[0] * d[1] - a.axis[1] * d[0]};
        const double n = sqrt(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
        out = n == 0.0 ? 0.f : (float)((th[0] * (double)w[0] + th[1] * (double)w[1] + th[2] * (double)w[2]) / n);
      }
    }
    a.sigma[I] = out;
  });
}

}  // namespace

int derived_run(Sim& s, int kind, int comp, const double* U, const double* axis, const double* center) {
  if (kind < WL_DF_KE || kind > WL_DF_OMEGA_THETA) return wl_fail(WL_ERR_ARG, "wl_derived: bad kind");
  if (kind == WL_DF_CURL && (comp < 1 || comp > 3 || (s.D == 2 && comp != 3))) return wl_fail(WL_ERR_ARG, "wl_derived: curl component must be 1..3 (3 in 2-D)");
  if (kind >= WL_DF_OMEGA_MAG && s.D != 3) return wl_fail(WL_ERR_ARG, "wl_derived: this diagnostic needs a 3-D flow");
  if (kind == WL_DF_OMEGA_THETA && (!axis || !center)) return wl_fail(WL_ERR_ARG, "wl_derived: ω_θ needs an axis and a centre");
  DerivedArgs a{};
  a.g = s.g; a.u = s.u; a.sigma = s.sigma; a.kind = kind; a.comp = comp; a.hasU = U ? 1 : 0;
  for (int k = 0; k < 3; ++k) {
    a.U[k] = (U && k < s.D) ? U[k] : 0.0;
    a.axis[k] = (axis && k < s.D) ? axis[k] : 0.0;
    a.center[k] = (center && k < s.D) ? center[k] : 0.0;
  }
  const dim3 blk = wl_pick_block(s.g);
  const int grid = wl_grid_for(s, box_tiles(box_interior(s.g), blk));
  if (s.D == 2) k_derived<2><<<grid, blk, 0, s.stream>>>(a);
  else k_derived<3><<<grid, blk, 0, s.stream>>>(a);
  ++s.launches;
  return wl_cuda_check(s, cudaGetLastError(), "k_derived");
}

int metrics_run(Sim& s, int which, const double* x0, double* out) {
  MetricArgs a{};
  a.g = s.g; a.p = s.p; a.u = s.u; a.f = s.f; a.nds = s.nds;
  a.mnu = -s.nu;
  a.which = which;
  for (int k = 0; k < 3; ++k) a.x0[k] = (x0 && k < s.D) ? x0[k] : 0.0;
  a.out = s.ws.partials + (size_t)WL_MAXBLOCKS * WL_MAXRED;  // spare slots behind the partials (8 doubles)
  // df .= 0 (Metrics.jl:97,115,138): ghost cells included
  s.f_stale = false;  // flow.f is the scratch of the Metrics functions (Metrics.jl:96): overwritten whole
  int rc = wl_fill(s, s.f - s.g.lead, 0.f, s.g.comp * s.D);
  if (rc) return rc;
  const dim3 blk = wl_pick_block(s.g);
  const int grid = wl_grid_for(s, box_tiles(box_interior(s.g), blk));
  if (s.D == 2) k_metric<2><<<grid, blk, 0, s.stream>>>(a, s.ws);
  else k_metric<3><<<grid, blk, 0, s.stream>>>(a, s.ws);
  ++s.launches;
  if ((rc = wl_cuda_check(s, cudaGetLastError(), "k_metric"))) return rc;
  if ((rc = dist_allreduce(s, a.out, 3, false))) return rc;  // slab ranks: every rank gets the global sums
  if ((rc = wl_read_scalars(s, a.out, 3 * sizeof(double)))) return rc;
  double h[3];
  memcpy(h, s.h_pinned, sizeof(h));
  for (int i = 0; i < s.D; ++i) out[i] = h[i];
  return 0;
}
