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

}  // namespace

int metrics_run(Sim& s, int which, const double* x0, double* out) {
  MetricArgs a{};
  a.g = s.g; a.p = s.p; a.u = s.u; a.f = s.f; a.nds = s.nds;
  a.mnu = -s.nu;
  a.which = which;
  for (int k = 0; k < 3; ++k) a.x0[k] = (x0 && k < s.D) ? x0[k] : 0.0;
  a.out = s.ws.partials + (size_t)WL_MAXBLOCKS * WL_MAXRED;  // spare slots behind the partials (8 doubles)
  // df .= 0 (Metrics.jl:97,115,138): ghost cells included
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
