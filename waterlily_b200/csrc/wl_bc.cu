// wl_bc.cu — boundary-condition kernels: BC!, perBC!, exitBC! (reference src/util.jl:192-231).
//
// The reference applies BC! as D² sequential passes (component i outer, direction j inner) of
// three tiny launches each, and relies on the pass order to fill edges and corners transitively.
// Here ONE kernel per BC! resolves every ghost cell directly: walking the directions from the
// last to the first, a cell written by pass j takes the value its source cell had when pass j
// ran, which only earlier passes (smaller j) can have changed — so the source is resolved
// recursively over smaller j until a cell no pass writes (or a Dirichlet constant) is reached.
// All reads therefore hit cells the kernel never writes: no ordering hazards, same values.
#include "wl_internal.h"

namespace {

struct BCArgs {
  float A[3];
  int per[3];
  int saveexit;
};

// BC!(a,A,saveexit,perdir) (util.jl:192-210)
// Work units are rows of the planes c[j] ∈ {0, 1, n-1} of every direction j: a block decodes its unit once (32-bit),
// its threads sweep the row — along x for the y- and z-planes (coalesced), along y for the x-planes — and fetch the D
// source values of a cell before storing any (independent loads).  A flat per-cell enumeration with 64-bit div/mod
// spent 127 µs on the 2.4 M ghost cells of 514³; edge cells are visited by more than one unit and receive the same value.
__device__ __forceinline__ int bc_units(const GridD& g, int j) { return 3 * (g.D == 3 ? g.n[j == 2 ? 1 : 2] : 1); }
__global__ void __launch_bounds__(256) k_bc_vec(GridD g, float* __restrict__ a, BCArgs args, int nunits) {
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    int u = unit, j = 0;
    for (; j < g.D - 1; ++j) {
      const int nu = bc_units(g, j);
      if (u < nu) break;
      u -= nu;
    }
    const int nb = bc_units(g, j) / 3, pl = u / nb, cb = u - pl * nb;
    const int cj = pl == 0 ? 0 : (pl == 1 ? 1 : g.n[j] - 1);
    const int na = g.n[j == 0 ? 1 : 0];
    for (int ca = threadIdx.x; ca < na; ca += blockDim.x) {
      // 3-D: j=0 → (cj, ca, cb); j=1 → (ca, cj, cb); j=2 → (ca, cb, cj).  2-D: j=0 → (cj, ca); j=1 → (ca, cj)
      const int c0[3] = {j == 0 ? cj : ca, j == 0 ? ca : (j == 1 ? cj : cb), g.D == 3 ? (j == 2 ? cj : cb) : 0};
      // z-planes owned by a neighbouring slab are filled by the halo exchange, not by BC!
      if (g.D > 2 && ((c0[2] == 0 && g.zlo_halo) || (c0[2] == g.n[2] - 1 && g.zhi_halo))) continue;
      const int64_t I0 = c0[0] + c0[1] * g.s[1] + c0[2] * g.s[2];
      float v[3];
      bool wr[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        wr[i] = false;
        if (i >= g.D) continue;
        int c[3] = {c0[0], c0[1], c0[2]};
        bool written = false, constant = false;
#pragma unroll
        for (int jj = 2; jj >= 0; --jj) {
          if (jj >= g.D || constant) continue;
          const int n = g.n[jj];
          const bool lo = !(jj == 2 && g.zlo_halo), hi = !(jj == 2 && g.zhi_halo);  // physical boundary on that side?
          if (args.per[jj]) {
            if (c[jj] == 0 && lo) { c[jj] = n - 2; written = true; }
            else if (c[jj] == n - 1 && hi) { c[jj] = 1; written = true; }
          } else if (i == jj) {
            if ((c[jj] <= 1 && lo) || (c[jj] == n - 1 && hi && !(args.saveexit && i == 0))) { constant = true; written = true; }
          } else {
            if (c[jj] == 0 && lo) { c[jj] = 1; written = true; }
            else if (c[jj] == n - 1 && hi) { c[jj] = n - 2; written = true; }
          }
        }
        wr[i] = written;
        if (written) v[i] = constant ? args.A[i] : a[i * g.comp + c[0] + c[1] * g.s[1] + c[2] * g.s[2]];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (wr[i]) a[i * g.comp + I0] = v[i];
    }
  }
}

// perBC!(a,perdir) (util.jl:227-231); `active` (may be null) predicates the launch on a running pcg!
__global__ void __launch_bounds__(256) k_per_bc(GridD g, float* __restrict__ a, int p0, int p1, int p2, int64_t count, const int* active) {
  if (active && !*active) return;
  const int per[3] = {p0, p1, p2};
  // every ghost cell of the global domain takes the value of its fully wrapped image — what the reference's ordered
  // per-direction slice copies leave behind (edges and corners included); ghost cells of non-periodic directions only stay
  for_global_ghosts(g, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, [&](const int (&c0)[3]) {
    int c[3] = {c0[0], c0[1], c0[2]};
    bool written = false;
    for (int j = g.D - 1; j >= 0; --j) {
      if (!per[j]) continue;
      if (c[j] == 0) { c[j] = g.n[j] - 2; written = true; }
      else if (c[j] == g.n[j] - 1) { c[j] = 1; written = true; }
    }
    if (written) a[c0[0] + c0[1] * g.s[1] + c0[2] * g.s[2]] = a[c[0] + c[1] * g.s[1] + c[2] * g.s[2]];
  });
}

// exitBC! (util.jl:216-222): 1-D convective update of the exit plane + its mean
__global__ void __launch_bounds__(256) k_exit1(GridD g, float* __restrict__ u, const float* __restrict__ u0, double Udt, ReduceWS ws, FinArgs fa) {
  const int ny = g.n[1] - 2, nz = g.D > 2 ? g.n[2] - 2 : 1;
  const int64_t count = (int64_t)ny * nz;
  const int i = g.n[0] - 1;
  double acc[1] = {0.0};
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = 1 + (int)(idx % ny), k = g.D > 2 ? 1 + (int)(idx / ny) : 0;
    const int64_t I = i + j * g.s[1] + k * g.s[2];
    const float v = (float)((double)u0[I] - Udt * (double)(u0[I] - u0[I - 1]));
    u[I] = v;
    acc[0] += (double)v;
  }
  double tot[1];
  if (grid_reduce<1>(acc, ws, tot)) finish_reduction(FIN_EXIT, fa, tot[0], 0.0);
}
__global__ void __launch_bounds__(256) k_exit2(GridD g, float* __restrict__ u, const MiscScalars* misc) {
  const int ny = g.n[1] - 2, nz = g.D > 2 ? g.n[2] - 2 : 1;
  const int64_t count = (int64_t)ny * nz;
  const int i = g.n[0] - 1;
  const double du = (double)misc->exit_mean_minus_U;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = 1 + (int)(idx % ny), k = g.D > 2 ? 1 + (int)(idx / ny) : 0;
    const int64_t I = i + j * g.s[1] + k * g.s[2];
    u[I] = (float)((double)u[I] - du);
  }
}

inline int flat_grid(int64_t count, int cap) {
  int64_t b = (count + 255) / 256;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

}  // namespace

int bc_vec(Sim& s, const GridD& g, float* a, const double* A, bool saveexit, const int* per, int halo_depth) {
  BCArgs args;
  for (int k = 0; k < 3; ++k) { args.A[k] = k < g.D ? (float)A[k] : 0.f; args.per[k] = k < g.D ? per[k] : 0; }
  args.saveexit = saveexit ? 1 : 0;
  int nunits = 0;
  for (int j = 0; j < g.D; ++j) nunits += 3 * (g.D == 3 ? g.n[j == 2 ? 1 : 2] : 1);
  k_bc_vec<<<nunits < s.nsm * 16 ? nunits : s.nsm * 16, 256, 0, s.stream>>>(g, a, args, nunits);
  ++s.launches;
  int rc = wl_cuda_check(s, cudaGetLastError(), "k_bc_vec");
  if (rc) return rc;
  return dist_halo(s, g, a, g.D, halo_depth);  // slab faces: full planes incl. the x/y ghosts just set
}

int bc_per(Sim& s, const GridD& g, float* a, const int* per, const int* active) {
  bool any = false;
  for (int k = 0; k < g.D; ++k) any |= per[k] != 0;
  if (any) {  // perBC!(a,::Tuple{}) = nothing
    const int64_t cnt = global_ghost_count(g);
    k_per_bc<<<flat_grid(cnt, s.nsm * 8), 256, 0, s.stream>>>(g, a, per[0], g.D > 1 ? per[1] : 0, g.D > 2 ? per[2] : 0, cnt, active);
    ++s.launches;
    const int rc = wl_cuda_check(s, cudaGetLastError(), "k_per_bc");
    if (rc) return rc;
  }
  return dist_halo(s, g, a, 1, 1);
}

int bc_exit(Sim& s, const double* U, float dt) {
  const GridD& g = s.g;
  const int64_t cnt = (int64_t)(g.n[1] - 2) * (g.D > 2 ? g.n[2] - 2 : 1);
  FinArgs fa = fin_misc(s);
  fa.aux = U[0];
  fa.count = (double)((int64_t)(g.n[1] - 2) * (g.D > 2 ? s.Nz_global : 1));  // length(exitR) of the whole domain
  k_exit1<<<flat_grid(cnt, 256), 256, 0, s.stream>>>(g, s.u, s.u0, U[0] * (double)dt, s.ws, fa);
  ++s.launches;
  int rc = post_reduce(s, FIN_EXIT, fa);
  if (rc) return rc;
  k_exit2<<<flat_grid(cnt, 256), 256, 0, s.stream>>>(g, s.u, s.d_misc);
  ++s.launches;
  return wl_cuda_check(s, cudaGetLastError(), "k_exit");
}
