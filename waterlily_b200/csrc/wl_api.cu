// wl_api.cu — the C ABI of include/waterlily_b200.h: handle lifetime, field transfer, and thin
// exports of each reference function on the path.  Every export catches its own errors and
// returns a status; nothing throws across the boundary.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "wl_internal.h"

namespace {
thread_local std::string g_err;
}

int wl_fail(int code, const std::string& msg) { g_err = msg; return code; }
int wl_cuda_check(Sim& s, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return s.sticky;
  if (!s.sticky) s.sticky = WL_ERR_CUDA;
  return wl_fail(WL_ERR_CUDA, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

dim3 wl_pick_block(const GridD& g) {
  const int nx = g.n[0];
  const int bx = nx >= 96 ? 128 : nx >= 48 ? 64 : nx >= 24 ? 32 : nx >= 12 ? 16 : 8;
  if (g.D == 2) return dim3(bx, 256 / bx, 1);
  switch (bx) {
    case 128: return dim3(128, 2, 1);
    case 64: return dim3(64, 2, 2);
    case 32: return dim3(32, 4, 2);
    case 16: return dim3(16, 4, 4);
    default: return dim3(8, 8, 4);
  }
}
int wl_grid_for(const Sim& s, int64_t ntiles) {
  const int64_t cap = std::min<int64_t>((int64_t)s.nsm * 8, WL_MAXBLOCKS);
  return (int)std::max<int64_t>(1, std::min(ntiles, cap));
}

void wl_prof_begin(Sim& s, int cls) {
  if (!s.profiling) return;
  ProfClass& p = s.prof[cls];
  if (p.used + 2 > p.ev.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    p.ev.push_back(a); p.ev.push_back(b);
  }
  cudaEventRecord(p.ev[p.used], s.stream);
}
void wl_prof_end(Sim& s, int cls, double bytes) {
  if (!s.profiling) return;
  ProfClass& p = s.prof[cls];
  cudaEventRecord(p.ev[p.used + 1], s.stream);
  p.used += 2;
  p.launches += 1;
  p.bytes += bytes;
}

namespace {
__global__ void k_fill(float* p, float v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

void grid_set(GridD& g, int D, const int* Ng, const int* per) {
  g.D = D;
  for (int k = 0; k < 3; ++k) { g.n[k] = k < D ? Ng[k] : 1; g.per[k] = (k < D && per) ? (per[k] != 0) : 0; }
  const int64_t pitch = ((int64_t)g.n[0] + WL_OX + 3) / 4 * 4;  // rows start 16-byte aligned at i=1
  g.s[0] = 1; g.s[1] = pitch; g.s[2] = pitch * g.n[1];
  // 3-D fields carry one spare z-plane below plane 0 and above plane n-1: the depth-2 halo of u/u⁰
  // in a z-slab decomposition lives there (QUICK reads I-2δz from the first interior plane)
  const int64_t spare = D == 3 ? g.s[2] : 0;
  g.lead = WL_OX + spare;
  g.comp = ((g.lead + pitch * g.n[1] * g.n[2] + spare + 8) + 3) / 4 * 4;
  g.interior = 1;
  for (int k = 0; k < D; ++k) g.interior *= g.n[k] - 2;
  g.interior_global = g.interior;
  g.zlo_halo = g.zhi_halo = 0;
  g.zoff = 0;
}

// Fields are bump-allocated from ONE arena per handle (a single cudaMalloc): the arena is what the
// z-slab ranks export to each other over CUDA IPC, and since every rank runs the same allocation
// sequence, a field sits at the same offset in every rank's arena (wl_p2p.cuh).  A first, dry pass
// of the allocation sequence (s.arena == nullptr) only measures.
void* arena_take(Sim& s, size_t bytes, int* rc) {
  const size_t off = (s.arena_used + 255) / 256 * 256;
  s.arena_used = off + bytes;
  if (!s.arena) return nullptr;  // dry pass
  if (s.arena_used > s.arena_bytes) { if (!*rc) *rc = wl_fail(WL_ERR_ARG, "internal: arena overflow"); return nullptr; }
  return s.arena + off;
}
float* dev_alloc(Sim& s, const GridD& g, int ncomp, int* rc) {
  float* p = (float*)arena_take(s, sizeof(float) * (size_t)g.comp * ncomp, rc);
  return p ? p + g.lead : nullptr;
}

float* field_ptr(Sim& s, int field, int* ncomp) {
  const int D = s.D;
  switch (field) {
    case WL_U: *ncomp = D; return s.u;
    case WL_U0: *ncomp = D; return s.u0;
    case WL_F: *ncomp = D; return s.f;
    case WL_P: *ncomp = 1; return s.p;
    case WL_SIGMA: *ncomp = 1; return s.sigma;
    case WL_V: *ncomp = D; return s.V;
    case WL_MU0: *ncomp = D; return s.mu0;
    case WL_MU1: *ncomp = D * D; return s.mu1;
  }
  return nullptr;
}
float* level_ptr(Sim& s, int level, int lf, int* ncomp) {
  if (level < 0 || level >= (int)s.levels.size()) return nullptr;
  Lvl& l = s.levels[level];
  *ncomp = 1;
  switch (lf) {
    case WL_LVL_L: *ncomp = s.D; return l.L;
    case WL_LVL_D: return l.Dg;
    case WL_LVL_ID: return l.iD;
    case WL_LVL_X: return l.x;
    case WL_LVL_EPS: return l.eps;
    case WL_LVL_R: return l.r;
    case WL_LVL_Z: return l.z;
  }
  return nullptr;
}
// dense rows (n0 floats) <-> re-pitched rows (pitch s1) for `rows` consecutive rows of one component
template <class T>
__global__ void k_repitch(T* __restrict__ dev, T* __restrict__ dense, int n0, int64_t s1, int64_t rows, int to_device) {
  const int64_t total = rows * n0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / n0;
    const int i = (int)(t - r * n0);
    if (to_device) dev[r * s1 + i] = dense[t];
    else dense[t] = dev[r * s1 + i];
  }
}

// host (dense, column-major, padded, GLOBAL domain) <-> device (re-pitched rows, this rank's slab).
// `slab`: the field lives on a z-distributed grid; the host array always has the global shape.
// Download writes only the planes this rank owns (its interior planes and global ghost planes).
// The PCIe copies are CONTIGUOUS chunks of whole planes through a device staging buffer and a
// re-pitch kernel: a strided cudaMemcpy2D of ~2 KB rows is descriptor-bound (measured 2.4 GB/s at 514³).
template <class T>
int copy_field(Sim& s, const GridD& g, T* dev, T* host, int ncomp, bool to_device, int nz_global_padded) {
  const int64_t dense = (int64_t)g.n[0] * g.n[1] * nz_global_padded;
  int k0 = 0, k1 = g.n[2] - 1;
  if (!to_device) {  // planes of a neighbouring slab are not ours to report; the global ghost planes are
    if (g.zlo_halo && g.zoff != 0) k0 = 1;
    if (g.zhi_halo && g.zoff + g.n[2] != nz_global_padded) k1 = g.n[2] - 2;
  }
  const int64_t plane = (int64_t)g.n[0] * g.n[1];  // floats of one dense plane
  if (!s.stage) {
    const size_t want = (size_t)64 << 20;
    if (cudaMalloc(&s.stage, want) != cudaSuccess) return wl_cuda_check(s, cudaGetLastError(), "cudaMalloc (staging)");
    s.stage_bytes = want;
  }
  const int64_t cap = (int64_t)(s.stage_bytes / sizeof(T));
  const int pchunk = (int)std::max<int64_t>(1, std::min<int64_t>(cap / plane, k1 - k0 + 1));
  if (plane > cap) return wl_fail(WL_ERR_ARG, "field plane larger than the staging buffer");
  for (int c = 0; c < ncomp; ++c) {
    for (int k = k0; k <= k1; k += pchunk) {
      const int np = std::min(pchunk, k1 - k + 1);
      T* d = dev + c * g.comp + (int64_t)k * g.s[2];
      T* h = host + c * dense + (int64_t)(g.zoff + k) * plane;
      const size_t bytes = sizeof(T) * (size_t)plane * np;
      const int64_t rows = (int64_t)g.n[1] * np;
      const int grid = (int)std::min<int64_t>((rows * g.n[0] + 255) / 256, (int64_t)s.nsm * 16);
      cudaError_t e = cudaSuccess;
      if (to_device) {
        e = cudaMemcpyAsync(s.stage, h, bytes, cudaMemcpyHostToDevice, s.stream);
        if (e == cudaSuccess) k_repitch<T><<<grid, 256, 0, s.stream>>>(d, (T*)s.stage, g.n[0], g.s[1], rows, 1);
      } else {
        k_repitch<T><<<grid, 256, 0, s.stream>>>(d, (T*)s.stage, g.n[0], g.s[1], rows, 0);
        e = cudaMemcpyAsync(h, s.stage, bytes, cudaMemcpyDeviceToHost, s.stream);
      }
      ++s.launches;
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e != cudaSuccess) return wl_cuda_check(s, e, "field copy");
    }
  }
  return wl_cuda_check(s, cudaStreamSynchronize(s.stream), "cudaStreamSynchronize");
}
// padded global z-size of the grid a level lives on
int nz_padded_global(const Sim& s, int level) {
  const GridD& g = s.levels[level].g;
  if (s.D < 3) return 1;
  return (s.dist.nranks > 1 && s.lvl_dist[level]) ? (g.n[2] - 2) * s.dist.nranks + 2 : g.n[2];
}
inline bool divisible(int N) { return N % 2 == 0 && N > 4; }  // MultiLevelPoisson.jl:36
}  // namespace

int wl_fill(Sim& s, float* p, float v, int64_t n) {
  k_fill<<<(int)std::min<int64_t>((n + 255) / 256, s.nsm * 8), 256, 0, s.stream>>>(p, v, n);
  ++s.launches;
  return wl_cuda_check(s, cudaGetLastError(), "k_fill");
}
int wl_read_scalars(Sim& s, const void* dptr, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(s.h_pinned, dptr, bytes, cudaMemcpyDeviceToHost, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  return wl_cuda_check(s, e, "scalar readback");
}

#define S (*(Sim*)s)
#define GUARD(s)                                                         \
  if (!(s)) return wl_fail(WL_ERR_ARG, "null handle");                   \
  if (S.sticky) return wl_fail(S.sticky, "handle is in a failed state: " + g_err); \
  cudaSetDevice(S.device)

extern "C" {

#ifndef WL_DIST_REPL_CELLS_DEFAULT
#define WL_DIST_REPL_CELLS_DEFAULT (1 << 21)  // 128³: measured at 512³ on 4 GPUs 6.16 → 6.01 ms/step, on 8 GPUs 4.24 → 4.22 (profiles/r02_bench_{4,8}gpu_*)
#endif
// Which multigrid levels are z-distributed (1) and which are replicated on every rank (0).
static int plan_levels(int D, const int* N, int nranks, const std::vector<std::vector<int>>& dims, std::vector<int>& lvl_dist) {
  lvl_dist.assign(dims.size(), 0);
  if (nranks <= 1) return 0;
  if (D != 3) return wl_fail(WL_ERR_ARG, "wl_create_dist: the z-slab decomposition needs D = 3");
  if (N[2] % nranks != 0 || (N[2] / nranks) % 2 != 0)
    return wl_fail(WL_ERR_ARG, "wl_create_dist: N[3] must split into an even number of z-planes per rank");
  const char* e = getenv("WL_DIST_MIN_NZ");
  const int min_nz = e ? atoi(e) : 16;  // coarser levels are agglomerated: replicated on every rank
  // … and so is every level small enough that one GPU works it off faster than the ranks can shake hands about it: a
  // distributed level costs a peer-memory halo or all-reduce (6–9 µs) in most of its ≈21 phases per V-cycle
  const char* e2 = getenv("WL_DIST_REPL_CELLS");
  const long long repl_cells = e2 ? atoll(e2) : WL_DIST_REPL_CELLS_DEFAULT;
  lvl_dist[0] = 1;
  for (size_t l = 1; l < dims.size(); ++l) {
    const int nz = dims[l][2] - 2;
    const int nzf_loc = (dims[l - 1][2] - 2) / nranks;
    const long long cells = (long long)(dims[l][0] - 2) * (dims[l][1] - 2) * nz;
    lvl_dist[l] = lvl_dist[l - 1] && nz % nranks == 0 && nz / nranks >= min_nz && nzf_loc % 2 == 0 && cells > repl_cells;
  }
  for (size_t l = 1; l < dims.size(); ++l)
    if (lvl_dist[l - 1] && !lvl_dist[l] && (((dims[l - 1][2] - 2) / nranks) % 2 != 0 || (dims[l][2] - 2) % nranks != 0))
      return wl_fail(WL_ERR_ARG, "wl_create_dist: cannot agglomerate — odd number of z-planes per rank on the last distributed level");
  return 0;
}
static std::vector<std::vector<int>> level_dims(int D, const int* N, int pois_kind) {
  std::vector<std::vector<int>> dims;
  std::vector<int> g0(3, 1);
  for (int k = 0; k < D; ++k) g0[k] = N[k] + 2;
  dims.push_back(g0);
  if (pois_kind == WL_POIS_MULTILEVEL) {
    auto div_ok = [&](const std::vector<int>& n) { for (int k = 0; k < D; ++k) if (!divisible(n[k])) return false; return true; };
    while (div_ok(dims.back()) && (int)dims.size() <= 10) {  // MultiLevelPoisson.jl:53: length(levels) <= maxlevels (=10) → at most 11 levels
      std::vector<int> c(3, 1);
      for (int k = 0; k < D; ++k) c[k] = 1 + dims.back()[k] / 2;  // restrictML :20
      dims.push_back(c);
    }
  }
  return dims;
}

static int create_impl(wl_sim** out, int D, const int* N, double nu, double dt0, const int* perdir, int exitBC, int pois_kind, int device,
                       int rank, int nranks, const char* nccl_id) {
  if (!out || !N || (D != 2 && D != 3)) return wl_fail(WL_ERR_ARG, "wl_create: D must be 2 or 3");
  for (int k = 0; k < D; ++k) if (N[k] < 1) return wl_fail(WL_ERR_ARG, "wl_create: N must be positive");
  if (nranks < 1 || rank < 0 || rank >= nranks) return wl_fail(WL_ERR_ARG, "wl_create_dist: bad rank/nranks");
  // the MultiLevelPoisson size assertion comes first, as in the reference constructor
  const std::vector<std::vector<int>> dims = level_dims(D, N, pois_kind);  // GLOBAL padded sizes per level
  if (pois_kind == WL_POIS_MULTILEVEL && dims.size() <= 2) return wl_fail(WL_ERR_ARG, "MultiLevelPoisson requires size=a2ⁿ, where n>2");
  // z-slab decomposition: which levels are distributed, which replicated (coarse tail)
  std::vector<int> lvl_dist(dims.size(), 0);
  {
    const int rc0 = plan_levels(D, N, nranks, dims, lvl_dist);
    if (rc0) return rc0;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return wl_fail(WL_ERR_CUDA, "wl_create: no CUDA device available (libwaterlily_b200 has no CPU fallback)");
  if (device < 0 || device >= ndev) return wl_fail(WL_ERR_ARG, "wl_create: bad device ordinal");
  Sim* sp = new (std::nothrow) Sim;
  if (!sp) return wl_fail(WL_ERR_ARG, "out of host memory");
  Sim& a = *sp;
  a.D = D; a.device = device; a.nu = (float)nu; a.dt.push_back((float)dt0); a.exitBC = exitBC != 0; a.pois_kind = pois_kind;
  a.lvl_dist = lvl_dist;
  a.Nz_global = D == 3 ? N[2] : 1;
  a.per_z_global = (D == 3 && perdir && perdir[2]) ? 1 : 0;
  int rc = 0;
  cudaDeviceProp prop;
  if ((rc = wl_cuda_check(a, cudaSetDevice(device), "cudaSetDevice")) ||
      (rc = wl_cuda_check(a, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties"))) { delete sp; return rc; }
  a.nsm = prop.multiProcessorCount;
  if ((rc = wl_cuda_check(a, cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking), "cudaStreamCreate"))) { delete sp; return rc; }
  a.own_stream = true;
  if ((rc = dist_init(a, rank, nranks, nccl_id))) { wl_destroy((wl_sim*)sp); return rc; }
  // local grid of a level: a distributed level keeps Nz/R planes (+ghosts); z faces towards a
  // neighbour (or the periodic wrap) become halo faces and the local grid is not z-periodic
  auto make_grid = [&](GridD& g, const std::vector<int>& gd, bool distributed) {
    int n3[3] = {gd[0], gd[1], gd[2]};
    int per3[3] = {perdir ? perdir[0] : 0, (perdir && D > 1) ? perdir[1] : 0, (perdir && D > 2) ? perdir[2] : 0};
    int64_t glob = 1;
    for (int k = 0; k < D; ++k) glob *= gd[k] - 2;
    if (distributed) {
      const int nzl = (gd[2] - 2) / nranks;
      n3[2] = nzl + 2;
      per3[2] = 0;
    }
    grid_set(g, D, n3, per3);
    g.interior_global = glob;
    if (distributed) {
      g.zoff = rank * (n3[2] - 2);
      g.zlo_halo = (rank > 0 || a.per_z_global) ? 1 : 0;
      g.zhi_halo = (rank < nranks - 1 || a.per_z_global) ? 1 : 0;
    }
  };
  make_grid(a.g, dims[0], nranks > 1);
  // every field of the handle, in a fixed order (identical on every rank: the arenas are symmetric).
  // levels[1] aliases the flow arrays (WaterLily.jl:77); r, ϵ, D, iD are its own (Poisson.jl:33-34)
  auto alloc_fields = [&](int* prc) {
    a.arena_used = WL_P2P_FLAG_BYTES;  // the peer-communication flag block leads the arena
    a.u = dev_alloc(a, a.g, D, prc); a.u0 = dev_alloc(a, a.g, D, prc); a.f = dev_alloc(a, a.g, D, prc); a.w = dev_alloc(a, a.g, D, prc);
    a.p = dev_alloc(a, a.g, 1, prc); a.sigma = dev_alloc(a, a.g, 1, prc); a.V = dev_alloc(a, a.g, D, prc);
    a.mu0 = dev_alloc(a, a.g, D, prc); a.mu1 = dev_alloc(a, a.g, D * D, prc);
    {
      unsigned char* m = (unsigned char*)arena_take(a, (size_t)a.g.comp, prc);
      a.body_mask = m ? m + a.g.lead : nullptr;
      // clean map of the conv_diff! core tiles (64 x 8 cells): one byte per tile and plane
      const size_t ntile = (size_t)((a.g.n[0] + 63) / 64 + 1) * ((a.g.n[1] + 7) / 8 + 1) * (size_t)a.g.n[2];
      a.tile_clean = (unsigned char*)arena_take(a, ntile + 16, prc);
    }
    a.levels.clear();
    for (size_t l = 0; l < dims.size(); ++l) {
      Lvl lv{};
      make_grid(lv.g, dims[l], nranks > 1 && lvl_dist[l]);
      lv.sc = a.d_scalars + l;
      if (l == 0) { lv.x = a.p; lv.L = a.mu0; lv.z = a.sigma; }
      else { lv.L = dev_alloc(a, lv.g, D, prc); lv.x = dev_alloc(a, lv.g, 1, prc); lv.z = dev_alloc(a, lv.g, 1, prc); }
      lv.r = dev_alloc(a, lv.g, 1, prc); lv.eps = dev_alloc(a, lv.g, 1, prc);
      lv.Dg = dev_alloc(a, lv.g, 1, prc); lv.iD = dev_alloc(a, lv.g, 1, prc);
      // uniform-coefficient pack mask (one byte per four cells) and this level's constants: L = μ₀ = 1 on the
      // finest level, ×2^(D-2) per restrictL (½ · 2^(D-1) fine faces, MultiLevelPoisson.jl:10-16); D and iD as set_diag! computes them
      unsigned char* um = (lv.g.n[0] - 2) % 4 == 0 ? (unsigned char*)arena_take(a, (size_t)lv.g.comp / 4 + 16, prc) : nullptr;
      if (l < a.umask_store.size()) a.umask_store[l] = um; else a.umask_store.push_back(um);
      lv.umask = nullptr;
      lv.uL = l == 0 ? 1.f : a.levels[l - 1].uL * (D == 3 ? 2.f : 1.f);
      float sD = 0.f;
      for (int k = 0; k < D; ++k) sD -= (lv.uL + lv.uL);
      lv.uD = sD;
      lv.uiD = (sD * sD < 2 * 1.1920929e-07f) ? 0.f : 1.f / sD;
      lv.upad = 0;
      a.levels.push_back(lv);
    }
  };
  // solver scalars, reduction workspace, pinned readback (rank-local, outside the arena)
  void* tmp = nullptr;
  if (!rc && cudaMalloc(&tmp, sizeof(LevelScalars) * dims.size()) == cudaSuccess) { a.allocs.push_back(tmp); a.d_scalars = (LevelScalars*)tmp; cudaMemsetAsync(tmp, 0, sizeof(LevelScalars) * dims.size(), a.stream); } else if (!rc) rc = wl_fail(WL_ERR_CUDA, "cudaMalloc scalars");
  if (!rc && cudaMalloc(&tmp, sizeof(MiscScalars)) == cudaSuccess) { a.allocs.push_back(tmp); a.d_misc = (MiscScalars*)tmp; cudaMemsetAsync(tmp, 0, sizeof(MiscScalars), a.stream); } else if (!rc) rc = wl_fail(WL_ERR_CUDA, "cudaMalloc misc");
  if (!rc && cudaMalloc(&tmp, sizeof(double) * (WL_MAXBLOCKS * WL_MAXRED + 8)) == cudaSuccess) { a.allocs.push_back(tmp); a.ws.partials = (double*)tmp; } else if (!rc) rc = wl_fail(WL_ERR_CUDA, "cudaMalloc partials");
  if (!rc && cudaMalloc(&tmp, sizeof(unsigned int) * (dims.size() + 4)) == cudaSuccess) { a.allocs.push_back(tmp); a.d_nonuni = (unsigned int*)tmp; cudaMemsetAsync(tmp, 0, sizeof(unsigned int) * (dims.size() + 4), a.stream); } else if (!rc) rc = wl_fail(WL_ERR_CUDA, "cudaMalloc counters");
  { const char* e = getenv("WL_UNIFORM"); a.uniform_path = !(e && atoi(e) == 0); }
  { const char* e = getenv("WL_MARCH"); a.march_on = !(e && atoi(e) == 0); }
  { const char* e = getenv("WL_STAGE"); a.stage_on = !(e && atoi(e) == 0); }
  { const char* e = getenv("WL_RESJAC"); a.resjac_on = !(e && atoi(e) == 0); }
  { const char* e = getenv("WL_CFL_FUSE"); a.cflfuse_on = !(e && atoi(e) == 0); }
  { const char* e = getenv("WL_XDIV"); a.xdiv_on = e && atoi(e) != 0; }
  { const char* e = getenv("WL_RESJAC_DIV"); a.resjac_div = e && atoi(e) != 0; }
  { const char* e = getenv("WL_BDIM_FUSE"); a.bdim_fuse = !(e && atoi(e) == 0); }
  if (!rc && cudaMalloc(&tmp, 64) == cudaSuccess) { a.allocs.push_back(tmp); a.ws.ticket = (unsigned int*)tmp; cudaMemsetAsync(tmp, 0, 64, a.stream); } else if (!rc) rc = wl_fail(WL_ERR_CUDA, "cudaMalloc ticket");
  if (!rc && cudaMallocHost((void**)&a.h_pinned, 64 * sizeof(float)) != cudaSuccess) rc = wl_fail(WL_ERR_CUDA, "cudaMallocHost");
  if (rc) { wl_destroy((wl_sim*)sp); return rc; }
  alloc_fields(&rc);  // dry pass: measures
  a.arena_bytes = (a.arena_used + 255) / 256 * 256;
  {
    void* m = nullptr;
    cudaError_t e = cudaMalloc(&m, a.arena_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(m, 0, a.arena_bytes, a.stream);
    if (e != cudaSuccess) { rc = wl_cuda_check(a, e, "cudaMalloc (field arena)"); wl_destroy((wl_sim*)sp); return rc; }
    a.arena = (char*)m;
  }
  alloc_fields(&rc);
  if (!rc) rc = wl_cuda_check(a, cudaMemsetAsync(a.body_mask - a.g.lead, 1, (size_t)a.g.comp, a.stream), "mask init");
  // peer mapping of the arenas (slab ranks): halos and reductions then bypass NCCL
  if (!rc) rc = dist_p2p_setup(a);
  if (!rc) rc = vcycle_fused_setup(a);
  if (!rc) rc = convdiff_tma_setup(a);
  if (rc) { wl_destroy((wl_sim*)sp); return rc; }
  // μ₀ = 1, BC!(μ₀,0) (Flow.jl:118-119); then the Poisson constructors' set_diag!/restrictL!
  const double zero[3] = {0, 0, 0};
  rc = wl_fill(a, a.mu0 - a.g.lead, 1.f, a.g.comp * D);
  if (!rc) rc = bc_vec(a, a.g, a.mu0, zero, false, a.g.per);
  if (!rc) rc = pois_update(a);
  if (!rc) rc = wl_cuda_check(a, cudaStreamSynchronize(a.stream), "wl_create sync");
  if (rc) { wl_destroy((wl_sim*)sp); return rc; }
  a.dist.ready = true;
  *out = (wl_sim*)sp;
  return WL_OK;
}

int wl_create(wl_sim** out, int D, const int* N, double nu, double dt0, const int* perdir, int exitBC, int pois_kind, int device) {
  return create_impl(out, D, N, nu, dt0, perdir, exitBC, pois_kind, device, 0, 1, nullptr);
}
int wl_dist_unique_id(char* out128) { return out128 ? dist_unique_id(out128) : wl_fail(WL_ERR_ARG, "null buffer"); }
int wl_create_dist(wl_sim** out, int D, const int* N, double nu, double dt0, const int* perdir, int exitBC, int pois_kind, int device,
                   int rank, int nranks, const char* nccl_id128) {
  if (nranks > 1 && !nccl_id128) return wl_fail(WL_ERR_ARG, "wl_create_dist: NCCL id required");
  return create_impl(out, D, N, nu, dt0, perdir, exitBC, pois_kind, device, rank, nranks, nccl_id128);
}
// Host-only: the slab plan of rank `rank` — per level: distributed flag, local interior planes, global
// offset of local plane 0, halo flags of the lower/upper z face.  No GPU is touched (CPU-testable).
int wl_dist_plan(int D, const int* N, const int* perdir, int pois_kind, int rank, int nranks, int max_levels, int* nlevels,
                 int* distributed, int* nz_local, int* zoff, int* zlo_halo, int* zhi_halo) {
  if (!N || (D != 2 && D != 3)) return wl_fail(WL_ERR_ARG, "wl_dist_plan: D must be 2 or 3");
  const auto dims = level_dims(D, N, pois_kind);
  if (pois_kind == WL_POIS_MULTILEVEL && dims.size() <= 2) return wl_fail(WL_ERR_ARG, "MultiLevelPoisson requires size=a2ⁿ, where n>2");
  std::vector<int> ld;
  const int rc = plan_levels(D, N, nranks, dims, ld);
  if (rc) return rc;
  const int perz = (D == 3 && perdir && perdir[2]) ? 1 : 0;
  if (nlevels) *nlevels = (int)dims.size();
  for (int l = 0; l < (int)dims.size() && l < max_levels; ++l) {
    const bool d = nranks > 1 && ld[l];
    const int nzg = D == 3 ? dims[l][2] - 2 : 1;
    const int nzl = d ? nzg / nranks : nzg;
    if (distributed) distributed[l] = d ? 1 : 0;
    if (nz_local) nz_local[l] = nzl;
    if (zoff) zoff[l] = d ? rank * nzl : 0;
    if (zlo_halo) zlo_halo[l] = d && (rank > 0 || perz) ? 1 : 0;
    if (zhi_halo) zhi_halo[l] = d && (rank < nranks - 1 || perz) ? 1 : 0;
  }
  return WL_OK;
}
int wl_dist_finalize(void) { return dist_finalize(); }
int wl_dist_info(wl_sim* s, int* rank, int* nranks, int* zoff, int* nz_local, int64_t* collectives) {
  if (!s) return wl_fail(WL_ERR_ARG, "null handle");
  Sim& a = *(Sim*)s;
  if (rank) *rank = a.dist.rank;
  if (nranks) *nranks = a.dist.nranks;
  if (zoff) *zoff = a.g.zoff;
  if (nz_local) *nz_local = a.D == 3 ? a.g.n[2] - 2 : 1;
  if (collectives) *collectives = a.collectives;
  return WL_OK;
}

int wl_destroy(wl_sim* s) {
  if (!s) return WL_OK;
  Sim& a = S;
  cudaSetDevice(a.device);
  if (a.stream) cudaStreamSynchronize(a.stream);
  dist_destroy(a);
  vcycle_fused_destroy(a);
  convdiff_tma_destroy(a);
  if (a.arena) cudaFree(a.arena);
  if (a.stage) cudaFree(a.stage);
  for (void* p : a.allocs) cudaFree(p);
  if (a.h_pinned) cudaFreeHost(a.h_pinned);
  for (auto& pc : a.prof) for (cudaEvent_t e : pc.ev) cudaEventDestroy(e);
  if (a.own_stream && a.stream) cudaStreamDestroy(a.stream);
  delete &a;
  return WL_OK;
}
const char* wl_last_error(void) { return g_err.c_str(); }
int wl_set_stream(wl_sim* s, void* stream) {
  GUARD(s);
  cudaStreamSynchronize(S.stream);
  if (S.own_stream && S.stream) cudaStreamDestroy(S.stream);
  S.own_stream = false;
  S.stream = (cudaStream_t)stream;
  return WL_OK;
}
// every collective entry point ends here: a peer wait that timed out inside one of its kernels
// surfaces as an error of THIS call, not of some later wl_mom_step
int wl_sync(wl_sim* s) {
  GUARD(s);
  const int rc = wl_cuda_check(S, cudaStreamSynchronize(S.stream), "wl_sync");
  return rc ? rc : dist_check_p2p_error(S);
}

int wl_field_count(wl_sim* s, int field, int64_t* count) {
  GUARD(s);
  int nc; if (!field_ptr(S, field, &nc)) return wl_fail(WL_ERR_ARG, "bad field");
  *count = (int64_t)S.g.n[0] * S.g.n[1] * nz_padded_global(S, 0) * nc;
  return WL_OK;
}
int wl_upload(wl_sim* s, int field, const float* host, int64_t count) {
  GUARD(s);
  int nc; float* d = field_ptr(S, field, &nc);
  const int nzg = nz_padded_global(S, 0);
  if (!d || count != (int64_t)S.g.n[0] * S.g.n[1] * nzg * nc) return wl_fail(WL_ERR_ARG, "wl_upload: bad field or size");
  if (field == WL_V || field == WL_MU0 || field == WL_MU1) S.mask_dirty = true;
  if (field == WL_MU0) pois_invalidate_uniform(S, 0);  // levels[1].L ≡ μ₀: stale until update!
  int rc = 0;
  if (field == WL_P) { S.p_pend = 0; S.xdiv_valid = false; }              // replaced wholesale
  if (field == WL_F) S.f_stale = false;                                   // replaced wholesale
  else if (field == WL_U0 || field == WL_V) rc = mom_materialize_f(S);     // inputs of the on-demand f are about to change
  if (rc) return rc;
  rc = copy_field(S, S.g, d, const_cast<float*>(host), nc, true, nzg);
  // slab faces: refresh the halo planes from the neighbours (collective: every rank uploads the field)
  if (!rc) rc = dist_halo(S, S.g, d, nc, (field == WL_U || field == WL_U0) ? 2 : 1);
  return rc ? rc : wl_sync(s);
}
int wl_download(wl_sim* s, int field, float* host, int64_t count) {
  GUARD(s);
  int nc; float* d = field_ptr(S, field, &nc);
  const int nzg = nz_padded_global(S, 0);
  if (!d || count != (int64_t)S.g.n[0] * S.g.n[1] * nzg * nc) return wl_fail(WL_ERR_ARG, "wl_download: bad field or size");
  if (field == WL_F) { const int rc = mom_materialize_f(S); if (rc) return rc; }
  if (field == WL_P) { const int rc = mom_materialize_p(S); if (rc) return rc; }
  return copy_field(S, S.g, d, host, nc, false, nzg);
}
int wl_nlevels(wl_sim* s) { return s ? (int)S.levels.size() : 0; }
int wl_level_dims(wl_sim* s, int level, int* Ng) {
  GUARD(s);
  if (level < 0 || level >= (int)S.levels.size()) return wl_fail(WL_ERR_ARG, "bad level");
  for (int k = 0; k < S.D; ++k) Ng[k] = S.levels[level].g.n[k];
  if (S.D == 3) Ng[2] = nz_padded_global(S, level);  // always the global shape
  return WL_OK;
}
int wl_level_upload(wl_sim* s, int level, int lf, const float* host, int64_t count) {
  GUARD(s);
  int nc; float* d = level_ptr(S, level, lf, &nc);
  if (!d) return wl_fail(WL_ERR_ARG, "wl_level_upload: bad level/field");
  const GridD& g = S.levels[level].g;
  const int nzg = nz_padded_global(S, level);
  if (count != (int64_t)g.n[0] * g.n[1] * nzg * nc) return wl_fail(WL_ERR_ARG, "wl_level_upload: bad size");
  if (level == 0 && lf == WL_LVL_X) S.p_pend = 0;         // levels[1].x ≡ p, replaced wholesale
  if (level == 0) S.xdiv_valid = false;                   // x or ϵ may be the one replaced
  if (level == 0 && lf == WL_LVL_L) S.mask_dirty = true;  // levels[1].L ≡ μ₀
  if (lf == WL_LVL_L || lf == WL_LVL_D || lf == WL_LVL_ID) pois_invalidate_uniform(S, level);
  int rc = copy_field(S, g, d, const_cast<float*>(host), nc, true, nzg);
  if (!rc) rc = dist_halo(S, g, d, nc, 1);
  return rc ? rc : wl_sync(s);
}
int wl_level_download(wl_sim* s, int level, int lf, float* host, int64_t count) {
  GUARD(s);
  int nc; float* d = level_ptr(S, level, lf, &nc);
  if (!d) return wl_fail(WL_ERR_ARG, "wl_level_download: bad level/field");
  const GridD& g = S.levels[level].g;
  const int nzg = nz_padded_global(S, level);
  if (count != (int64_t)g.n[0] * g.n[1] * nzg * nc) return wl_fail(WL_ERR_ARG, "wl_level_download: bad size");
  if (level == 0 && lf == WL_LVL_X) { const int rc = mom_materialize_p(S); if (rc) return rc; }
  return copy_field(S, g, d, host, nc, false, nzg);
}

static void per_flags(Sim& a, const int* perdir, int* out) {
  for (int k = 0; k < 3; ++k) out[k] = perdir ? (k < a.D && perdir[k] != 0) : a.g.per[k];
}
int wl_bc(wl_sim* s, int field, const double* A, int saveexit, const int* perdir) {
  GUARD(s);
  int nc; float* d = field_ptr(S, field, &nc);
  if (!d || nc != S.D) return wl_fail(WL_ERR_ARG, "wl_bc: not a vector field");
  int per[3]; per_flags(S, perdir, per);
  if (field == WL_V || field == WL_MU0) S.mask_dirty = true;
  if (field == WL_MU0) pois_invalidate_uniform(S, 0);
  if (field == WL_F || field == WL_U0 || field == WL_V) { const int rc0 = mom_materialize_f(S); if (rc0) return rc0; }
  int rc = bc_vec(S, S.g, d, A, saveexit != 0, per, (field == WL_U || field == WL_U0) ? 2 : 1);
  return rc ? rc : wl_sync(s);
}
int wl_exit_bc(wl_sim* s, const double* U, double dt) {
  GUARD(s);
  int rc = bc_exit(S, U, (float)dt);
  return rc ? rc : wl_sync(s);
}
int wl_per_bc(wl_sim* s, int field, const int* perdir) {
  GUARD(s);
  int nc; float* d = field_ptr(S, field, &nc);
  if (!d || nc != 1) return wl_fail(WL_ERR_ARG, "wl_per_bc: not a scalar field");
  int per[3]; per_flags(S, perdir, per);
  int rc = field == WL_P ? mom_materialize_p(S) : 0;
  if (!rc) rc = bc_per(S, S.g, d, per);
  return rc ? rc : wl_sync(s);
}
int wl_conv_diff(wl_sim* s, int src) {
  GUARD(s);
  if (src != WL_U && src != WL_U0) return wl_fail(WL_ERR_ARG, "wl_conv_diff: src must be WL_U or WL_U0");
  S.f_stale = false;  // f is overwritten
  int rc = mom_conv_diff(S, src == WL_U ? S.u : S.u0, false, nullptr);
  return rc ? rc : wl_sync(s);
}
int wl_accelerate(wl_sim* s, const double* g) { GUARD(s); int rc = mom_materialize_f(S); if (!rc) rc = mom_accelerate(S, g); return rc ? rc : wl_sync(s); }
int wl_bdim(wl_sim* s) {
  GUARD(s);
  int rc = mom_materialize_f(S);
  if (!rc) rc = mom_bdim1(S);
  if (!rc) rc = mom_bdim2(S, true, 1.f);
  return rc ? rc : wl_sync(s);
}
int wl_scale_u(wl_sim* s, double scale) { GUARD(s); int rc = mom_scale_u(S, scale); return rc ? rc : wl_sync(s); }
int wl_project(wl_sim* s, double w, int* n) { GUARD(s); int rc = mom_materialize_p(S); if (!rc) rc = mom_project(S, w, n); return rc ? rc : wl_sync(s); }
int wl_cfl(wl_sim* s, double* dt) { GUARD(s); return mom_cfl(S, dt); }

int wl_pois_update(wl_sim* s) { GUARD(s); int rc = pois_update(S); return rc ? rc : wl_sync(s); }
int wl_pois_mult(wl_sim* s) { GUARD(s); int rc = mom_materialize_p(S);
  if (!rc) rc = pois_mult(S); return rc ? rc : wl_sync(s); }
int wl_pois_residual(wl_sim* s, int level, double* r2) {
  GUARD(s);
  if (level < 0 || level >= (int)S.levels.size()) return wl_fail(WL_ERR_ARG, "bad level");
  int rc = mom_materialize_p(S);
  if (!rc) rc = pois_residual(S, S.levels[level]);
  if (!rc && r2) rc = pois_dot_rr(S, S.levels[level], r2);
  return rc ? rc : wl_sync(s);
}
int wl_pois_jacobi(wl_sim* s, int level, int it) {
  GUARD(s);
  if (level < 0 || level >= (int)S.levels.size()) return wl_fail(WL_ERR_ARG, "bad level");
  int rc = mom_materialize_p(S);
  if (!rc) rc = pois_jacobi(S, S.levels[level], it);
  if (!rc) rc = bc_per(S, S.levels[level].g, S.levels[level].eps, S.levels[level].g.per);  // perBC!(ϵ) made visible
  return rc ? rc : wl_sync(s);
}
int wl_pois_pcg(wl_sim* s, int level, int it) {
  GUARD(s);
  if (level < 0 || level >= (int)S.levels.size()) return wl_fail(WL_ERR_ARG, "bad level");
  int rc = mom_materialize_p(S);
  if (!rc) rc = pois_pcg(S, level, it);
  if (!rc) rc = bc_per(S, S.levels[level].g, S.levels[level].eps, S.levels[level].g.per);
  return rc ? rc : wl_sync(s);
}
int wl_pois_head(wl_sim* s) {
  GUARD(s);
  if (S.pois_kind != WL_POIS_MULTILEVEL || S.levels.size() < 2) return wl_fail(WL_ERR_ARG, "MultiLevelPoisson with at least two levels required");
  int rc = mom_materialize_p(S);
  if (!rc) rc = pois_head(S);
  return rc ? rc : wl_sync(s);
}
int wl_pois_vcycle(wl_sim* s, int level) {
  GUARD(s);
  if (level < 0 || level + 1 >= (int)S.levels.size()) return wl_fail(WL_ERR_ARG, "bad level");
  int rc = mom_materialize_p(S);
  if (!rc) rc = pois_vcycle(S, level);
  if (!rc) rc = pois_materialize_eps(S, level);
  return rc ? rc : wl_sync(s);
}
int wl_pois_solve(wl_sim* s, double tol, int itmx, int* n) {
  GUARD(s);
  if (tol < 0) tol = 1e-4;
  if (itmx < 0) itmx = S.pois_kind == WL_POIS_MULTILEVEL ? 32 : 1000;
  int rc = mom_materialize_p(S);
  if (!rc) rc = pois_solve(S, tol, itmx, n);
  return rc ? rc : wl_sync(s);
}
// nds(body, loc(0,I,Float64), t) from the host → device (same re-pitched indexing as the fields, Float64)
static int metrics_nds(Sim& a, const double* nds, int64_t count) {
  if (!nds) return a.nds ? 0 : wl_fail(WL_ERR_STATE, "no nds array has been passed yet");
  const int nzg = nz_padded_global(a, 0);
  if (count != (int64_t)a.g.n[0] * a.g.n[1] * nzg * a.D) return wl_fail(WL_ERR_ARG, "nds: bad size");
  if (!a.nds) {
    void* p = nullptr;
    const size_t bytes = sizeof(double) * (size_t)a.g.comp * a.D;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, bytes, a.stream);
    if (e != cudaSuccess) return wl_cuda_check(a, e, "cudaMalloc (nds)");
    a.allocs.push_back(p);
    a.nds = (double*)p + a.g.lead;
  }
  return copy_field<double>(a, a.g, a.nds, const_cast<double*>(nds), a.D, true, nzg);
}
int wl_pressure_force(wl_sim* s, const double* nds, int64_t count, double* force) {
  GUARD(s);
  if (!force) return wl_fail(WL_ERR_ARG, "null output");
  int rc = mom_materialize_p(S);
  if (!rc) rc = metrics_nds(S, nds, count);
  return rc ? rc : metrics_run(S, 0, nullptr, force);
}
int wl_viscous_force(wl_sim* s, const double* nds, int64_t count, double* force) {
  GUARD(s);
  if (!force) return wl_fail(WL_ERR_ARG, "null output");
  int rc = metrics_nds(S, nds, count);
  return rc ? rc : metrics_run(S, 1, nullptr, force);
}
int wl_pressure_moment(wl_sim* s, const double* x0, const double* nds, int64_t count, double* moment) {
  GUARD(s);
  if (!moment || !x0) return wl_fail(WL_ERR_ARG, "null argument");
  int rc = mom_materialize_p(S);
  if (!rc) rc = metrics_nds(S, nds, count);
  return rc ? rc : metrics_run(S, 2, x0, moment);
}
int wl_derived(wl_sim* s, int kind, int comp, const double* U, const double* axis, const double* center) {
  GUARD(s);
  int rc = derived_run(S, kind, comp, U, axis, center);
  return rc ? rc : wl_sync(s);
}
int wl_mom_step(wl_sim* s, const double* U, const double* gp, const double* gc, double* dt_next, int* n_pred, int* n_corr) {
  GUARD(s);
  if (!U) return wl_fail(WL_ERR_ARG, "wl_mom_step: U is required");
  return mom_step(S, U, gp, gc, dt_next, n_pred, n_corr);
}
int wl_get_dt(wl_sim* s, double* dt) { GUARD(s); *dt = (double)S.dt.back(); return WL_OK; }
int wl_set_dt(wl_sim* s, double dt) { GUARD(s); S.dt.back() = (float)dt; return WL_OK; }
int64_t wl_launch_count(wl_sim* s) { return s ? S.launches : 0; }
int wl_profile(wl_sim* s, int on) {
  GUARD(s);
  S.profiling = on != 0;
  for (auto& p : S.prof) { p.used = 0; p.ms = 0; p.launches = 0; p.bytes = 0; p.d_counter = nullptr; }
  // pcg! kernels can be predicated off on the device: their traffic is counted from device counters
  LevelScalars* sc = S.levels[0].sc;
  const double cb = 4.0 * (double)S.g.interior;
  int rc = pois_uniform_fraction(S);  // algorithmic bytes depend on how many packs take the uniform-coefficient path
  if (rc) return rc;
  S.prof[WL_PROF_PCG_MULT].d_counter = &sc->cnt_mult;     S.prof[WL_PROF_PCG_MULT].bytes_per_active = wl_words(S, S.D + 3, S.D + 1) * cb;
  S.prof[WL_PROF_PCG_UPDATE].d_counter = &sc->cnt_update; S.prof[WL_PROF_PCG_UPDATE].bytes_per_active = wl_words(S, 7, 1) * cb;
  S.prof[WL_PROF_PCG_DIR].d_counter = &sc->cnt_dir;       S.prof[WL_PROF_PCG_DIR].bytes_per_active = wl_words(S, 4, 1) * cb;
  rc = wl_read_scalars(S, sc, sizeof(LevelScalars));
  if (rc) return rc;
  LevelScalars h; memcpy(&h, S.h_pinned, sizeof(h));
  S.prof[WL_PROF_PCG_MULT].counter_base = h.cnt_mult;
  S.prof[WL_PROF_PCG_UPDATE].counter_base = h.cnt_update;
  S.prof[WL_PROF_PCG_DIR].counter_base = h.cnt_dir;
  return WL_OK;
}
int wl_profile_read(wl_sim* s, int which, double* ms, int64_t* launches, double* bytes) {
  GUARD(s);
  if (which < 0 || which >= WL_PROF_N) return wl_fail(WL_ERR_ARG, "bad profile class");
  int rc = wl_cuda_check(S, cudaStreamSynchronize(S.stream), "profile sync");
  if (rc) return rc;
  ProfClass& p = S.prof[which];
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    float t = 0;
    if (cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]) == cudaSuccess) p.ms += t;
  }
  p.used = 0;
  if (p.d_counter) {  // active launches only
    if ((rc = wl_read_scalars(S, p.d_counter, sizeof(int)))) return rc;
    int c; memcpy(&c, S.h_pinned, sizeof(int));
    p.launches = c - p.counter_base;
    p.bytes = p.bytes_per_active * (double)p.launches;
  }
  if (ms) *ms = p.ms;
  if (launches) *launches = p.launches;
  if (bytes) *bytes = p.bytes;
  return WL_OK;
}

// development aid, not part of the header: phase log of the persistent solver kernel (WL_VC_TRACE=1)
int wl_debug_vc_trace(wl_sim* s, unsigned long long* out, int max_pairs, int* n) { GUARD(s); return vcycle_fused_trace(S, out, max_pairs, n); }
int wl_profile_count(void) { return WL_PROF_N; }
const char* wl_profile_name(int which) {
  static const char* names[WL_PROF_N] = {"conv_diff+bdim1", "bdim2", "pcg_mult", "pcg_update", "pcg_dir", "pcg_init",
                                         "increment", "residual", "project_misc", "vcycle_tail", "solver", "mom_step"};
  return (which >= 0 && which < WL_PROF_N) ? names[which] : "";
}

}  // extern "C"
