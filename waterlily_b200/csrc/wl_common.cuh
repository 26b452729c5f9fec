// wl_common.cuh — shared device/host definitions of libwaterlily_b200.so (sm_100a only).
//
// Data layout in HBM: every field keeps the reference's padded, x-fastest,
// component-last (SoA) layout, but rows are re-pitched so that the first interior cell of each
// row is 16-byte aligned (float4 loads) — cell (i,j,k) (0-based, ghosts at 0 and n-1) of a scalar
// field lives at  base + i + j*s1 + k*s2,  base = alloc + OX, s1 = pitch (multiple of 4),
// s2 = s1*n[1].  Component c of a vector field starts at c*comp floats.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wl_p2p.cuh"

#define WL_OX 3  // row offset: interior cell i=1 lands on a 16-byte boundary

struct GridD {
  int D;
  int n[3];        // padded size (1 for dims >= D)
  int per[3];      // periodic flags
  int64_t s[3];    // strides in floats: 1, pitch, pitch*n[1]
  int64_t comp;    // floats per scalar field (allocation, incl. WL_OX and row padding)
  int64_t interior;  // product of (n-2) over the D dims (this rank's slab)
  int64_t interior_global;  // the same over the whole domain
  // z-slab decomposition: a z-boundary flagged `halo` is served by a neighbour exchange instead of
  // a boundary condition (interior slab faces, and the periodic wrap when there are several ranks)
  int zlo_halo, zhi_halo;
  int zoff;        // global z index of local plane 0 (global k = zoff + local k)
  int64_t lead;    // floats between the start of an allocation and cell (0,0,0): WL_OX + one spare z-plane in 3-D
};

// device-resident solver scalars of one Poisson level (src/Poisson.jl:123-144 keeps these on the
// host; here every data-dependent decision of pcg! is taken on the device)
struct LevelScalars {
  double raw[2];    // local reduction totals awaiting a cross-rank all-reduce (deferred finalisation)
  double ghost_dot; // Σ over the ghost shell of z·ϵ (level 1 only: z ≡ σ holds stale Φ)
  float rho;        // r⋅z
  float alpha;
  float beta;
  float r2;         // L₂ = r⋅r, refreshed by every kernel that writes r
  float rsum;       // mean of r in residual!
  int active;       // pcg! still iterating
  int need_sub;     // residual!: |s| > 2eps → subtract the mean
  int cnt_mult, cnt_update, cnt_dir;  // pcg! kernels that actually ran (not predicated off)
};

struct Lvl {
  GridD g;
  float *L, *Dg, *iD, *x, *eps, *r, *z;
  LevelScalars* sc;
  // Uniform-coefficient packs.  Away from bodies and non-periodic boundaries every face coefficient of a
  // level has the same value uL (μ₀ = 1 on the finest level, MultiLevelPoisson.jl:10-16 scales it by
  // 2^(D-2) per coarsening), so D = uD and iD = uiD there.  umask holds one byte per aligned pack of four
  // x-cells — 0 when all 2D faces of all four cells equal uL — written by set_diag! (k_set_diag) from the
  // actual contents of L.  A stencil on such a pack uses the three constants instead of loading L (D words),
  // D and iD: same operands, same operations, same bits, 5 words per cell less traffic.  Null when the
  // mask is stale (L, D or iD changed without update!) or the level's rows are not a multiple of four.
  const unsigned char* umask;
  float uL, uD, uiD;
  int upad;
};

// device scratch scalars shared by the momentum kernels
struct MiscScalars {
  double raw[2];
  float exit_mean_minus_U;  // ∮u of exitBC!
  float cfl_max;            // max over padded σ
  int solve_n[2];           // V-cycle counts of the predictor / corrector solver! (whole-solver kernel)
};

// What to do with a finished reduction.  On one GPU the last block of the producing kernel runs
// `finalize` itself; with several ranks the totals are parked in raw[], all-reduced by NCCL, and a
// one-thread kernel runs the same `finalize` on every rank (identical bits → identical decisions).
enum FinStage {
  FIN_RESIDUAL = 0,     // t0 = Σr, t1 = Σr²   → rsum, need_sub, r2
  FIN_R2,               // t0 = Σr²            → r2
  FIN_R2_IF_SUB,        // same, only if residual! subtracted the mean
  FIN_PCG_INIT,         // t0 = r⋅z            → rho, active
  FIN_INC_INIT,         // t0 = r⋅z, t1 = Σr²  → rho, active, r2   (increment! fused with the head of pcg!)
  FIN_GHOST_DOT,        // t0 = Σ_shell z·ϵ    → ghost_dot
  FIN_PCG_MULT,         // t0 = z⋅ϵ (interior) → alpha, active
  FIN_PCG_UPDATE,       // t0 = r⋅z, t1 = Σr²  → r2, beta, rho, active
  FIN_PCG_UPDATE_LAST,  // same, last sweep    → r2, active = 0
  FIN_EXIT,             // t0 = Σu on the exit plane → exit_mean_minus_U
  FIN_CFL,              // t0 = max σ (interior)
  FIN_CFL_SHELL,        // t0 = max σ (ghost shell), folded into cfl_max
  FIN_DOT              // t0 → *out
};
struct FinArgs {
  LevelScalars* sc;
  MiscScalars* misc;
  double* raw;      // where deferred totals are parked
  double* out;
  double count;     // global cell count the stage divides by
  double aux;       // U[1] for FIN_EXIT
  const P2PDev* p2p; // peer-memory context (defer == 2)
  int defer;        // 0: finalise locally; 1: park the totals for an NCCL all-reduce + k_finalize;
                    // 2: the finishing thread all-reduces over peer memory and finalises in place
  int has_ghost;    // this level's z⋅ϵ has a ghost-shell part (level 1 only)
};

// scratch for deterministic two-stage reductions: per-block partials + a ticket
struct ReduceWS {
  double* partials;   // [maxBlocks * WL_MAXRED]
  unsigned int* ticket;
};
#define WL_MAXLEVELS 12
#define WL_MAXRED 4
#define WL_MAXBLOCKS 4096

#ifdef __CUDACC__

#define WL_EPS32 1.1920929e-07f

__device__ __forceinline__ bool fin_predicated_off(int stage, const FinArgs& fa) {
  if (stage == FIN_R2_IF_SUB) return !fa.sc->need_sub;
  return (stage == FIN_GHOST_DOT || stage == FIN_PCG_MULT || stage == FIN_PCG_UPDATE || stage == FIN_PCG_UPDATE_LAST) && !fa.sc->active;
}
__device__ __forceinline__ void finalize(int stage, const FinArgs& fa, double t0, double t1) {
  LevelScalars* sc = fa.sc;
  switch (stage) {
    case FIN_RESIDUAL: {  // Poisson.jl:94-95
      const float s = (float)t0 / (float)fa.count;
      sc->rsum = s;
      sc->need_sub = fabsf(s) > 2 * WL_EPS32 ? 1 : 0;
      sc->r2 = (float)t1;
    } break;
    case FIN_R2:
    case FIN_R2_IF_SUB: sc->r2 = (float)t0; break;
    case FIN_PCG_INIT: {  // Poisson.jl:126-127
      const float rho = (float)t0;
      sc->rho = rho;
      sc->active = fabsf(rho) < 10 * WL_EPS32 ? 0 : 1;
    } break;
    case FIN_INC_INIT: {
      const float rho = (float)t0;
      sc->r2 = (float)t1;
      sc->rho = rho;
      sc->active = fabsf(rho) < 10 * WL_EPS32 ? 0 : 1;
    } break;
    case FIN_GHOST_DOT: sc->ghost_dot = t0; break;  // single GPU only; deferred mode adds it in FIN_PCG_MULT
    case FIN_PCG_MULT: {  // Poisson.jl:131-132
      ++sc->cnt_mult;
      // deferred (multi-rank): t1 is the all-reduced ghost-shell part, parked by k_shell_dot in raw[1]
      const float ze = (float)(t0 + (fa.defer ? (fa.has_ghost ? t1 : 0.0) : sc->ghost_dot));
      const float alpha = sc->rho / ze;
      sc->alpha = alpha;
      const double aa = fabs((double)alpha);
      if (aa < 1e-2 || aa > 1e2) sc->active = 0;
    } break;
    case FIN_PCG_UPDATE:
    case FIN_PCG_UPDATE_LAST: {  // Poisson.jl:135-141
      sc->r2 = (float)t1;
      ++sc->cnt_update;
      if (stage == FIN_PCG_UPDATE_LAST) {
        sc->active = 0;
      } else {
        const float rho2 = (float)t0;
        if (fabsf(rho2) < 10 * WL_EPS32) {
          sc->active = 0;
        } else {
          sc->beta = rho2 / sc->rho;
          sc->rho = rho2;
        }
      }
    } break;
    case FIN_EXIT: {  // util.jl:220
      const float mean = (float)t0 / (float)fa.count;
      fa.misc->exit_mean_minus_U = (float)((double)mean - fa.aux);
    } break;
    case FIN_CFL: fa.misc->cfl_max = (float)t0; break;
    case FIN_CFL_SHELL: fa.misc->cfl_max = fmaxf(fa.misc->cfl_max, (float)t0); break;
    case FIN_DOT: *fa.out = t0; break;
  }
}
// called by the thread that holds the grid totals
__device__ __forceinline__ void finish_reduction(int stage, const FinArgs& fa, double t0, double t1) {
  if (!fa.defer) { finalize(stage, fa, t0, t1); return; }
  if (stage == FIN_GHOST_DOT) { fa.raw[1] = t0; return; }        // rides along with the z⋅ϵ all-reduce
  if (fa.defer == 2) {
    double v0 = t0, v1 = stage == FIN_PCG_MULT ? fa.raw[1] : t1;
    p2p_allreduce2(*fa.p2p, v0, v1, stage == FIN_CFL || stage == FIN_CFL_SHELL);
    finalize(stage, fa, v0, v1);
    return;
  }
  fa.raw[0] = t0;
  if (stage != FIN_PCG_MULT) fa.raw[1] = t1;                      // keep the parked ghost part
}

// acc + a·b for the canonical Float64 dots: the product of two Float32 values is exact in Float64 (48 significant
// bits), so the fused multiply-add rounds once exactly where the separate exact multiply + add would — same bits,
// one DFMA instead of DMUL + DADD (the dots are a fifth of the instructions of the streaming kernels)
__device__ __forceinline__ double dfma(float a, float b, double acc) { return fma((double)a, (double)b, acc); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-level sum of NR per-thread Float64 accumulators (warp shuffles, then one smem hop),
// followed by a grid-level combine through per-block partials: the LAST block to arrive (ticket)
// adds the partials in a fixed order in double and returns true on its thread 0 with the totals.
// Reductions are canonical: products and sums in Float64, rounded to Float32 once by the caller
// (see oracle/wl_oracle.cpp "reductions" for why).  All threads of the block must call it.
template <int NR, bool MAX = false>
__device__ __forceinline__ bool grid_reduce(const double (&v)[NR], ReduceWS ws, double (&tot)[NR]) {
  __shared__ double sm[NR][32];
  __shared__ bool last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, warp = tid >> 5, nwarp = (nthr + 31) >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const double w = MAX ? warp_max(v[r]) : warp_sum(v[r]);
    if (lane == 0) sm[r][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double a = lane < nwarp ? sm[r][lane] : (MAX ? -1e300 : 0.0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double b = __shfl_xor_sync(0xffffffffu, a, o);
        a = MAX ? fmax(a, b) : a + b;
      }
      if (lane == 0) ws.partials[(size_t)blockIdx.x * WL_MAXRED + r] = a;
    }
  }
  if (tid == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(ws.ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  // last block: fixed-order combine of gridDim.x partials
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double a = MAX ? -1e300 : 0.0;
    for (unsigned int b = tid; b < gridDim.x; b += nthr) {
      const double p = ((volatile double*)ws.partials)[(size_t)b * WL_MAXRED + r];
      a = MAX ? fmax(a, p) : a + p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double b = __shfl_xor_sync(0xffffffffu, a, o);
      a = MAX ? fmax(a, b) : a + b;
    }
    __syncthreads();
    if (lane == 0) sm[r][warp] = a;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double a = MAX ? -1e300 : 0.0;
      for (int w = 0; w < nwarp; ++w) a = MAX ? fmax(a, sm[r][w]) : a + sm[r][w];
      tot[r] = a;
    }
    *ws.ticket = 0u;  // re-arm for the next launch on this stream
    return true;
  }
  return false;
}

#endif  // __CUDACC__

#ifdef __CUDACC__
// Walk the cell box [lo,hi] (0-based, inclusive, per dim) with a persistent grid: tiles of
// blockDim cells, grid-stride over tiles, threadIdx.x along x (coalesced rows).
struct Box { int lo[3], hi[3]; };
__host__ __device__ inline Box box_interior(const GridD& g) {
  Box b;
  for (int k = 0; k < 3; ++k) { b.lo[k] = k < g.D ? 1 : 0; b.hi[k] = k < g.D ? g.n[k] - 2 : 0; }
  return b;
}
__host__ __device__ inline Box box_all(const GridD& g) {
  Box b;
  for (int k = 0; k < 3; ++k) { b.lo[k] = 0; b.hi[k] = g.n[k] - 1; }
  return b;
}
__host__ inline int64_t box_tiles(const Box& b, dim3 blk) {
  const int e[3] = {b.hi[0] - b.lo[0] + 1, b.hi[1] - b.lo[1] + 1, b.hi[2] - b.lo[2] + 1};
  const int bd[3] = {(int)blk.x, (int)blk.y, (int)blk.z};
  int64_t n = 1;
  for (int k = 0; k < 3; ++k) n *= (e[k] + bd[k] - 1) / bd[k];
  return n;
}
template <class F>
__device__ __forceinline__ void for_cells(const GridD& g, const Box& b, F&& f) {
  const int ntx = (b.hi[0] - b.lo[0] + blockDim.x) / blockDim.x;
  const int nty = (b.hi[1] - b.lo[1] + blockDim.y) / blockDim.y;
  const int ntz = (b.hi[2] - b.lo[2] + blockDim.z) / blockDim.z;
  const int ntiles = ntx * nty * ntz;  // 32-bit on purpose: 64-bit div/mod per tile costs ~100 instructions
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tx = t % ntx;
    const int q = t / ntx;
    const int ty = q % nty, tz = q / nty;
    const int i = b.lo[0] + tx * blockDim.x + threadIdx.x;
    const int j = b.lo[1] + ty * blockDim.y + threadIdx.y;
    const int k = b.lo[2] + tz * blockDim.z + threadIdx.z;
    if (i <= b.hi[0] && j <= b.hi[1] && k <= b.hi[2]) f(i, j, k, (int64_t)i + j * g.s[1] + k * g.s[2]);
  }
}
// unique enumeration of the boundary shell of thickness t (cells with any coordinate in [0,t) or
// [n-t,n)): the last dim's 2t planes in full, then the next dim's planes restricted to the core of
// the dims above it, ...   t=1 is the ghost layer.
__device__ __forceinline__ bool shell_decode(const GridD& g, int64_t idx, int c[3], int t = 1) {
  c[0] = c[1] = c[2] = 0;
  for (int d = g.D - 1; d >= 0; --d) {
    int64_t below = 1, above = 1;
    for (int k = 0; k < d; ++k) below *= g.n[k];
    for (int k = d + 1; k < g.D; ++k) above *= g.n[k] - 2 * t;
    const int64_t plane = below * above, cnt = 2 * t * plane;
    if (idx < cnt) {
      const int pl = (int)(idx / plane);
      c[d] = pl < t ? pl : g.n[d] - 2 * t + pl;
      int64_t q = idx - (int64_t)pl * plane;
      for (int k = 0; k < d; ++k) { c[k] = (int)(q % g.n[k]); q /= g.n[k]; }
      for (int k = d + 1; k < g.D; ++k) { c[k] = t + (int)(q % (g.n[k] - 2 * t)); q /= (g.n[k] - 2 * t); }
      return true;
    }
    idx -= cnt;
  }
  return false;
}
// a cell of the local ghost shell is a ghost cell of the global domain owned by this rank unless it
// lies in a z-plane that belongs to a neighbouring slab (there it is the neighbour's cell)
__device__ __forceinline__ bool is_global_ghost(const GridD& g, const int c[3]) {
  if (g.D > 2 && ((c[2] == 0 && g.zlo_halo) || (c[2] == g.n[2] - 1 && g.zhi_halo))) return false;
  return true;
}
__host__ __device__ inline int64_t shell_count(const GridD& g, int t = 1) {
  int64_t all = 1, in = 1;
  for (int k = 0; k < g.D; ++k) { all *= g.n[k]; in *= g.n[k] - 2 * t; }
  return all - in;
}
// ---- the ghost shell (t = 1) in 3-D, fast: 32-bit arithmetic, x fastest, and only the cells this rank owns as ghost cells of
// the GLOBAL domain — a z-plane that is a slab halo is not enumerated at all (shell_decode + is_global_ghost walk it and
// discard it, with 64-bit divisions per cell: in the O(N²) kernels that was most of their time).
// Groups: the owned z-planes in full; the two y-planes over the local interior planes; the two x-planes over the interior
// rows of those planes — the same set, each cell once.
struct Shell3 { unsigned int pz, py, px, total; int nzp, zpl[2]; };
__host__ __device__ inline Shell3 shell3(const GridD& g) {
  Shell3 s;
  s.pz = (unsigned int)g.n[0] * (unsigned int)g.n[1];
  s.py = (unsigned int)g.n[0] * (unsigned int)(g.n[2] - 2);
  s.px = (unsigned int)(g.n[1] - 2) * (unsigned int)(g.n[2] - 2);
  s.nzp = 0;
  s.zpl[0] = s.zpl[1] = 0;
  if (!g.zlo_halo) s.zpl[s.nzp++] = 0;
  if (!g.zhi_halo) s.zpl[s.nzp++] = g.n[2] - 1;
  s.total = (unsigned int)s.nzp * s.pz + 2u * s.py + 2u * s.px;
  return s;
}
__device__ __forceinline__ void shell3_decode(const GridD& g, const Shell3& s, unsigned int idx, int c[3]) {
  const unsigned int nz = (unsigned int)s.nzp * s.pz;
  if (idx < nz) {
    const unsigned int pl = idx >= s.pz ? 1u : 0u, q = idx - pl * s.pz, row = q / (unsigned int)g.n[0];
    c[0] = (int)(q - row * (unsigned int)g.n[0]); c[1] = (int)row; c[2] = s.zpl[pl];
    return;
  }
  idx -= nz;
  if (idx < 2u * s.py) {
    const unsigned int pl = idx >= s.py ? 1u : 0u, q = idx - pl * s.py, k = q / (unsigned int)g.n[0];
    c[0] = (int)(q - k * (unsigned int)g.n[0]); c[1] = pl ? g.n[1] - 1 : 0; c[2] = 1 + (int)k;
    return;
  }
  idx -= 2u * s.py;
  const unsigned int pl = idx >= s.px ? 1u : 0u, q = idx - pl * s.px, m = (unsigned int)(g.n[1] - 2), k = q / m;
  c[0] = pl ? g.n[0] - 1 : 0; c[1] = 1 + (int)(q - k * m); c[2] = 1 + (int)k;
}
// f(c) for every ghost cell of the global domain this rank owns; thread `first` of `stride` (grid-stride or block-stride)
template <class F>
__device__ __forceinline__ void for_global_ghosts(const GridD& g, int64_t first, int64_t stride, F&& f) {
  if (g.D == 3) {
    const Shell3 s = shell3(g);
    for (int64_t idx = first; idx < (int64_t)s.total; idx += stride) {
      int c[3];
      shell3_decode(g, s, (unsigned int)idx, c);
      f(c);
    }
  } else {
    const int64_t cnt = shell_count(g);
    for (int64_t idx = first; idx < cnt; idx += stride) {
      int c[3];
      if (shell_decode(g, idx, c) && is_global_ghost(g, c)) f(c);
    }
  }
}
// (host) how many cells that is
inline int64_t global_ghost_count(const GridD& g) { return g.D == 3 ? (int64_t)shell3(g).total : shell_count(g); }
// threads enumerate, for every direction j, the planes c[j] ∈ {0,1,n-1} in full (edge cells are
// visited more than once and receive the same value each time)
__device__ __forceinline__ bool planes_decode(const GridD& g, int64_t idx, int nplanes, int c[3]) {
  c[0] = c[1] = c[2] = 0;
  for (int j = 0; j < g.D; ++j) {
    int64_t other = 1;
    for (int k = 0; k < g.D; ++k) if (k != j) other *= g.n[k];
    const int64_t cnt = nplanes * other;
    if (idx < cnt) {
      const int pl = (int)(idx / other);
      int64_t q = idx % other;
      for (int k = 0; k < g.D; ++k) {
        if (k == j) continue;
        c[k] = (int)(q % g.n[k]);
        q /= g.n[k];
      }
      c[j] = nplanes == 3 ? (pl == 0 ? 0 : (pl == 1 ? 1 : g.n[j] - 1)) : (pl == 0 ? 0 : g.n[j] - 1);
      return true;
    }
    idx -= cnt;
  }
  return false;
}
__host__ __device__ inline int64_t planes_count(const GridD& g, int nplanes) {
  int64_t tot = 0;
  for (int j = 0; j < g.D; ++j) {
    int64_t other = 1;
    for (int k = 0; k < g.D; ++k) if (k != j) other *= g.n[k];
    tot += nplanes * other;
  }
  return tot;
}

#endif
