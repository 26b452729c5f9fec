// wl_vcycle.cu — the multigrid solver as ONE persistent cooperative kernel
// (reference src/MultiLevelPoisson.jl:70-99, src/Poisson.jl:91-144).
//
// Below the finest levels a V-cycle is pure latency: ~30 launches per level, each a few
// microseconds of work, plus (across slab ranks) a halo exchange or an all-reduce between most of
// them.  k_vcycle_tail runs in one launch either
//
//   tail mode  — levels lf … nl-1 of one V-cycle:
//       down:  Jacobi!(l) ; restrict!(l → l+1) ; fill!(x_{l+1},0)               l = lf … nl-2
//              pcg!(nl-1)
//       up:    prolongate!(l+1 → l) ; increment!(l) ; pcg!(l)                    l = nl-2 … lf
//   solve mode — the whole solver!(ml; tol, itmx): residual! (+ mean subtraction), then
//       { Jacobi!(1) ; restrict! ; the tail ; prolongate! ; increment! ; pcg!(1) ; L₂ test } until
//       converged, then perBC!(x); chosen when the finest level has ≤ 24 M cells per rank (WL_FUSED_SOLVE_CELLS).
//
// Phases are separated by a grid barrier (~1-2 µs) instead of a kernel boundary, every scalar
// decision of pcg! (|ρ|<10eps, |α|∉[1e-2,1e2]) and of solver! (L₂ < tol) is taken redundantly and
// identically by all blocks from the same reduced bits, the smallest levels (≤ 16³) are run by one
// block with __syncthreads as the barrier, and on z-distributed levels the halo planes and
// reduction partials go to the neighbours' memory directly from inside the kernel (wl_p2p.cuh).
// In solve mode the host reads nothing until the end of the time step.
//
// Arithmetic is that of the per-function kernels in wl_poisson.cu (same expressions, -fmad=false,
// Float64 dot accumulation); periodic ghosts of ϵ are not materialised — the stencil wraps its
// index instead, which reads the same value — and prolongate!/increment!/the head of pcg! run as
// one phase with ϵ evaluated from the coarse x on the fly (wl_vec.cuh).
#include <algorithm>
#include <cstdlib>

#include "wl_internal.h"
#include "wl_vec.cuh"

namespace {

#define VC_THREADS 256
#ifndef WL_VC_MINB
#define WL_VC_MINB 2  // resident blocks per SM the register allocation aims for
#endif

struct FusedParams {
  int D, nl, lf, lb;          // dims, levels, first fused level, first single-block level
  Lvl lv[WL_MAXLEVELS];
  int dist[WL_MAXLEVELS];     // level is z-distributed (halo exchange + all-reduce)
  int bxlog[WL_MAXLEVELS];    // log2 of the threads along x used for this level
  int nranks, rank;
  unsigned int* bar;          // grid barrier: {count, generation}
  double* partials;           // [gridDim.x][2]
  const P2PDev* p2p;          // null on one rank
  MiscScalars* misc;          // solve mode: V-cycle count out
  int nzg0, perz_global;      // interior z-planes of the whole finest level; the global domain is periodic in z
  // z-marching tiles (for_columns / march_staged, wl_vec.cuh) of the levels large enough to be bandwidth-bound:
  // march[l] = 1: tile_prolong runs on column tiles; 2: also the shared-memory staged stencil phases
  MarchGeom mg[WL_MAXLEVELS];
  int march[WL_MAXLEVELS];
  unsigned long long* trace;  // WL_VC_TRACE=1: block 0 logs (tag, globaltimer) after every phase (development aid), else null
};
// per-launch arguments of the whole-solver mode (solver!(ml), MultiLevelPoisson.jl:87-99)
struct SolveArgs {
  int solve;         // 0: coarse tail only (levels lf…), 1: the whole solver! including level 1 of the reference
  float tol;
  int itmx;
  int ghosts_valid;  // x's ghost cells / halo planes are consistent (project! scaled them with the interior)
  int has_ghost;     // z⋅ϵ has a ghost-shell part (z ≡ σ keeps conv_diff!'s stale Φ)
  int slot;          // which entry of MiscScalars::solve_n receives the count
  const float* divu; // non-null: residual! evaluates z = div(u) on the spot instead of reading z (project! inside mom_step!)
};

extern __shared__ __align__(16) float vc_ring[];  // operand ring of the staged stencil phases

struct Shared {
  double red[2][VC_THREADS / 32];
  double tot[2];
  double mail[WL_MAXRANKS][2];
  P2PDev c;
};

// state every thread carries (uniform across the grid)
struct Run {
  Shared* sh;
  unsigned long long push_epoch, ar_epoch;
  unsigned int bar_gen;
  int trn;
};
#define VC_TRACE_MAX 4096
// tag = 1000·level + phase: 1 residual, 2 sub-mean, 3 jacobi ϵ, 4 increment, 5 restrict, 6 prolong+increment+init, 7 pcg init,
// 10+3s pcg mult / 11+3s update / 12+3s dir of sweep s, 90 single-block sub-cycle, 95 perBC, 99 start/end
__device__ __forceinline__ void trace(const FusedParams& p, Run& run, int tag) {
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && run.trn < VC_TRACE_MAX) {
    p.trace[2 * run.trn] = (unsigned long long)tag;
    p.trace[2 * run.trn + 1] = p2p_now();
    ++run.trn;
  }
}

// ---- barriers -----------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(const FusedParams& p, Run& run) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int gen = run.bar_gen;
    __threadfence();
    if (atomicAdd(&p.bar[0], 1u) == gridDim.x - 1) {
      p.bar[0] = 0u;
      __threadfence();
      atomicAdd(&p.bar[1], 1u);
    } else {
      while (*((volatile unsigned int*)&p.bar[1]) == gen) {}
    }
    __threadfence();
  }
  ++run.bar_gen;
  __syncthreads();
}
template <bool GRID>
__device__ __forceinline__ void phase_sync(const FusedParams& p, Run& run) {
  if (GRID) grid_barrier(p, run);
  else __syncthreads();
}

// ---- iteration ----------------------------------------------------------------------------
// interior rows of level l, V cells per thread; f(i, j, k, I)
template <int V, bool GRID, class F>
__device__ __forceinline__ void rows(const FusedParams& p, int l, F&& f) {
  const GridD& g = p.lv[l].g;
  const int bxl = p.bxlog[l], bx = 1 << bxl, by = VC_THREADS >> bxl;
  const int tx = threadIdx.x & (bx - 1), ty = threadIdx.x >> bxl;
  const int nj = g.D > 1 ? g.n[1] - 2 : 1, nk = g.D > 2 ? g.n[2] - 2 : 1;
  const int njt = (nj + by - 1) / by;
  const int nunits = njt * nk;
  const int nxi = g.n[0] - 2;
  const int bid = GRID ? blockIdx.x : 0, nb = GRID ? gridDim.x : 1;
  for (int unit = bid; unit < nunits; unit += nb) {
    const int kk = unit / njt, jt = unit - kk * njt;
    const int j = 1 + jt * by + ty;
    if (j > nj) continue;
    const int k = g.D > 2 ? 1 + kk : 0;
    const int64_t row = j * g.s[1] + k * g.s[2];
    for (int i = 1 + V * tx; i <= nxi; i += V * bx) f(i, j, k, row + i);
  }
}

// ---- reductions ---------------------------------------------------------------------------
// Sum of NR per-thread Float64 accumulators over the whole level (all blocks, all ranks); every
// thread of every block returns the same totals.  Contains the phase barrier.
template <int NR, bool GRID>
__device__ __forceinline__ void reduce(const FusedParams& p, Run& run, bool distributed, double (&v)[NR]) {
  Shared& sh = *run.sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const double w = warp_sum(v[r]);
    if (lane == 0) sh.red[r][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double a = lane < VC_THREADS / 32 ? sh.red[r][lane] : 0.0;
      a = warp_sum(a);
      if (lane == 0) {
        if (GRID) p.partials[(size_t)blockIdx.x * 2 + r] = a;
        else sh.tot[r] = a;
      }
    }
  }
  if (!GRID) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; ++r) v[r] = sh.tot[r];
    __syncthreads();
    return;
  }
  grid_barrier(p, run);
  if (warp == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double a = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) a += ((volatile double*)p.partials)[(size_t)b * 2 + r];
      a = warp_sum(a);
      if (lane == 0) sh.tot[r] = a;
    }
  }
  __syncthreads();
  if (distributed) {
    // block 0 mails this rank's totals to every rank; every block adds the R mails in rank order
    const P2PDev& c = sh.c;
    const unsigned long long e = ++run.ar_epoch;
    const int slot = (int)(e & 1ull);
    if (tid < c.nranks) {
      if (blockIdx.x == 0) {
        P2PMail* m = &p2p_flags(c, tid)->mbox[slot][c.rank];
        ((volatile double*)m->v)[0] = sh.tot[0];
        ((volatile double*)m->v)[1] = NR > 1 ? sh.tot[1] : 0.0;
        __threadfence_system();
        p2p_st_release(&m->epoch, e);
      }
      const P2PMail* in = &p2p_flags(c, c.rank)->mbox[slot][tid];
      p2p_wait<true>(c, &in->epoch, e);
      sh.mail[tid][0] = ((const volatile double*)in->v)[0];
      sh.mail[tid][1] = ((const volatile double*)in->v)[1];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        double a = 0.0;
        for (int q = 0; q < c.nranks; ++q) a += sh.mail[q][r];
        sh.tot[r] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) v[r] = sh.tot[r];
  __syncthreads();
}

// ---- peer pushes from inside the kernel ------------------------------------------------------
__device__ __forceinline__ void post_ready(const FusedParams& p, Run& run) {
  // "all my reads of ghost data delivered by pushes ≤ push_epoch are complete" — to every peer
  if (!p.p2p || blockIdx.x != 0) return;
  const P2PDev& c = run.sh->c;
  if (threadIdx.x < c.nranks && threadIdx.x != c.rank) p2p_st_release(&p2p_flags(c, threadIdx.x)->ready[c.rank], run.push_epoch + 1);
}
struct Seg { int peer; int64_t src_off, dst_off, n4; };  // float offsets are converted to byte offsets by the caller
#define VC_CHUNK 1024  // float4 per chunk
// the caller has passed a grid barrier since the data was produced
__device__ __forceinline__ void push_segments(const FusedParams& p, Run& run, const Seg* seg, int nseg, unsigned int peer_mask) {
  const P2PDev& c = run.sh->c;
  P2PFlags* mine = p2p_flags(c, c.rank);
  const unsigned long long e = ++run.push_epoch;
  const int t = threadIdx.x;
  const bool mypeer = t < c.nranks && ((peer_mask >> t) & 1u);
  if (mypeer) p2p_wait<false>(c, &mine->ready[t], e);
  __syncthreads();
  int64_t nchunk = 0;
  for (int s = 0; s < nseg; ++s) nchunk += (seg[s].n4 + VC_CHUNK - 1) / VC_CHUNK;
  const int npart = (int)min((int64_t)gridDim.x, nchunk);
  if ((int)blockIdx.x < npart) {
    int64_t first = 0;
    for (int s = 0; s < nseg; ++s) {
      const int64_t nc = (seg[s].n4 + VC_CHUNK - 1) / VC_CHUNK;
      const float4* src = reinterpret_cast<const float4*>(c.base[c.rank] + seg[s].src_off);
      float4* dst = reinterpret_cast<float4*>(c.base[seg[s].peer] + seg[s].dst_off);
      int64_t k = ((int64_t)blockIdx.x - first % npart + npart) % npart;
      for (; k < nc; k += npart) {
        const int64_t lo = k * VC_CHUNK, hi = min(lo + (int64_t)VC_CHUNK, seg[s].n4);
        for (int64_t i = lo + t; i < hi; i += VC_THREADS) dst[i] = src[i];
      }
      first += nc;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool lastb;
    if (t == 0) lastb = (atomicAdd(&mine->ticket, 1u) == (unsigned int)npart - 1);
    __syncthreads();
    if (lastb) {
      __threadfence_system();
      if (mypeer) p2p_st_release(&p2p_flags(c, t)->done[c.rank], e);
      if (t == 0) mine->ticket = 0u;
    }
  }
  if (mypeer) p2p_wait<false>(c, &mine->done[t], e);
  __syncthreads();
}
// halo of a scalar field of distributed level l (one plane each way)
__device__ __forceinline__ void halo_push(const FusedParams& p, Run& run, int l, float* a) {
  const GridD& g = p.lv[l].g;
  const P2PDev& c = run.sh->c;
  const int R = c.nranks, me = c.rank, up = (me + 1) % R, dn = (me + R - 1) % R, n = g.n[2];
  Seg seg[2];
  int ns = 0;
  unsigned int mask = 0;
  const char* base = c.base[me];
  if (g.zhi_halo) {
    seg[ns++] = Seg{up, (int64_t)((const char*)(a + (int64_t)(n - 2) * g.s[2] - WL_OX) - base), (int64_t)((const char*)(a - WL_OX) - base), g.s[2] / 4};
    mask |= 1u << up;
  }
  if (g.zlo_halo) {
    seg[ns++] = Seg{dn, (int64_t)((const char*)(a + g.s[2] - WL_OX) - base), (int64_t)((const char*)(a + (int64_t)(n - 1) * g.s[2] - WL_OX) - base), g.s[2] / 4};
    mask |= 1u << dn;
  }
  push_segments(p, run, seg, ns, mask);
}

// ---- phases ----------------------------------------------------------------------------------
// increment! (Poisson.jl:99-103) of level l, ϵ already stored; on a distributed level ϵ's halo first
template <int D, int V, bool GRID>
__device__ __forceinline__ void increment(const FusedParams& p, Run& run, int l, float* r2 = nullptr) {
  const Lvl& L = p.lv[l];
  if (GRID && p.dist[l]) halo_push(p, run, l, L.eps);
  double acc[1] = {0.0};
  if (GRID && D == 3 && V == 4 && p.march[l] == 2) {
    for_columns_even(L.g, p.mg[l], [&](int i, int j, int k0, int k1) { tile_increment(L, p.mg[l], vc_ring, i, j, k0, k1, acc[0]); });
  } else
  rows<V, GRID>(p, l, [&](int i, int j, int k, int64_t I) {
    const Pack<V> ec = ldp<V>(L.eps + I), ae = mult_w<D, V>(L, L.eps, i, j, k, I, ec, pack_uniform<V>(L, I));
    Pack<V> r = ldp<V>(L.r + I), x = ldp<V>(L.x + I);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      r.v[q] = r.v[q] - ae.v[q];
      x.v[q] = x.v[q] + ec.v[q];
      acc[0] = dfma(r.v[q], r.v[q], acc[0]);
    }
    stp<V>(L.r + I, r);
    stp<V>(L.x + I, x);
  });
  if (r2) {  // L₂ (Poisson.jl:146) is only ever read on the finest level
    reduce<1, GRID>(p, run, GRID && p.dist[l], acc);
    *r2 = (float)acc[0];
  } else {
    phase_sync<GRID>(p, run);
  }
  if (GRID && p.dist[l]) post_ready(p, run);
  trace(p, run, 1000 * l + 4);
}
// Jacobi!(l; it=1) (Poisson.jl:110-113)
template <int D, int V, bool GRID>
__device__ __forceinline__ void jacobi(const FusedParams& p, Run& run, int l, bool eps_ready = false) {
  const Lvl& L = p.lv[l];
  if (eps_ready) { increment<D, V, GRID>(p, run, l); return; }  // residual! already stored ϵ = r·iD
  rows<V, GRID>(p, l, [&](int, int, int, int64_t I) {
    const Pack<V> r = ldp<V>(L.r + I), id = ld_iD<V>(L, I, pack_uniform<V>(L, I));
    Pack<V> e;
#pragma unroll
    for (int q = 0; q < V; ++q) e.v[q] = r.v[q] * id.v[q];
    stp<V>(L.eps + I, e);
  });
  phase_sync<GRID>(p, run);
  trace(p, run, 1000 * l + 3);
  increment<D, V, GRID>(p, run, l);
}
// pcg!(l; it=6) (Poisson.jl:123-144)
// `shell`: add Σ over the global ghost shell of z·ϵ to z⋅ϵ (level 1 of the reference only);
// `r2`: keep L₂ = r⋅r up to date (finest level only)
template <int D, int V, bool GRID>
__device__ __forceinline__ void pcg(const FusedParams& p, Run& run, int l, int it, bool shell = false, float* r2 = nullptr,
                                    const float* rho_in = nullptr) {
  const Lvl& L = p.lv[l];
  const bool dist = GRID && p.dist[l];
  float rho;
  if (rho_in) {  // ϵ = r·iD and ρ = r⋅ϵ came out of the fused prolongate!/increment! pass
    rho = *rho_in;
  } else {
    double acc[1] = {0.0};
    rows<V, GRID>(p, l, [&](int, int, int, int64_t I) {
      const Pack<V> r = ldp<V>(L.r + I), id = ld_iD<V>(L, I, pack_uniform<V>(L, I));
      Pack<V> e;
#pragma unroll
      for (int q = 0; q < V; ++q) { e.v[q] = r.v[q] * id.v[q]; acc[0] = dfma(r.v[q], e.v[q], acc[0]); }
      stp<V>(L.eps + I, e);
    });
    reduce<1, GRID>(p, run, dist, acc);
    rho = (float)acc[0];
    trace(p, run, 1000 * l + 7);
  }
  if (fabsf(rho) < 10 * WL_EPS32) return;
  for (int i = 1; i <= it; ++i) {
    if (dist) halo_push(p, run, l, L.eps);
    float alpha;
    {
      double acc[1] = {0.0};
      if (GRID && D == 3 && V == 4 && p.march[l] == 2) {
        for_columns_even(L.g, p.mg[l], [&](int ii, int j, int k0, int k1) { tile_pcg_mult(L, p.mg[l], vc_ring, ii, j, k0, k1, acc[0]); });
      } else
      rows<V, GRID>(p, l, [&](int ii, int j, int k, int64_t I) {
        const Pack<V> ec = ldp<V>(L.eps + I), z = mult_w<D, V>(L, L.eps, ii, j, k, I, ec, pack_uniform<V>(L, I));
#pragma unroll
        for (int q = 0; q < V; ++q) acc[0] = dfma(z.v[q], ec.v[q], acc[0]);
        stp<V>(L.z + I, z);
      });
      if (shell) {
        const GridD& g = L.g;
        for_global_ghosts(g, blockIdx.x * (int64_t)VC_THREADS + threadIdx.x, (int64_t)gridDim.x * VC_THREADS, [&](const int (&c)[3]) {
          int w[3] = {c[0], c[1], c[2]};
          for (int d = 0; d < g.D; ++d)
            if (g.per[d]) w[d] = c[d] == 0 ? g.n[d] - 2 : (c[d] == g.n[d] - 1 ? 1 : c[d]);
          acc[0] += (double)L.z[c[0] + c[1] * g.s[1] + c[2] * g.s[2]] * (double)L.eps[w[0] + w[1] * g.s[1] + w[2] * g.s[2]];
        });
      }
      reduce<1, GRID>(p, run, dist, acc);
      if (dist) post_ready(p, run);
      alpha = rho / (float)acc[0];
      trace(p, run, 1000 * l + 10 + 3 * i);
    }
    const double aa = fabs((double)alpha);
    if (aa < 1e-2 || aa > 1e2) return;
    float rho2;
    {
      double acc[2] = {0.0, 0.0};
      rows<V, GRID>(p, l, [&](int, int, int, int64_t I) {
        const Pack<V> e = ldp<V>(L.eps + I), z = ldp<V>(L.z + I), id = ld_iD<V>(L, I, pack_uniform<V>(L, I));
        Pack<V> x = ldp<V>(L.x + I), r = ldp<V>(L.r + I);
#pragma unroll
        for (int q = 0; q < V; ++q) {
          x.v[q] += alpha * e.v[q];
          r.v[q] -= alpha * z.v[q];
          acc[0] = dfma(r.v[q], (r.v[q] * id.v[q]), acc[0]);
          acc[1] = dfma(r.v[q], r.v[q], acc[1]);
        }
        stp<V>(L.x + I, x);
        stp<V>(L.r + I, r);
      });
      if (i == it && !r2) { phase_sync<GRID>(p, run); trace(p, run, 1000 * l + 11 + 3 * i); return; }  // the dots of the last sweep are never used below level 1
      reduce<2, GRID>(p, run, dist, acc);
      trace(p, run, 1000 * l + 11 + 3 * i);
      if (r2) *r2 = (float)acc[1];
      if (i == it) return;
      rho2 = (float)acc[0];
    }
    if (fabsf(rho2) < 10 * WL_EPS32) return;
    const float beta = rho2 / rho;
    rho = rho2;
    rows<V, GRID>(p, l, [&](int, int, int, int64_t I) {
      const Pack<V> r = ldp<V>(L.r + I), id = ld_iD<V>(L, I, pack_uniform<V>(L, I));
      Pack<V> e = ldp<V>(L.eps + I);
#pragma unroll
      for (int q = 0; q < V; ++q) e.v[q] = beta * e.v[q] + r.v[q] * id.v[q];
      stp<V>(L.eps + I, e);
    });
    phase_sync<GRID>(p, run);
    trace(p, run, 1000 * l + 12 + 3 * i);
  }
}
// restrict!(l → l+1) (MultiLevelPoisson.jl:33) + fill!(x_{l+1},0) (:75); a distributed → replicated
// transition computes this rank's coarse planes and pushes them to every peer
template <int D, bool GRID>
__device__ __forceinline__ void restrict_to(const FusedParams& p, Run& run, int l) {
  const Lvl& f = p.lv[l];
  const Lvl& c = p.lv[l + 1];
  const bool gather = GRID && p.dist[l] && !p.dist[l + 1];
  int ck0 = 1, ck1 = c.g.n[2] - 2, fshift = 0;
  if (gather) {
    const int cp = (c.g.n[2] - 2) / p.nranks;
    ck0 = 1 + p.rank * cp; ck1 = ck0 + cp - 1; fshift = -f.g.zoff;
  }
  rows<1, GRID>(p, l + 1, [&](int i0, int j0, int k0, int64_t I) {
    c.x[I] = 0.f;
    if (D > 2 && (k0 < ck0 || k0 > ck1)) return;
    const int cc[3] = {i0, j0, k0};
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < D; ++d) { lo[d] = 2 * cc[d] - 1; hi[d] = 2 * cc[d]; }
    if (D > 2) { lo[2] += fshift; hi[2] += fshift; }
    float s = 0.f;
    if (D == 3) {
      // the two x-cells of a fine row are one aligned 8-byte load (fine x = 2·i0−1 sits 2·i0+2 floats into a 16-byte aligned
      // row); four independent loads, then the additions in the reference's order (x fastest, then y, then z)
      const float* b = f.r + lo[0] + (int64_t)lo[1] * f.g.s[1] + (int64_t)lo[2] * f.g.s[2];
      const float2 v00 = *reinterpret_cast<const float2*>(b), v01 = *reinterpret_cast<const float2*>(b + f.g.s[1]);
      const float2 v10 = *reinterpret_cast<const float2*>(b + f.g.s[2]), v11 = *reinterpret_cast<const float2*>(b + f.g.s[1] + f.g.s[2]);
      s += v00.x; s += v00.y; s += v01.x; s += v01.y; s += v10.x; s += v10.y; s += v11.x; s += v11.y;
    } else {
      for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
          for (int x = lo[0]; x <= hi[0]; ++x) s += f.r[x + y * f.g.s[1] + z * f.g.s[2]];
    }
    c.r[I] = s;
  });
  phase_sync<GRID>(p, run);
  if (gather) {
    const P2PDev& pc = run.sh->c;
    const int64_t count = (int64_t)((c.g.n[2] - 2) / p.nranks) * c.g.s[2];
    const char* base = pc.base[pc.rank];
    const float* mine = c.r + c.g.s[2] + (int64_t)p.rank * count - WL_OX;
    Seg seg[WL_MAXRANKS];
    int ns = 0;
    unsigned int mask = 0;
    for (int r = 0; r < p.nranks; ++r) {
      if (r == p.rank) continue;
      const int64_t off = (int64_t)((const char*)mine - base);
      seg[ns++] = Seg{r, off, off, count / 4};
      mask |= 1u << r;
    }
    push_segments(p, run, seg, ns, mask);
  }
  trace(p, run, 1000 * l + 5);
}
// prolongate!(l+1 → l) + increment!(l) + the head of pcg!(l) in one phase (wl_vec.cuh); across slab faces
// ϵ is the neighbour's coarse x, so the halo that travels is the (4× smaller) coarse one
template <int D, int V, bool GRID>
__device__ __forceinline__ void prolong_inc_init(const FusedParams& p, Run& run, int l, float& rho, float* r2) {
  const Lvl& f = p.lv[l];
  const Lvl& c = p.lv[l + 1];
  const bool fdist = GRID && p.dist[l], cdist = GRID && p.dist[l + 1];
  if (fdist && cdist) halo_push(p, run, l + 1, c.x);
  ProMap m;
#pragma unroll
  for (int d = 0; d < 3; ++d) { m.per[d] = f.g.per[d]; m.n[d] = f.g.n[d]; }
  m.zmode = (fdist && !cdist) ? 1 : 0;
  m.zoff = f.g.zoff;
  m.nzg = p.nzg0 >> l;
  m.perz_global = p.perz_global;
  double acc[2] = {0.0, 0.0};
  if (GRID && D == 3 && V == 4 && p.march[l] >= 1) {
    for_columns_even(f.g, p.mg[l], [&](int i, int j, int k0, int k1) { tile_prolong(f, c, m, p.mg[l], f.x, i, j, k0, k1, acc[0], acc[1]); });
  } else {
    rows<V, GRID>(p, l, [&](int i, int j, int k, int64_t I) { prolong_increment_init<D, V>(f, c, m, f.x, i, j, k, I, acc[0], acc[1]); });
  }
  reduce<2, GRID>(p, run, fdist, acc);
  if (fdist && cdist) post_ready(p, run);
  rho = (float)acc[0];
  if (r2) *r2 = (float)acc[1];
  trace(p, run, 1000 * l + 6);
}

template <int D, bool GRID>
__device__ __forceinline__ void down_level(const FusedParams& p, Run& run, int l) {
  if ((p.lv[l].g.n[0] - 2) % 4 == 0 && GRID) jacobi<D, 4, GRID>(p, run, l);
  else jacobi<D, 1, GRID>(p, run, l);
  restrict_to<D, GRID>(p, run, l);
}
template <int D, bool GRID>
__device__ __forceinline__ void pcg_level(const FusedParams& p, Run& run, int l) {
  if ((p.lv[l].g.n[0] - 2) % 4 == 0 && GRID) pcg<D, 4, GRID>(p, run, l, 6);
  else pcg<D, 1, GRID>(p, run, l, 6);
}
template <int D, bool GRID>
__device__ __forceinline__ void up_level(const FusedParams& p, Run& run, int l) {
  float rho;
  if ((p.lv[l].g.n[0] - 2) % 4 == 0 && GRID) {
    prolong_inc_init<D, 4, GRID>(p, run, l, rho, nullptr);
    pcg<D, 4, GRID>(p, run, l, 6, false, nullptr, &rho);
  } else {
    prolong_inc_init<D, 1, GRID>(p, run, l, rho, nullptr);
    pcg<D, 1, GRID>(p, run, l, 6, false, nullptr, &rho);
  }
}

// levels lf … nl-1 of one V-cycle (the residual of level lf has just been restricted)
template <int D>
__device__ __forceinline__ void vcycle_tail(const FusedParams& p, Run& run, int lf) {
  const int last = p.nl - 1;
  const int lb = max(p.lb, lf);  // levels ≥ lb: block 0 alone
  // down-leg on the grid
  for (int l = lf; l < min(lb, last); ++l) down_level<D, true>(p, run, l);
  if (lb <= last) {
    // whole sub-cycle of the smallest levels in one block
    if (blockIdx.x == 0) {
      for (int l = lb; l < last; ++l) down_level<D, false>(p, run, l);
      pcg_level<D, false>(p, run, last);
      for (int l = last - 1; l >= lb; --l) up_level<D, false>(p, run, l);
    }
    grid_barrier(p, run);
    trace(p, run, 1000 * lb + 90);
  } else {
    pcg_level<D, true>(p, run, last);
  }
  for (int l = min(lb, last) - 1; l >= lf; --l) {
    // the replicated tail is finished: its gathered residual may be overwritten by the next V-cycle
    if (p.p2p && p.dist[l] && !p.dist[l + 1]) post_ready(p, run);
    up_level<D, true>(p, run, l);
  }
}

// ---- finest level (whole-solver mode) -------------------------------------------------------
// residual! (Poisson.jl:91-97) + the ϵ = r·iD of the Jacobi! that opens the first V-cycle
template <int D, int V>
__device__ __forceinline__ void residual0(const FusedParams& p, Run& run, const SolveArgs& sa, float& r2) {
  const Lvl& L = p.lv[0];
  const bool dist = p.dist[0] != 0;
  // x's periodic neighbours are read through wrapped indices; slab faces need the neighbours' planes
  if (dist && !sa.ghosts_valid) halo_push(p, run, 0, L.x);
  double acc[2] = {0.0, 0.0};
  if (D == 3 && V == 4 && p.march[0] == 2) {
    for_columns_even(L.g, p.mg[0], [&](int i, int j, int k0, int k1) { tile_residual(L, p.mg[0], vc_ring, i, j, k0, k1, 1, sa.divu, acc[0], acc[1]); });
  } else
  rows<V, true>(p, 0, [&](int i, int j, int k, int64_t I) {
    const bool uni = pack_uniform<V>(L, I);
    const Pack<V> xc = ldp<V>(L.x + I), ax = mult_w<D, V>(L, L.x, i, j, k, I, xc, uni), zc = sa.divu ? div_pack<D, V>(L.g, sa.divu, I) : ldp<V>(L.z + I),
                  id = ld_iD<V>(L, I, uni);
    Pack<V> r, e;
#pragma unroll
    for (int q = 0; q < V; ++q) {
      r.v[q] = id.v[q] == 0.f ? 0.f : zc.v[q] - ax.v[q];
      acc[0] += (double)r.v[q];
      acc[1] = dfma(r.v[q], r.v[q], acc[1]);
      e.v[q] = r.v[q] * id.v[q];
    }
    stp<V>(L.r + I, r);
    stp<V>(L.eps + I, e);
  });
  reduce<2, true>(p, run, dist, acc);
  if (dist && !sa.ghosts_valid) post_ready(p, run);
  const float rmean = (float)acc[0] / (float)L.g.interior_global;
  r2 = (float)acc[1];
  trace(p, run, 1);
  if (fabsf(rmean) > 2 * WL_EPS32) {
    rows<V, true>(p, 0, [&](int, int, int, int64_t I) {
      Pack<V> r = ldp<V>(L.r + I);
      const Pack<V> id = ld_iD<V>(L, I, pack_uniform<V>(L, I));
      Pack<V> e;
#pragma unroll
      for (int q = 0; q < V; ++q) { r.v[q] = r.v[q] - rmean; e.v[q] = r.v[q] * id.v[q]; }
      stp<V>(L.r + I, r);
      stp<V>(L.eps + I, e);
    });
    grid_barrier(p, run);
  }
}
// perBC!(x) (MultiLevelPoisson.jl:97): periodic ghost cells, then the slab halo planes
template <int D>
__device__ __forceinline__ void per_bc_x0(const FusedParams& p, Run& run) {
  const Lvl& L = p.lv[0];
  const GridD& g = L.g;
  bool any_per = false;
  for (int d = 0; d < D; ++d) any_per |= g.per[d] != 0;
  if (any_per) {
    for_global_ghosts(g, blockIdx.x * (int64_t)VC_THREADS + threadIdx.x, (int64_t)gridDim.x * VC_THREADS, [&](const int (&c0)[3]) {
      int c[3] = {c0[0], c0[1], c0[2]};
      bool written = false;
      for (int j = D - 1; j >= 0; --j) {
        if (!g.per[j]) continue;
        if (c[j] == 0) { c[j] = g.n[j] - 2; written = true; }
        else if (c[j] == g.n[j] - 1) { c[j] = 1; written = true; }
      }
      if (written) L.x[c0[0] + c0[1] * g.s[1] + c0[2] * g.s[2]] = L.x[c[0] + c[1] * g.s[1] + c[2] * g.s[2]];
    });
  }
  if (p.dist[0]) {
    grid_barrier(p, run);
    halo_push(p, run, 0, L.x);
  }
}

// Two modes, one flow (so that the coarse tail is instantiated once):
//  tail mode  — levels lf … nl-1 of one V-cycle (the finer levels run as separate kernels);
//  solve mode — solver!(ml; tol, itmx) (MultiLevelPoisson.jl:87-99) entirely on the device: residual!,
//               then {Vcycle!, pcg!(finest), L₂ test} until converged, then perBC!(x); the V-cycle count
//               and L₂ go to device memory, the host reads them when it needs them.
template <int D, int MINB>
__global__ void __launch_bounds__(VC_THREADS, MINB) k_vcycle_tail(FusedParams p, SolveArgs sa) {
  __shared__ Shared sh;
  Run run;
  run.sh = &sh;
  run.push_epoch = run.ar_epoch = 0;
  if (p.p2p) {
    if (threadIdx.x == 0) sh.c = *p.p2p;
    __syncthreads();
    const P2PFlags* mine = p2p_flags(sh.c, sh.c.rank);
    run.push_epoch = *((const volatile unsigned long long*)&mine->push_epoch);
    run.ar_epoch = *((const volatile unsigned long long*)&mine->ar_epoch);
    post_ready(p, run);  // every earlier kernel of this rank has retired
  }
  run.bar_gen = *((volatile unsigned int*)&p.bar[1]);
  run.trn = 0;
  // every block must have sampled the generation / epochs before anyone moves them
  grid_barrier(p, run);
  trace(p, run, 99);
  const bool solve = sa.solve != 0;
  const bool v4 = (p.lv[0].g.n[0] - 2) % 4 == 0;
  const int lf = solve ? 1 : p.lf;
  float r2 = 0.f;
  int n = 0;
  if (solve) {
    if (v4) residual0<D, 4>(p, run, sa, r2);
    else residual0<D, 1>(p, run, sa, r2);
  }
  for (;;) {
    if (solve) {  // Vcycle!(ml) on the finest level: Jacobi!, restrict!
      if (v4) jacobi<D, 4, true>(p, run, 0, n == 0);
      else jacobi<D, 1, true>(p, run, 0, n == 0);
      restrict_to<D, true>(p, run, 0);
    }
    vcycle_tail<D>(p, run, lf);
    if (!solve) break;
    if (p.p2p && p.dist[0] && !p.dist[1]) post_ready(p, run);
    float rho;
    if (v4) {
      prolong_inc_init<D, 4, true>(p, run, 0, rho, &r2);
      pcg<D, 4, true>(p, run, 0, 6, sa.has_ghost != 0, &r2, &rho);
    } else {
      prolong_inc_init<D, 1, true>(p, run, 0, rho, &r2);
      pcg<D, 1, true>(p, run, 0, 6, sa.has_ghost != 0, &r2, &rho);
    }
    ++n;
    if (r2 < sa.tol || n >= sa.itmx) break;
  }
  if (solve) {
    per_bc_x0<D>(p, run);
    trace(p, run, 95);
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) { p.trace[2 * VC_TRACE_MAX] = (unsigned long long)run.trn; }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      p.lv[0].sc->r2 = r2;
      p.misc->solve_n[sa.slot] = n;
    }
  }
  if (p.p2p && blockIdx.x == 0 && threadIdx.x == 0) {
    P2PFlags* mine = p2p_flags(sh.c, sh.c.rank);
    mine->push_epoch = run.push_epoch;
    mine->ar_epoch = run.ar_epoch;
  }
}

}  // namespace

// Host side: the fused coarse tail replaces pois_vcycle(level ≥ lf-1)'s recursion.
struct VcycleFused {
  FusedParams p;
  size_t smem = 0;   // dynamic shared memory: the operand ring of the staged stencil phases
  int grid = 0;
  int minb = 2;
  bool ok = false;
  bool solve = false;  // the whole solver! runs in the kernel (level 1 of the reference included)
};
// MINB = resident blocks per SM the register allocation aims for: 2 (128 registers, no spills) is best for the latency-bound
// coarse tail; the whole-solver mode also streams the finest level, where 3 blocks (80 registers) hide HBM latency better
// (256³ per rank: 5.30 → 4.96 ms/step)
static const void* tail_kernel(int D, int minb) {
  if (D == 2) return (const void*)k_vcycle_tail<2, 2>;
  return minb >= 3 ? (const void*)k_vcycle_tail<3, 3> : (const void*)k_vcycle_tail<3, 2>;
}

static int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; }

int vcycle_fused_setup(Sim& s) {
  s.vc_first = -1;
  const char* env = getenv("WL_FUSED_VCYCLE");
  if (env && atoi(env) == 0) return 0;
  if (s.pois_kind != WL_POIS_MULTILEVEL || s.levels.size() < 2 || s.levels.size() > WL_MAXLEVELS) return 0;
  if (s.dist.nranks > 1 && !s.dist.p2p) return 0;  // distributed levels need the peer-memory path
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s.device);
  if (!coop) return 0;
  VcycleFused* vf = new VcycleFused;
  FusedParams& p = vf->p;
  p.D = s.D;
  p.nl = (int)s.levels.size();
  p.nranks = s.dist.nranks;
  p.rank = s.dist.rank;
  p.p2p = s.dist.nranks > 1 ? s.dist.dev : nullptr;
  p.misc = s.d_misc;
  p.nzg0 = s.Nz_global;
  p.perz_global = s.per_z_global;
  // fuse from the first level below the finest whose kernels are launch-latency bound
  const char* e1 = getenv("WL_FUSED_MAX_CELLS");
  const int64_t max_cells = e1 ? atoll(e1) : (int64_t)1 << 22;
  int lf = 1;
  while (lf < p.nl - 1 && s.levels[lf].g.interior > max_cells) ++lf;
  p.lf = lf;
  const char* e2 = getenv("WL_FUSED_BLOCK_CELLS");
  // measured (profiles/r02_solver_kernel_phase_trace_256.txt): one block needs 7–18 µs per phase on a 16³ level, the whole grid
  // 4–5 µs including its barrier; from 8³ down the single block wins
  const int64_t blk_cells = e2 ? atoll(e2) : 512;
  int lb = p.nl;
  for (int l = p.nl - 1; l >= lf; --l) {
    const bool dist = s.dist.nranks > 1 && s.lvl_dist[l];
    if (dist || s.levels[l].g.interior > blk_cells) break;
    lb = l;
  }
  p.lb = lb;
  for (int l = 0; l < p.nl; ++l) {
    p.lv[l] = s.levels[l];
    p.dist[l] = (s.dist.nranks > 1 && s.lvl_dist[l]) ? 1 : 0;
    const int nxi = s.levels[l].g.n[0] - 2;
    const bool grid_mode = l < lb;
    const int V = (grid_mode && nxi % 4 == 0) ? 4 : 1;
    int bx = 8;
    while (bx < nxi / V && bx < 128) bx <<= 1;
    p.bxlog[l] = ilog2(bx);
  }
  // whole-solver mode: one launch and no host read per solver! when the finest level is small enough per rank that the
  // launch and host-read latencies of the per-level kernels outweigh their higher streaming bandwidth (measured at 512³,
  // round 2: 8 ranks, 16.8 M cells each: 4.22 vs 4.39 ms/step; 4 ranks, 33.5 M cells: 6.27 vs 6.01; 2 ranks, 67 M cells: 11.3 vs 10.1)
  const char* e5 = getenv("WL_FUSED_SOLVE_CELLS");
  const int64_t solve_cells = e5 ? atoll(e5) : (int64_t)24 << 20;
  vf->solve = s.levels[0].g.interior <= solve_cells;
  const char* e4 = getenv("WL_FUSED_MINB");
  vf->minb = e4 ? atoi(e4) : (vf->solve ? 3 : 2);
  // z-marching tiles for the levels that are bandwidth-bound (3-D, grid mode, large enough)
  const char* e6 = getenv("WL_FUSED_MARCH_CELLS");
  const int64_t march_cells = e6 ? atoll(e6) : (int64_t)1 << 20;
  vf->smem = 0;
  for (int l = 0; l < p.nl; ++l) {
    p.march[l] = 0;
    p.mg[l] = MarchGeom{};
    const MarchLaunch ml = march_launch(s.levels[l].g, s.nsm);
    if (!s.march_on || l >= lb || !ml.ok || s.levels[l].g.interior < march_cells) continue;
    p.mg[l] = ml.mg;
    p.march[l] = (ml.staged && s.stage_on) ? 2 : 1;
    if (p.march[l] == 2) vf->smem = std::max(vf->smem, ml.smem);
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tail_kernel(s.D, vf->minb), VC_THREADS, vf->smem);
  if (e != cudaSuccess || per_sm < 1) { cudaGetLastError(); delete vf; return 0; }
  const char* e3 = getenv("WL_FUSED_BLOCKS_PER_SM");
  const int want = e3 ? atoi(e3) : 4;
  vf->grid = s.nsm * std::max(1, std::min(per_sm, want));
  void* tmp = nullptr;
  if (cudaMalloc(&tmp, 64) != cudaSuccess) { delete vf; return wl_fail(WL_ERR_CUDA, "cudaMalloc barrier"); }
  s.allocs.push_back(tmp);
  cudaMemsetAsync(tmp, 0, 64, s.stream);
  p.bar = (unsigned int*)tmp;
  if (cudaMalloc(&tmp, sizeof(double) * 2 * vf->grid) != cudaSuccess) { delete vf; return wl_fail(WL_ERR_CUDA, "cudaMalloc partials"); }
  s.allocs.push_back(tmp);
  p.partials = (double*)tmp;
  p.trace = nullptr;
  if (getenv("WL_VC_TRACE") && atoi(getenv("WL_VC_TRACE"))) {
    if (cudaMalloc(&tmp, sizeof(unsigned long long) * (2 * VC_TRACE_MAX + 2)) == cudaSuccess) { s.allocs.push_back(tmp); p.trace = (unsigned long long*)tmp; cudaMemsetAsync(tmp, 0, 8 * (2 * VC_TRACE_MAX + 2), s.stream); }
  }
  vf->ok = true;
  s.vc_fused = vf;
  s.vc_first = lf;
  s.vc_solve = vf->solve;
  return 0;
}
void vcycle_fused_destroy(Sim& s) {
  delete (VcycleFused*)s.vc_fused;
  s.vc_fused = nullptr;
}
// runs levels lf … nl-1 of the V-cycle (the residual of level lf has just been restricted)
// the uniform-pack masks are (in)validated after setup: the kernel gets the current pointers
static void refresh_masks(Sim& s, VcycleFused* vf) {
  for (int l = 0; l < vf->p.nl; ++l) vf->p.lv[l].umask = s.levels[l].umask;
}
int vcycle_fused_run(Sim& s) {
  VcycleFused* vf = (VcycleFused*)s.vc_fused;
  refresh_masks(s, vf);
  SolveArgs sa{};
  void* args[] = {&vf->p, &sa};
  wl_prof_begin(s, WL_PROF_VCTAIL);
  cudaError_t e = cudaLaunchCooperativeKernel(tail_kernel(s.D, vf->minb), dim3(vf->grid), dim3(VC_THREADS), args, vf->smem, s.stream);
  ++s.launches;
  wl_prof_end(s, WL_PROF_VCTAIL, 0.0);
  return wl_cuda_check(s, e, "k_vcycle_tail");
}

// solver!(ml; tol, itmx) as ONE launch; the V-cycle count lands in MiscScalars::solve_n[slot]
int vcycle_fused_solve(Sim& s, double tol, int itmx, bool ghosts_valid, int slot, const float* divu) {
  VcycleFused* vf = (VcycleFused*)s.vc_fused;
  refresh_masks(s, vf);
  const Lvl& l = s.levels[0];
  bool ghosts = l.g.zlo_halo || l.g.zhi_halo;
  for (int k = 0; k < s.D; ++k) ghosts |= l.g.per[k] != 0;
  SolveArgs sa{};
  sa.solve = 1;
  sa.tol = (float)tol;
  sa.itmx = itmx;
  sa.ghosts_valid = ghosts_valid ? 1 : 0;
  sa.has_ghost = ghosts ? 1 : 0;
  sa.slot = slot;
  sa.divu = divu;
  void* args[] = {&vf->p, &sa};
  cudaError_t e = cudaLaunchCooperativeKernel(tail_kernel(s.D, vf->minb), dim3(vf->grid), dim3(VC_THREADS), args, vf->smem, s.stream);
  ++s.launches;
  return wl_cuda_check(s, e, "k_vcycle_tail (solver!)");
}

// development aid (WL_VC_TRACE=1): the phase log of the last whole-solver launch → out[2n] = tag, out[2n+1] = globaltimer ns
int vcycle_fused_trace(Sim& s, unsigned long long* out, int max_pairs, int* n) {
  VcycleFused* vf = (VcycleFused*)s.vc_fused;
  *n = 0;
  if (!vf || !vf->p.trace) return 0;
  unsigned long long cnt = 0;
  cudaStreamSynchronize(s.stream);
  cudaMemcpy(&cnt, vf->p.trace + 2 * VC_TRACE_MAX, sizeof(cnt), cudaMemcpyDeviceToHost);
  const int m = (int)std::min<unsigned long long>(cnt, (unsigned long long)std::min(max_pairs, VC_TRACE_MAX));
  cudaMemcpy(out, vf->p.trace, sizeof(unsigned long long) * 2 * m, cudaMemcpyDeviceToHost);
  *n = m;
  return 0;
}
