// wl_vec.cuh — row-vectorised iteration for the HBM-bound field kernels.
//
// Every field row starts 16-byte aligned at its first interior cell (WL_OX), so when the interior
// row length is a multiple of 4 a thread owns 4 consecutive cells and moves them with one
// 128-bit load/store per field (V=4); coarse multigrid levels with odd row lengths use V=1.
// A block is (BX,BY): BX threads sweep a row (stride V*BX), BY rows per block; blocks stride over
// (row-group, plane) units, so the only integer divisions are one 32-bit pair per unit — the
// per-cell 64-bit div/mod of a flat index cost more than the stencil itself (profiles/r01).
#pragma once
#include "wl_common.cuh"

#ifdef __CUDACC__
template <int V> struct Pack { float v[V]; };

template <int V> __device__ __forceinline__ Pack<V> ldp(const float* __restrict__ p);
template <> __device__ __forceinline__ Pack<4> ldp<4>(const float* __restrict__ p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  Pack<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ Pack<1> ldp<1>(const float* __restrict__ p) { Pack<1> r; r.v[0] = *p; return r; }

template <int V> __device__ __forceinline__ void stp(float* __restrict__ p, const Pack<V>& a);
template <> __device__ __forceinline__ void stp<4>(float* __restrict__ p, const Pack<4>& a) {
  *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
template <> __device__ __forceinline__ void stp<1>(float* __restrict__ p, const Pack<1>& a) { *p = a.v[0]; }

// values at x-1 / x+1 of each lane, given the centre pack loaded from p
template <int V> __device__ __forceinline__ Pack<V> ld_xm(const float* __restrict__ p, const Pack<V>& c) {
  Pack<V> r; r.v[0] = p[-1];
#pragma unroll
  for (int l = 1; l < V; ++l) r.v[l] = c.v[l - 1];
  return r;
}
template <int V> __device__ __forceinline__ Pack<V> ld_xp(const float* __restrict__ p, const Pack<V>& c) {
  Pack<V> r;
#pragma unroll
  for (int l = 0; l < V - 1; ++l) r.v[l] = c.v[l + 1];
  r.v[V - 1] = p[V];
  return r;
}

// div(I,u) (Flow.jl:11-17) of V cells — project!'s z = div u evaluated where residual! consumes it
template <int D, int V>
__device__ __forceinline__ Pack<V> div_pack(const GridD& g, const float* __restrict__ u, int64_t I) {
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
  return s;
}

// ---- uniform-coefficient packs (see Lvl in wl_common.cuh) -------------------------------------------
// true when the pack of V=4 cells starting at I has L ≡ uL on all its faces (then D ≡ uD, iD ≡ uiD)
template <int V>
__device__ __forceinline__ bool pack_uniform(const Lvl& l, int64_t I) {
  if (V != 4 || l.umask == nullptr) return false;
  return l.umask[(I + l.g.lead) >> 2] == 0;
}
template <int V>
__device__ __forceinline__ Pack<V> splat(float v) {
  Pack<V> r;
#pragma unroll
  for (int q = 0; q < V; ++q) r.v[q] = v;
  return r;
}
// iD of a pack: the constant on a uniform pack, a load otherwise
template <int V>
__device__ __forceinline__ Pack<V> ld_iD(const Lvl& l, int64_t I, bool uni) {
  return uni ? splat<V>(l.uiD) : ldp<V>(l.iD + I);
}

// mult (Poisson.jl:69-75) for V cells at (i..i+V-1, j, k); periodic neighbours are read through a
// wrapped index instead of a ghost copy.  `uni`: the pack is uniform — L, D come from constants.
template <int D, int V>
__device__ __forceinline__ Pack<V> mult_w(const Lvl& l, const float* x, int i, int j, int k, int64_t I, const Pack<V>& xc, bool uni = false) {
  const GridD& g = l.g;
  const Pack<V> Dc = uni ? splat<V>(l.uD) : ldp<V>(l.Dg + I);
  Pack<V> s;
#pragma unroll
  for (int q = 0; q < V; ++q) s.v[q] = xc.v[q] * Dc.v[q];
  {
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(l.uL); Lp = Lc; }
    else { Lc = ldp<V>(l.L + I); Lp = ld_xp<V>(l.L + I, Lc); }
    Pack<V> xm, xp;
    xm.v[0] = x[(g.per[0] && i == 1) ? I + (g.n[0] - 3) : I - 1];
#pragma unroll
    for (int q = 1; q < V; ++q) xm.v[q] = xc.v[q - 1];
#pragma unroll
    for (int q = 0; q < V - 1; ++q) xp.v[q] = xc.v[q + 1];
    xp.v[V - 1] = x[(g.per[0] && i + V - 1 == g.n[0] - 2) ? I + (V - 1) - (g.n[0] - 3) : I + V];
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
#pragma unroll
  for (int d = 1; d < D; ++d) {
    const int64_t st = g.s[d];
    const int cd = d == 1 ? j : k, n = g.n[d];
    const int64_t om = (g.per[d] && cd == 1) ? (int64_t)(n - 3) * st : -st;
    const int64_t op = (g.per[d] && cd == n - 2) ? -(int64_t)(n - 3) * st : st;
    const float* Li = l.L + d * g.comp;
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(l.uL); Lp = Lc; }
    else { Lc = ldp<V>(Li + I); Lp = ldp<V>(Li + I + st); }
    const Pack<V> xm = ldp<V>(x + I + om), xp = ldp<V>(x + I + op);
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
  return s;
}

// mult (Poisson.jl:69-75) of one pack in 3-D with every neighbour value supplied by the caller (registers or
// shared memory): xm1/xp4 are the x-neighbours of the pack's ends, ym/yp/zm/zp the neighbour packs.  Operands
// and operation order are those of mult_w.
template <int V>
__device__ __forceinline__ Pack<V> mult_vals(const Lvl& l, int64_t I, const Pack<V>& xc, float xm1, float xp4, const Pack<V>& ym, const Pack<V>& yp,
                                             const Pack<V>& zm, const Pack<V>& zp, bool uni) {
  const GridD& g = l.g;
  const Pack<V> Dc = uni ? splat<V>(l.uD) : ldp<V>(l.Dg + I);
  Pack<V> s;
#pragma unroll
  for (int q = 0; q < V; ++q) s.v[q] = xc.v[q] * Dc.v[q];
  {
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(l.uL); Lp = Lc; }
    else { Lc = ldp<V>(l.L + I); Lp = ld_xp<V>(l.L + I, Lc); }
    Pack<V> xm, xp;
    xm.v[0] = xm1;
#pragma unroll
    for (int q = 1; q < V; ++q) xm.v[q] = xc.v[q - 1];
#pragma unroll
    for (int q = 0; q < V - 1; ++q) xp.v[q] = xc.v[q + 1];
    xp.v[V - 1] = xp4;
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
  {
    const float* Li = l.L + g.comp;
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(l.uL); Lp = Lc; }
    else { Lc = ldp<V>(Li + I); Lp = ldp<V>(Li + I + g.s[1]); }
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (ym.v[q] * Lc.v[q] + yp.v[q] * Lp.v[q]);
  }
  {
    const float* Li = l.L + 2 * g.comp;
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(l.uL); Lp = Lc; }
    else { Lc = ldp<V>(Li + I); Lp = ldp<V>(Li + I + g.s[2]); }
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (zm.v[q] * Lc.v[q] + zp.v[q] * Lp.v[q]);
  }
  return s;
}
// the same with the x/y neighbours read from memory (wrapped indices) and the z-neighbours from registers
template <int V>
__device__ __forceinline__ Pack<V> mult_zreg(const Lvl& l, const float* x, int i, int j, int64_t I, const Pack<V>& xc, const Pack<V>& zm,
                                             const Pack<V>& zp, bool uni) {
  const GridD& g = l.g;
  const float xm1 = x[(g.per[0] && i == 1) ? I + (g.n[0] - 3) : I - 1];
  const float xp4 = x[(g.per[0] && i + V - 1 == g.n[0] - 2) ? I + (V - 1) - (g.n[0] - 3) : I + V];
  const int64_t st = g.s[1];
  const int n = g.n[1];
  const int64_t om = (g.per[1] && j == 1) ? (int64_t)(n - 3) * st : -st;
  const int64_t op = (g.per[1] && j == n - 2) ? -(int64_t)(n - 3) * st : st;
  const Pack<V> ym = ldp<V>(x + I + om), yp = ldp<V>(x + I + op);
  return mult_vals<V>(l, I, xc, xm1, xp4, ym, yp, zm, zp, uni);
}
// offsets of the z-neighbours of plane k (periodic wrap of a level that is periodic in z on this rank)
__device__ __forceinline__ int64_t zoff_m(const GridD& g, int k) { return (g.per[2] && k == 1) ? (int64_t)(g.n[2] - 3) * g.s[2] : -g.s[2]; }
__device__ __forceinline__ int64_t zoff_p(const GridD& g, int k) { return (g.per[2] && k == g.n[2] - 2) ? -(int64_t)(g.n[2] - 3) * g.s[2] : g.s[2]; }

// ---- z-marching iteration (3-D, V = 4) ---------------------------------------------------------------
// With one (row-group, plane) unit per block (for_rows) every 7-point stencil pulls its four y/z neighbour
// rows through L2 again: ~5 words per cell of L2 traffic for a kernel whose HBM traffic is 2 words (measured:
// k_pcg_mult stalled at 5.6 TB/s of L2 reads).  Here a block owns an (mx packs × my rows) tile and marches
// along z: a thread keeps the packs of planes k-1, k, k+1 of its column in registers (one new load per
// step), its y-neighbours were loaded one step earlier by the adjacent threads of the same block (L1 hits),
// and only the tile's perimeter rows come from L2.
struct MarchGeom { int mx, my, ntx, nty, ntz, zchunk; };
template <class F>
__device__ __forceinline__ void for_columns(const GridD& g, const MarchGeom& mg, F&& f) {
  const int ntiles = mg.ntx * mg.nty * mg.ntz;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tx = t % mg.ntx, q = t / mg.ntx;
    const int ty = q % mg.nty, tz = q / mg.nty;
    const int i = 1 + 4 * (tx * mg.mx + (int)threadIdx.x), j = 1 + ty * mg.my + (int)threadIdx.y;
    const int k0 = 1 + tz * mg.zchunk, k1 = min(k0 + mg.zchunk - 1, g.n[2] - 2);
    if (i <= g.n[0] - 2 && j <= g.n[1] - 2) f(i, j, k0, k1);
  }
}

// ---- z-marching with the stencil operand staged in shared memory --------------------------------------
// A kernel whose only HBM input is the stencil operand (pcg!'s z = Aϵ reads ϵ and nothing else) has one
// 16-byte load per thread and step in flight — at 32 warps per SM that is a quarter of what HBM latency ×
// bandwidth needs, and the kernel ran at 2.7 TB/s.  Here the planes of the operand travel through a ring of
// WL_NST shared-memory stages filled by cp.async (LDGSTS: no registers held while the miss is outstanding),
// two planes ahead of the one being differenced; every neighbour read is then an LDS with a compile-time
// shape: [my + 2 rows][4 pad | 4·mx | 4 pad floats], the x-halo cells in the inner float of each pad.
// Each thread copies its own pack; edge threads add the tile's halo rows / cells (periodic wrap in the address).
#define WL_NST 5
__device__ __forceinline__ unsigned int wl_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wl_smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wl_smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct StageCtx {
  float* ring;       // WL_NST planes
  int row, plane;    // floats per staged row / plane
  int mine;          // offset of this thread's pack inside a staged plane
  int64_t col;       // i + j*s1 of the thread's pack
  int64_t colm, colp;  // the same for the y-neighbour rows (wrapped)
  int64_t xm, xp;    // i-1 / i+4 (wrapped) + j*s1
  bool ylo, yhi, xlo, xhi;  // this thread also copies the tile's lower/upper halo row, left/right halo cell
};
__device__ __forceinline__ StageCtx stage_ctx(const GridD& g, float* ring, int i, int j) {
  StageCtx c;
  c.ring = ring;
  c.row = 4 * blockDim.x + 8;
  c.plane = (blockDim.y + 2) * c.row;
  c.mine = (threadIdx.y + 1) * c.row + 4 + 4 * threadIdx.x;
  const int64_t s1 = g.s[1];
  c.col = i + j * s1;
  const int n1 = g.n[1], n0 = g.n[0];
  c.colm = i + (int64_t)((g.per[1] && j == 1) ? n1 - 2 : j - 1) * s1;
  c.colp = i + (int64_t)((g.per[1] && j == n1 - 2) ? 1 : j + 1) * s1;
  c.xm = ((g.per[0] && i == 1) ? n0 - 2 : i - 1) + j * s1;
  c.xp = ((g.per[0] && i + 3 == n0 - 2) ? 1 : i + 4) + j * s1;
  c.ylo = threadIdx.y == 0; c.yhi = threadIdx.y == blockDim.y - 1;
  c.xlo = threadIdx.x == 0; c.xhi = threadIdx.x == blockDim.x - 1;
  return c;
}
// requests plane P (0 … n2-1; wrapped if the level is periodic in z) of field S into ring stage `st`
__device__ __forceinline__ void stage_issue(const StageCtx& c, const GridD& g, const float* __restrict__ S, int P, int st) {
  if (g.per[2]) P = P == 0 ? g.n[2] - 2 : (P == g.n[2] - 1 ? 1 : P);
  const float* src = S + (int64_t)P * g.s[2];
  float* dst = c.ring + st * c.plane;
  cp_async16(dst + c.mine, src + c.col);
  if (c.ylo) cp_async16(dst + c.mine - c.row, src + c.colm);
  if (c.yhi) cp_async16(dst + c.mine + c.row, src + c.colp);
  if (c.xlo) cp_async4(dst + c.mine - 1, src + c.xm);
  if (c.xhi) cp_async4(dst + c.mine + 4, src + c.xp);
}
__device__ __forceinline__ Pack<4> lds4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  Pack<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
// Marches a column block over planes k0 … k1 with field S staged: body(k, I, xc, Ax-operands…) is called with
// the seven neighbour values of the thread's pack.  `pre(k, I)` runs before the step's wait + barrier (issue the
// step's own global loads there: they overlap the wait).  All threads of the block must call it (barriers).
template <class Pre, class Body>
__device__ __forceinline__ void march_staged(const GridD& g, const float* __restrict__ S, float* ring, int i, int j, int k0, int k1, Pre&& pre, Body&& body) {
  const StageCtx c = stage_ctx(g, ring, i, j);
  // planes k0-1, k0, k0+1, k0+2 → stages 0, 1, 2, 3 (one group each)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (k0 - 1 + q <= k1 + 1) stage_issue(c, g, S, k0 - 1 + q, q);
    cp_async_commit_group();
  }
  int st = 1;  // stage of plane k
  int64_t I = c.col + (int64_t)k0 * g.s[2];
  for (int k = k0; k <= k1; ++k, I += g.s[2]) {
    pre(k, I);
    cp_async_wait_group<1>();   // planes ≤ k+1 have landed (the newest group, k+2, may still be in flight)
    __syncthreads();             // … for every thread of the block; and everyone is done with plane k-2
    {                            // plane k+3 takes the stage plane k-2 occupied
      const int P = k + 3;
      if (P <= k1 + 1) stage_issue(c, g, S, P, (st + 3) % WL_NST);
      cp_async_commit_group();
    }
    const float* pc = c.ring + st * c.plane + c.mine;
    const float* pm = c.ring + ((st + WL_NST - 1) % WL_NST) * c.plane + c.mine;
    const float* pp = c.ring + ((st + 1) % WL_NST) * c.plane + c.mine;
    body(k, I, lds4(pc), pc[-1], pc[4], lds4(pc - c.row), lds4(pc + c.row), lds4(pm), lds4(pp));
    st = (st + 1) % WL_NST;
  }
  cp_async_wait_group<0>();
  __syncthreads();  // the ring is reused by the block's next tile
}

// ---- prolongate! + increment! + the head of pcg! in one pass -------------------------------------
// Vcycle! ends with  ϵ_f = x_c[down(I)] (prolongate!, MultiLevelPoisson.jl:34) ; perBC!(ϵ) ; r -= Aϵ ;
// x += ϵ (increment!, Poisson.jl:99-103), and the pcg! that follows opens with ϵ = r·iD ; ρ = r⋅ϵ
// (:125-126).  ϵ_f is a pure function of the coarse x, so it is evaluated on the fly at the cell and
// its 2D neighbours (periodic wrap / slab halo handled in the index map) instead of being stored and
// re-read: r, x and the NEW ϵ are written in place with no read-after-write hazard.  13.1 → 10.1 words.
struct ProMap {
  int per[3];       // wrap the fine index in this dimension (periodic and not z-distributed)
  int n[3];         // fine padded sizes of this rank
  int zmode;        // 1: the coarse level is replicated while the fine one is z-distributed
  int zoff;         // global fine padded z index of local plane 0            (zmode 1)
  int nzg;          // interior z-planes of the whole fine level             (zmode 1)
  int perz_global;  // the global domain is periodic in z                     (zmode 1)
};
__device__ __forceinline__ int pro_cidx(const ProMap& m, int d, int v) {
  if (d == 2 && m.zmode == 1) {
    int kg = v + m.zoff;
    if (m.perz_global) kg = kg == 0 ? m.nzg : (kg == m.nzg + 1 ? 1 : kg);
    return (kg + 1) >> 1;
  }
  if (m.per[d]) v = v == 0 ? m.n[d] - 2 : (v == m.n[d] - 1 ? 1 : v);
  return (v + 1) >> 1;
}
template <int D, int V>
__device__ __forceinline__ void prolong_increment_init(const Lvl& f, const Lvl& c, const ProMap& m, int i, int j, int k, int64_t I,
                                                       double& acc_rho, double& acc_r2) {
  const float* __restrict__ xc = c.x;
  const int64_t cs1 = c.g.s[1], cs2 = c.g.s[2];
  int ci[V];
#pragma unroll
  for (int q = 0; q < V; ++q) ci[q] = (i + q + 1) >> 1;  // interior cells never wrap
  const int64_t J = (D > 1 ? pro_cidx(m, 1, j) * cs1 : 0) + (D > 2 ? pro_cidx(m, 2, k) * cs2 : 0);
  Pack<V> e, xm, xp;
#pragma unroll
  for (int q = 0; q < V; ++q) e.v[q] = xc[J + ci[q]];
  xm.v[0] = xc[J + pro_cidx(m, 0, i - 1)];
#pragma unroll
  for (int q = 1; q < V; ++q) xm.v[q] = e.v[q - 1];
#pragma unroll
  for (int q = 0; q < V - 1; ++q) xp.v[q] = e.v[q + 1];
  xp.v[V - 1] = xc[J + pro_cidx(m, 0, i + V)];
  const bool uni = pack_uniform<V>(f, I);
  const Pack<V> Dc = uni ? splat<V>(f.uD) : ldp<V>(f.Dg + I);
  Pack<V> s;
#pragma unroll
  for (int q = 0; q < V; ++q) s.v[q] = e.v[q] * Dc.v[q];
  {
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(f.uL); Lp = Lc; }
    else { Lc = ldp<V>(f.L + I); Lp = ld_xp<V>(f.L + I, Lc); }
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
#pragma unroll
  for (int d = 1; d < D; ++d) {
    const int cd = d == 1 ? j : k;
    const int64_t cst = d == 1 ? cs1 : cs2;
    const int64_t Jo = J - pro_cidx(m, d, cd) * cst;  // J without dimension d's term
    const int64_t Jm = Jo + pro_cidx(m, d, cd - 1) * cst, Jp = Jo + pro_cidx(m, d, cd + 1) * cst;
    const float* Li = f.L + d * f.g.comp;
    Pack<V> Lc, Lp;
    if (uni) { Lc = splat<V>(f.uL); Lp = Lc; }
    else { Lc = ldp<V>(Li + I); Lp = ldp<V>(Li + I + f.g.s[d]); }
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xc[Jm + ci[q]] * Lc.v[q] + xc[Jp + ci[q]] * Lp.v[q]);
  }
  Pack<V> r = ldp<V>(f.r + I), x = ldp<V>(f.x + I);
  const Pack<V> id = ld_iD<V>(f, I, uni);
  Pack<V> en;
#pragma unroll
  for (int q = 0; q < V; ++q) {
    r.v[q] = r.v[q] - s.v[q];
    x.v[q] = x.v[q] + e.v[q];
    en.v[q] = r.v[q] * id.v[q];
    acc_rho += (double)r.v[q] * (double)en.v[q];
    acc_r2 += (double)r.v[q] * (double)r.v[q];
  }
  stp<V>(f.r + I, r);
  stp<V>(f.x + I, x);
  stp<V>(f.eps + I, en);
}

// interior rows; f(i, j, k, I) is called for the first cell of each V-group
template <int V, class F>
__device__ __forceinline__ void for_rows(const GridD& g, F&& f) {
  const int nj = g.D > 1 ? g.n[1] - 2 : 1, nk = g.D > 2 ? g.n[2] - 2 : 1;
  const int njt = (nj + blockDim.y - 1) / blockDim.y;
  const int nunits = njt * nk;
  const int nxi = g.n[0] - 2;
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int kk = unit / njt, jt = unit - kk * njt;
    const int j = 1 + jt * blockDim.y + threadIdx.y;
    if (j > nj) continue;
    const int k = g.D > 2 ? 1 + kk : 0;
    const int64_t row = j * g.s[1] + k * g.s[2];
    for (int i = 1 + V * threadIdx.x; i <= nxi; i += V * blockDim.x) f(i, j, k, row + i);
  }
}
#endif

// host-side launch geometry for for_rows kernels
struct RowLaunch { int V; dim3 blk; int nunits; };
inline RowLaunch row_launch(const GridD& g) {
  RowLaunch r;
  const int nxi = g.n[0] - 2;
  r.V = (nxi % 4 == 0) ? 4 : 1;
  const int per = nxi / r.V;
  int bx = 8;
  while (bx < per && bx < 128) bx <<= 1;
  r.blk = dim3(bx, 256 / bx, 1);
  const int nj = g.D > 1 ? g.n[1] - 2 : 1, nk = g.D > 2 ? g.n[2] - 2 : 1;
  r.nunits = ((nj + (int)r.blk.y - 1) / (int)r.blk.y) * nk;
  return r;
}

// host-side geometry of the z-marching kernels; ok = false when the level should use for_rows
struct MarchLaunch { bool ok; bool staged; MarchGeom mg; dim3 blk; int grid; size_t smem; };
inline MarchLaunch march_launch(const GridD& g, int nsm) {
  MarchLaunch m{};
  const int nxi = g.n[0] - 2, nyi = g.n[1] - 2, nzi = g.D > 2 ? g.n[2] - 2 : 0;
  m.ok = g.D == 3 && nxi % 4 == 0 && nxi >= 32 && nyi >= 8 && nzi >= 8;
  if (!m.ok) return m;
  const int packs = nxi / 4;
  int mx = 8;
  while (mx < packs && mx < 32) mx <<= 1;
  const int my = 256 / mx;
  m.mg.mx = mx; m.mg.my = my;
  m.mg.ntx = (packs + mx - 1) / mx;
  m.mg.nty = (nyi + my - 1) / my;
  int zc = 64;
  while (zc > 8 && (int64_t)m.mg.ntx * m.mg.nty * ((nzi + zc - 1) / zc) < (int64_t)nsm * 8) zc >>= 1;  // enough blocks to fill the machine
  m.mg.zchunk = zc;
  m.mg.ntz = (nzi + zc - 1) / zc;
  m.blk = dim3(mx, my, 1);
  const int64_t ntiles = (int64_t)m.mg.ntx * m.mg.nty * m.mg.ntz;
  m.grid = (int)(ntiles < WL_MAXBLOCKS ? ntiles : WL_MAXBLOCKS);
  // the shared-memory staged kernels have block-wide barriers in the march: every thread of every tile must own a pack
  m.staged = packs % mx == 0 && nyi % my == 0;
  m.smem = sizeof(float) * (size_t)WL_NST * (my + 2) * (4 * mx + 8);
  return m;
}

// launch a for_rows kernel<D,V> on the interior of grid `grd`
#define RLAUNCH(s, KERN, grd, ...)                                                               \
  do {                                                                                            \
    const RowLaunch rl_ = row_launch(grd);                                                        \
    const int gs_ = wl_grid_for((s), rl_.nunits);                                                 \
    if ((s).D == 2) {                                                                             \
      if (rl_.V == 4) KERN<2, 4><<<gs_, rl_.blk, 0, (s).stream>>>(__VA_ARGS__);                   \
      else KERN<2, 1><<<gs_, rl_.blk, 0, (s).stream>>>(__VA_ARGS__);                              \
    } else {                                                                                      \
      if (rl_.V == 4) KERN<3, 4><<<gs_, rl_.blk, 0, (s).stream>>>(__VA_ARGS__);                   \
      else KERN<3, 1><<<gs_, rl_.blk, 0, (s).stream>>>(__VA_ARGS__);                              \
    }                                                                                             \
    ++(s).launches;                                                                               \
  } while (0)

