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
// The staged marches read the mask byte of their pack once per plane.  The byte is tiny but its load is a full HBM
// round trip: consumed in the step that issued it (or the next: a step of a fully fed SM is ~0.4 µs, the loaded HBM
// latency ~1 µs) it was the largest stall of the staged kernels (32 % of the warp samples of k_pcg_mult_s waited on the
// compare).  MaskPipe keeps the raw byte of the next plane in a register, prefetches the byte WL_MASK_PF planes ahead
// into L1, and compares only where the flag is used.
#ifndef WL_MASK_PF
#define WL_MASK_PF 6
#endif
// planes ahead at which the per-step operand loads of the marching tiles (r, x, div u, coarse x) are prefetched into L1
#ifndef WL_PF
#define WL_PF 3
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
struct MaskPipe {
  const unsigned char* p;  // byte of the plane after the one `nxt` belongs to
  int64_t stp;             // bytes per plane (s2 / 4: rows and planes are multiples of four cells)
  unsigned int cur, nxt;
  int pf;                  // steps that still have a plane WL_MASK_PF ahead inside the array
  // before the march: I0 = first cell of the pack on plane k0
  __device__ __forceinline__ void init(const Lvl& l, int64_t I0, int k0) {
    p = l.umask; cur = 1u; nxt = 1u; stp = l.g.s[2] >> 2; pf = l.g.n[2] - k0 - WL_MASK_PF;
    if (p != nullptr) { p += (I0 + l.g.lead) >> 2; nxt = *p; p += stp; }
  }
  // top of every step, first at plane k0; planes ≤ n2 - 2, so the byte of the next plane exists
  __device__ __forceinline__ void step() {
    cur = nxt;
    if (p == nullptr) return;
    nxt = *p;
    if (pf > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (WL_MASK_PF - 1) * stp));
    --pf;
    p += stp;
  }
  __device__ __forceinline__ bool uni() const {
    unsigned int m = cur;
    asm volatile("" : "+r"(m));  // the compare stays at the point of use
    return m == 0u;
  }
};
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
  Pack<V> xm, xp, s;
  xm.v[0] = xm1;
#pragma unroll
  for (int q = 1; q < V; ++q) xm.v[q] = xc.v[q - 1];
#pragma unroll
  for (int q = 0; q < V - 1; ++q) xp.v[q] = xc.v[q + 1];
  xp.v[V - 1] = xp4;
  if (uni) {  // a branch of its own: the constants are used as scalar operands (a select per coefficient cost 24 moves per pack)
    // uL is a power of two (1 on the finest level, ×2 per restrictL in 3-D): a·L is exact, a·L + b·L = (a + b)·L with the
    // single rounding of the sum, and t + (…)·L rounds once — (a + b) then fmaf(·, L, t) gives the bits of the
    // generic expression x[I-δ]·L[I] + x[I+δ]·L[I+δ] in 7 instead of 13 operations per cell
    const float L = l.uL, Dd = l.uD;
#pragma unroll
    for (int q = 0; q < V; ++q) {
      float t = xc.v[q] * Dd;
      t = fmaf(xm.v[q] + xp.v[q], L, t);
      t = fmaf(ym.v[q] + yp.v[q], L, t);
      t = fmaf(zm.v[q] + zp.v[q], L, t);
      s.v[q] = t;
    }
    return s;
  }
  const Pack<V> Dc = ldp<V>(l.Dg + I);
#pragma unroll
  for (int q = 0; q < V; ++q) s.v[q] = xc.v[q] * Dc.v[q];
  {
    const Pack<V> Lc = ldp<V>(l.L + I), Lp = ld_xp<V>(l.L + I, Lc);
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (xm.v[q] * Lc.v[q] + xp.v[q] * Lp.v[q]);
  }
  {
    const float* Li = l.L + g.comp;
    const Pack<V> Lc = ldp<V>(Li + I), Lp = ldp<V>(Li + I + g.s[1]);
#pragma unroll
    for (int q = 0; q < V; ++q) s.v[q] += (ym.v[q] * Lc.v[q] + yp.v[q] * Lp.v[q]);
  }
  {
    const float* Li = l.L + 2 * g.comp;
    const Pack<V> Lc = ldp<V>(Li + I), Lp = ldp<V>(Li + I + g.s[2]);
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
struct MarchGeom { int mx, my, ntx, nty, ntz, zchunk; };  // mx (a power of two) · my = threads of the block
// position of the calling thread inside its mx × my tile (blocks may be 1-D: the persistent solver kernel)
struct TilePos { int tx, ty; };
__device__ __forceinline__ TilePos tile_pos(const MarchGeom& mg) {
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  TilePos t; t.tx = tid & (mg.mx - 1); t.ty = tid / mg.mx;
  return t;
}
template <class F>
__device__ __forceinline__ void for_columns(const GridD& g, const MarchGeom& mg, F&& f) {
  const int ntiles = mg.ntx * mg.nty * mg.ntz;
  const TilePos tp = tile_pos(mg);
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tx = t % mg.ntx, q = t / mg.ntx;
    const int ty = q % mg.nty, tz = q / mg.nty;
    const int i = 1 + 4 * (tx * mg.mx + tp.tx), j = 1 + ty * mg.my + tp.ty;
    const int k0 = 1 + tz * mg.zchunk, k1 = min(k0 + mg.zchunk - 1, g.n[2] - 2);
    if (i <= g.n[0] - 2 && j <= g.n[1] - 2) f(i, j, k0, k1);
  }
}

// The persistent solver kernel has a fixed number of blocks: the (tile, plane) pairs are dealt out in equal
// contiguous runs, planes fastest, so a block marches through ~total/gridDim planes of (at most two) tiles with one
// pipeline fill per run instead of one per short z-chunk.
template <class F>
__device__ __forceinline__ void for_columns_even(const GridD& g, const MarchGeom& mg, F&& f) {
  const int nz = g.n[2] - 2;
  const int64_t total = (int64_t)mg.ntx * mg.nty * nz;
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  int64_t lo = per * blockIdx.x;
  const int64_t hi = min(total, lo + per);
  const TilePos tp = tile_pos(mg);
  while (lo < hi) {
    const int t = (int)(lo / nz), kz = (int)(lo - (int64_t)t * nz);
    const int len = (int)min(hi - lo, (int64_t)(nz - kz));
    const int tx = t % mg.ntx, ty = t / mg.ntx;
    const int i = 1 + 4 * (tx * mg.mx + tp.tx), j = 1 + ty * mg.my + tp.ty;
    if (i <= g.n[0] - 2 && j <= g.n[1] - 2) f(i, j, 1 + kz, kz + len);
    lo += len;
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
#ifndef WL_AHEAD
#define WL_AHEAD 2          // planes in flight beyond the three being read
#endif
#define WL_NST (3 + WL_AHEAD)
__device__ __forceinline__ unsigned int wl_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wl_smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem, const float* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wl_smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Everything the inner loop needs is kept as running pointers / shared-memory byte addresses that advance by a
// constant per step: the first version recomputed 64-bit addresses and ring slots (mod 5) from indices every step
// and spent 270 instructions per step on ~60 of arithmetic — it was issue-bound, not memory-bound.
struct StageCtx {
  unsigned int ring0;   // shared-memory byte address of this thread's pack in stage 0
  unsigned int pbytes;  // bytes per staged plane
  int rowb;             // bytes per staged row
  int64_t col;          // i + j*s1 of the thread's pack
  int colm, colp;       // offsets from col to the y-neighbour rows (wrapped)
  int xm, xp;           // … to cell i-1 / i+4 of the row (wrapped)
  bool ylo, yhi, xlo, xhi;  // this thread also copies the tile's lower/upper halo row, left/right halo cell
};
__device__ __forceinline__ StageCtx stage_ctx(const GridD& g, const MarchGeom& mg, float* ring, int i, int j) {
  StageCtx c;
  const TilePos tp = tile_pos(mg);
  const int row = 4 * mg.mx + 8;
  c.rowb = row * (int)sizeof(float);
  c.pbytes = (unsigned int)((mg.my + 2) * row * sizeof(float));
  c.ring0 = wl_smem_u32(ring) + (unsigned int)(((tp.ty + 1) * row + 4 + 4 * tp.tx) * sizeof(float));
  const int64_t s1 = g.s[1];
  c.col = i + j * s1;
  const int n1 = g.n[1], n0 = g.n[0];
  c.colm = (int)((((g.per[1] && j == 1) ? n1 - 2 : j - 1) - j) * s1);
  c.colp = (int)((((g.per[1] && j == n1 - 2) ? 1 : j + 1) - j) * s1);
  c.xm = ((g.per[0] && i == 1) ? n0 - 2 : i - 1) - i;
  c.xp = ((g.per[0] && i + 3 == n0 - 2) ? 1 : i + 4) - i;
  c.ylo = tp.ty == 0; c.yhi = tp.ty == mg.my - 1;
  c.xlo = tp.tx == 0; c.xhi = tp.tx == mg.mx - 1;
  return c;
}
__device__ __forceinline__ void cp_async16s(unsigned int smem, const float* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4s(unsigned int smem, const float* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(gmem) : "memory");
}
// requests the plane at `src` (pointer to its cell (0,0)) into the stage whose pack address is `dst`
__device__ __forceinline__ void stage_issue(const StageCtx& c, const float* __restrict__ src, unsigned int dst) {
  const float* p = src + c.col;
  cp_async16s(dst, p);
  if (c.ylo) cp_async16s(dst - c.rowb, p + c.colm);
  if (c.yhi) cp_async16s(dst + c.rowb, p + c.colp);
  if (c.xlo) cp_async4s(dst - 4, p + c.xm);
  if (c.xhi) cp_async4s(dst + 16, p + c.xp);
}
// plane P (0 … n2-1) of a field, wrapped if the level is periodic in z
__device__ __forceinline__ const float* plane_ptr(const GridD& g, const float* __restrict__ S, int P) {
  if (g.per[2]) P = P == 0 ? g.n[2] - 2 : (P == g.n[2] - 1 ? 1 : P);
  return S + (int64_t)P * g.s[2];
}
__device__ __forceinline__ Pack<4> lds4(unsigned int a) {
  Pack<4> r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(a));
  return r;
}
__device__ __forceinline__ float lds1(unsigned int a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
// Marches a column block over planes k0 … k1 with field S staged: body(k, I, xc, xm1, xp4, ym, yp, zm, zp) is called with
// the neighbour values of the thread's pack.  `pre(k, I)` runs before the step's wait + barrier (issue the step's own
// global loads there: they overlap the wait).  All threads of the block must call it (barriers).
template <class Pre, class Body>
__device__ __forceinline__ void march_staged(const GridD& g, const MarchGeom& mg, const float* __restrict__ S, float* ring, int i, int j, int k0, int k1,
                                             Pre&& pre, Body&& body) {
  const StageCtx c = stage_ctx(g, mg, ring, i, j);
  const unsigned int ring_end = c.ring0 + WL_NST * c.pbytes;
  // planes k0-1 … k0+WL_AHEAD → stages 0 … 1+WL_AHEAD (one group each)
#pragma unroll
  for (int q = 0; q < 2 + WL_AHEAD; ++q) {
    if (k0 - 1 + q <= k1 + 1) stage_issue(c, plane_ptr(g, S, k0 - 1 + q), c.ring0 + q * c.pbytes);
    cp_async_commit_group();
  }
  // stages of planes k-1, k, k+1 and of the plane requested at step k (k+1+WL_AHEAD: it takes the stage of plane k-2)
  unsigned int am = c.ring0, ac = c.ring0 + c.pbytes, ap = c.ring0 + 2 * c.pbytes, ai = c.ring0 + (WL_NST - 1) * c.pbytes;
  const float* nxt = S + (int64_t)(k0 + 1 + WL_AHEAD) * g.s[2];  // (a periodic wrap can only concern the last plane of the level: handled below)
  const int klast = k1 - WL_AHEAD;                                // last step that still requests a plane (k+1+WL_AHEAD ≤ k1+1)
  const int kwrap = g.per[2] ? g.n[2] - 2 - WL_AHEAD : -1;        // the step whose request is the upper ghost plane = image of plane 1
  const float* const pwrap = S + g.s[2];
  int64_t I = c.col + (int64_t)k0 * g.s[2];
#pragma unroll 1
  for (int k = k0; k <= k1; ++k) {
    pre(k, I);
    cp_async_wait_group<WL_AHEAD - 1>();   // planes ≤ k+1 have landed (the newer groups may still be in flight)
    __syncthreads();                        // … for every thread of the block; and everyone is done with plane k-2
    if (k <= klast) stage_issue(c, k == kwrap ? pwrap : nxt, ai);
    cp_async_commit_group();
    body(k, I, lds4(ac), lds1(ac - 4), lds1(ac + 16), lds4(ac - c.rowb), lds4(ac + c.rowb), lds4(am), lds4(ap));
    am = ac; ac = ap;
    ap += c.pbytes; if (ap == ring_end) ap = c.ring0;
    ai += c.pbytes; if (ai == ring_end) ai = c.ring0;
    nxt += g.s[2];
    I += g.s[2];
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
__device__ __forceinline__ void prolong_increment_init(const Lvl& f, const Lvl& c, const ProMap& m, const float* __restrict__ xsrc, int i, int j, int k,
                                                       int64_t I, double& acc_rho, double& acc_r2) {
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
  Pack<V> r = ldp<V>(f.r + I), x = ldp<V>(xsrc + I);  // xsrc: f.x, or f.eps when the fused head of solver! parked x there (wl_resjac.cu)
  const Pack<V> id = ld_iD<V>(f, I, uni);
  Pack<V> en;
#pragma unroll
  for (int q = 0; q < V; ++q) {
    r.v[q] = r.v[q] - s.v[q];
    x.v[q] = x.v[q] + e.v[q];
    en.v[q] = r.v[q] * id.v[q];
    acc_rho = dfma(r.v[q], en.v[q], acc_rho);
    acc_r2 = dfma(r.v[q], r.v[q], acc_r2);
  }
  stp<V>(f.r + I, r);
  stp<V>(f.x + I, x);
  stp<V>(f.eps + I, en);
}

// ---- the stencil phases on one column tile (3-D, V = 4) — shared by the per-level kernels (wl_poisson.cu) and
// ---- the persistent solver kernel (wl_vcycle.cu).  Arithmetic per cell is that of the for_rows kernels.
// pcg!: z = Aϵ, acc += z·ϵ (Poisson.jl:130-131)
__device__ __forceinline__ void tile_pcg_mult(const Lvl& l, const MarchGeom& mg, float* ring, int i, int j, int k0, int k1, double& acc) {
  // the pack-mask byte of the NEXT plane is requested one step early: a load issued and consumed within the same step
  // exposed a full HBM latency per step (the largest stall of the first profile)
  MaskPipe mp;
  mp.init(l, i + j * l.g.s[1] + (int64_t)k0 * l.g.s[2], k0);
  march_staged(l.g, mg, l.eps, ring, i, j, k0, k1,
    [&](int, int64_t) { mp.step(); },
    [&](int, int64_t I, const Pack<4>& ec, float xm1, float xp4, const Pack<4>& ym, const Pack<4>& yp, const Pack<4>& zm, const Pack<4>& zp) {
      const Pack<4> z = mult_vals<4>(l, I, ec, xm1, xp4, ym, yp, zm, zp, mp.uni());
#pragma unroll
      for (int q = 0; q < 4; ++q) acc = dfma(z.v[q], ec.v[q], acc);
      stp<4>(l.z + I, z);
    });
}
// increment!: r -= Aϵ, x += ϵ, acc += r·r (Poisson.jl:101-102, :146); r and x of the next plane travel one step ahead too
__device__ __forceinline__ void tile_increment(const Lvl& l, const MarchGeom& mg, float* ring, int i, int j, int k0, int k1, double& acc) {
  MaskPipe mp;
  mp.init(l, i + j * l.g.s[1] + (int64_t)k0 * l.g.s[2], k0);
  Pack<4> r, x, rn, xn;
  march_staged(l.g, mg, l.eps, ring, i, j, k0, k1,
    [&](int k, int64_t I) {
      mp.step();
      if (k == k0) { r = ldp<4>(l.r + I); x = ldp<4>(l.x + I); }
      else { r = rn; x = xn; }
      const int64_t In = I + l.g.s[2];  // plane k1+1 is a ghost / halo plane of the same arrays: readable, never used
      rn = ldp<4>(l.r + In); xn = ldp<4>(l.x + In);
      if (WL_PF > 1 && k + WL_PF <= l.g.n[2] - 1) { prefetch_l1(l.r + I + WL_PF * l.g.s[2]); prefetch_l1(l.x + I + WL_PF * l.g.s[2]); }
    },
    [&](int, int64_t I, const Pack<4>& ec, float xm1, float xp4, const Pack<4>& ym, const Pack<4>& yp, const Pack<4>& zm, const Pack<4>& zp) {
      const Pack<4> ae = mult_vals<4>(l, I, ec, xm1, xp4, ym, yp, zm, zp, mp.uni());
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        r.v[q] = r.v[q] - ae.v[q];
        x.v[q] = x.v[q] + ec.v[q];
        acc = dfma(r.v[q], r.v[q], acc);
      }
      stp<4>(l.r + I, r);
      stp<4>(l.x + I, x);
    });
}
// residual!: r = iD == 0 ? 0 : z - Ax (z read, or evaluated as div(divu)), a0 += r, a1 += r·r; optionally ϵ = r·iD
__device__ __forceinline__ void tile_residual(const Lvl& l, const MarchGeom& mg, float* ring, int i, int j, int k0, int k1, int write_eps,
                                              const float* __restrict__ divu, double& a0, double& a1) {
  MaskPipe mp;
  Pack<4> zc;
  const TilePos tpos = tile_pos(mg);
  const bool xhi = tpos.tx == mg.mx - 1, yhi = tpos.ty == mg.my - 1;
  mp.init(l, i + j * l.g.s[1] + (int64_t)k0 * l.g.s[2], k0);
  march_staged(l.g, mg, l.x, ring, i, j, k0, k1,
    [&](int k, int64_t I) {
      mp.step();
      zc = divu ? div_pack<3, 4>(l.g, divu, I) : ldp<4>(l.z + I);
      // the operands of a later plane's z: own pack of each component (the neighbour threads fetch theirs), the cell
      // right of / row above the tile by its edge threads; u₃ of plane k+1 is the pack fetched one step later
      if (WL_PF > 0 && k + WL_PF <= l.g.n[2] - 1) {
        if (divu) {
          const float* q = divu + I + WL_PF * l.g.s[2];
          prefetch_l1(q); prefetch_l1(q + l.g.comp); prefetch_l1(q + 2 * l.g.comp);
          if (xhi) prefetch_l1(q + 4);
          if (yhi) prefetch_l1(q + l.g.comp + l.g.s[1]);
        } else prefetch_l1(l.z + I + WL_PF * l.g.s[2]);
      }
    },
    [&](int, int64_t I, const Pack<4>& xc, float xm1, float xp4, const Pack<4>& ym, const Pack<4>& yp, const Pack<4>& zm, const Pack<4>& zp) {
      const bool uni = mp.uni();
      const Pack<4> ax = mult_vals<4>(l, I, xc, xm1, xp4, ym, yp, zm, zp, uni), id = ld_iD<4>(l, I, uni);
      Pack<4> r;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        r.v[q] = id.v[q] == 0.f ? 0.f : zc.v[q] - ax.v[q];
        a0 += (double)r.v[q];
        a1 = dfma(r.v[q], r.v[q], a1);
      }
      stp<4>(l.r + I, r);
      if (write_eps) {
        Pack<4> e;
#pragma unroll
        for (int q = 0; q < 4; ++q) e.v[q] = r.v[q] * id.v[q];
        stp<4>(l.eps + I, e);
      }
    });
}
// prolongate! + increment! + head of pcg! (see prolong_increment_init): ϵ of a fine pack (i = 1+4t) is {c, c, c', c'}
// with (c, c') the aligned coarse pair at ci = 1+2t, so a pack costs one 8-byte load per coarse row it touches
__device__ __forceinline__ Pack<4> pro_pack(const float* __restrict__ crow, int ci) {
  const float2 c = *reinterpret_cast<const float2*>(crow + ci);
  Pack<4> e; e.v[0] = c.x; e.v[1] = c.x; e.v[2] = c.y; e.v[3] = c.y;
  return e;
}
__device__ __forceinline__ void tile_prolong(const Lvl& f, const Lvl& c, const ProMap& m, const MarchGeom& mg, const float* __restrict__ xsrc, int i, int j,
                                             int k0, int k1, double& a0, double& a1) {
  const GridD& g = f.g;
  const TilePos tpos = tile_pos(mg);
  const bool xlo = tpos.tx == 0, xhi = tpos.tx == mg.mx - 1, ylo = tpos.ty == 0, yhi = tpos.ty == mg.my - 1;
  const float* __restrict__ xc = c.x;
  const int64_t cs1 = c.g.s[1], cs2 = c.g.s[2];
  int64_t I = i + j * g.s[1] + (int64_t)k0 * g.s[2];
  const int ci = (i + 1) >> 1;
  const int cim = pro_cidx(m, 0, i - 1), cip = pro_cidx(m, 0, i + 4);
  const int64_t Jy = pro_cidx(m, 1, j) * cs1, Jym = pro_cidx(m, 1, j - 1) * cs1, Jyp = pro_cidx(m, 1, j + 1) * cs1;
  for (int k = k0; k <= k1; ++k, I += g.s[2]) {
    // Every load of a step is consumed within the step: without help each step exposes a full memory latency (17 warps
    // per issue waited on the scoreboard).  r and x of a later plane, and the coarse-x rows that plane will read (own
    // pair; the tile's edge threads add the rows / cells beyond the tile), are prefetched into L1.
    if (WL_PF > 0 && k + WL_PF <= g.n[2] - 2) {
      prefetch_l1(f.r + I + WL_PF * g.s[2]); prefetch_l1(xsrc + I + WL_PF * g.s[2]);
      const float* q = xc + pro_cidx(m, 2, k + WL_PF + 1) * cs2;
      prefetch_l1(q + Jy + ci);
      if (ylo) prefetch_l1(q + Jym + ci);
      if (yhi) prefetch_l1(q + Jyp + ci);
      if (xlo) prefetch_l1(q + Jy + cim);
      if (xhi) prefetch_l1(q + Jy + cip);
    }
    const int64_t Jz = pro_cidx(m, 2, k) * cs2, Jzm = pro_cidx(m, 2, k - 1) * cs2, Jzp = pro_cidx(m, 2, k + 1) * cs2;
    const float* row = xc + Jy + Jz;
    const Pack<4> e = pro_pack(row, ci);
    const bool uni = pack_uniform<4>(f, I);
    const Pack<4> s = mult_vals<4>(f, I, e, row[cim], row[cip], pro_pack(xc + Jym + Jz, ci), pro_pack(xc + Jyp + Jz, ci), pro_pack(xc + Jy + Jzm, ci),
                                   pro_pack(xc + Jy + Jzp, ci), uni);
    Pack<4> r = ldp<4>(f.r + I), x = ldp<4>(xsrc + I);
    const Pack<4> id = ld_iD<4>(f, I, uni);
    Pack<4> en;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r.v[q] = r.v[q] - s.v[q];
      x.v[q] = x.v[q] + e.v[q];
      en.v[q] = r.v[q] * id.v[q];
      a0 = dfma(r.v[q], en.v[q], a0);
      a1 = dfma(r.v[q], r.v[q], a1);
    }
    stp<4>(f.r + I, r);
    stp<4>(f.x + I, x);
    stp<4>(f.eps + I, en);
  }
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

