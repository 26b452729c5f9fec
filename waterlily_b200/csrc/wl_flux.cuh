// wl_flux.cuh — QUICK flux helpers shared by the conv_diff! kernels (reference src/Flow.jl:1-9,25-34).
#pragma once
#include "wl_common.cuh"

// median (Flow.jl:25-34), branch-free: max(min(a,b), min(max(a,b),c)) is the same value for every
// ordering and tie of finite inputs (only the sign of a zero can differ)
__device__ __forceinline__ float median3(float a, float b, float c) { return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c)); }
// x/6 correctly rounded without the IEEE-division subroutine: one Newton step on q = x·RN(1/6).
// Checked exhaustively against x/6.0f over all 2³² bit patterns: identical for 1e-30 < |x| < 1e30
// and for ±0 (tools/check_div6.c); below 1e-30 it may differ in the last place of a number that small.
__device__ __forceinline__ float div6(float x) {
  const float y = 1.0f / 6.0f;
  const float q = x * y;
  return fmaf(fmaf(-6.0f, q, x), y, q);
}
// quick (Flow.jl:4): median(q1, c, median(q2, c, d)) with q1 = (5c+2d-u)/6, q2 = 10c-9u.  Both medians
// share the endpoint c, so the nest collapses to "c clamped to the range of {q1, q2, d}":
//   d ≥ c:  min(max(q1,c), max(q2,c), d) = max(c, min(q1,q2,d))   (and max(q1,q2,d) ≥ d ≥ c is inactive)
//   d < c:  the mirror image.
// Same value for every finite input (checked on 6·10⁶ random and tie-heavy triples); 2 FMNMX3 + 2 FMNMX
// on sm_100a instead of 8 FMNMX — the ALU pipe these run on is what bounds conv_diff!.
__device__ __forceinline__ float quick(float u, float c, float d) {
  const float q1 = div6(5 * c + 2 * d - u), q2 = 10 * c - 9 * u;
  const float lo = fminf(fminf(q1, q2), d), hi = fmaxf(fmaxf(q1, q2), d);
  return fminf(fmaxf(c, lo), hi);
}

// tile of the two-cells-per-thread core kernel (wl_convdiff_tma.cu): 64 x 8 cells starting at cell (3, 2)
#define WL_CD_TX 64
#define WL_CD_TY 8
#define WL_CD_X0 3
#define WL_CD_Y0 2

struct CDArgs {
  GridD g;
  const float* u;    // field being differenced (u⁰ in the predictor, u in the corrector)
  const float* u0;   // BDIM stage 1 operands
  const float* V;
  float* f;
  float* sigma;      // Φ scratch ≡ flow.σ: only its ghost layer is materialised
  float nu, dt;
  double gacc[3];
  int has_g;
  int fuse_bdim1;
  // BDIM! stage 2 folded into the store (3-D core kernel only): on a tile-plane whose cells — and, for μ₁, whose
  // neighbours — are away from every body (`clean` map, see k_tile_clean) stage 2 is u = (u_acc + f)·scale, so the
  // kernel writes u straight to `ufused` and neither f nor V is touched there.  Null = write f everywhere.
  float* ufused;
  const unsigned char* clean;   // [plane][tile row][tile column], 1 = clean
  float bscale;                 // scale_u! folded into stage 2: 1 (predictor), 0.5 (corrector)
  // boundary-shell cells (k_conv_diff_shell): f is always stored; where the cell's own μ₁ is zero (mask bit 1 clear) stage 2
  // needs no neighbour and is applied on the spot: u_dst = (u_acc + V + μ₀·f)·scale
  const unsigned char* mask;
  const float* mu0;
  const float* uacc;            // null: predictor (u_acc = 0)
};

// one face flux Φ = uf·quick(upwind samples) − ν(u_c − u_{c-1}) (Flow.jl:45): the reference rounds the
// Float64 expression once, which is what the single fmaf does
__device__ __forceinline__ float flux1(float uf, float um2, float um1, float uc, float up1, float nu) {
  const float q = uf > 0 ? quick(um2, um1, uc) : quick(up1, uc, um1);
  return fmaf(uf, q, -(nu * (uc - um1)));
}

// ---- packed Float32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: add/mul/fma.rn.f32x2) -------------------------
// The two-cells-per-thread conv_diff! kernel evaluates every flux for the cell pair (x, x+1) with identical
// operations: issued as packed instructions they take one issue slot per pair instead of two (the kernel is bound
// by instruction issue).  Each lane rounds to nearest exactly like the scalar instruction, so results are bit-identical.
// CAUTION: ptxas (12.9) contracts mul.rn.f32x2 feeding add/sub.rn.f32x2 into FFMA2 even with --fmad=false; wherever
// the reference rounds a product before adding it, the product is formed with the scalar __fmul_rn below.
typedef unsigned long long wl_u64;
__device__ __forceinline__ wl_u64 pk2(float2 a) { wl_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2 upk2(wl_u64 a) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { wl_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { wl_u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { wl_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b))); return upk2(r); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  wl_u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return upk2(r);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// scalar products that must be rounded on their own (see CAUTION above)
__device__ __forceinline__ float2 mul2_sep(float2 a, float b) { return make_float2(__fmul_rn(a.x, b), __fmul_rn(a.y, b)); }
// div6 (above) for a pair: q is consumed by fma only
__device__ __forceinline__ float2 div6x2(float2 x) {
  const float2 y = splat2(1.0f / 6.0f);
  const float2 q = mul2(x, y);
  return fma2(fma2(splat2(-6.0f), q, x), y, q);
}
// flux1 (above) for the cell pair: same operations per lane
__device__ __forceinline__ float2 flux2(float2 uf, float2 um2, float2 um1, float2 uc, float2 up1, float nu) {
  float2 u, c, d;
  u.x = uf.x > 0 ? um2.x : up1.x; c.x = uf.x > 0 ? um1.x : uc.x; d.x = uf.x > 0 ? uc.x : um1.x;
  u.y = uf.y > 0 ? um2.y : up1.y; c.y = uf.y > 0 ? um1.y : uc.y; d.y = uf.y > 0 ? uc.y : um1.y;
  // q1 = (5c + 2d - u)/6: 2d is exact, so fma(2, d, RN(5c)) = RN(RN(5c) + 2d)
  const float2 q1 = div6x2(sub2(fma2(splat2(2.f), d, mul2(splat2(5.f), c)), u));
  // q2 = 10c - 9u: both products rounded on their own
  const float2 q2 = sub2(mul2_sep(c, 10.f), mul2_sep(u, 9.f));
  float2 q;
  {
    const float lo = fminf(fminf(q1.x, q2.x), d.x), hi = fmaxf(fmaxf(q1.x, q2.x), d.x);
    q.x = fminf(fmaxf(c.x, lo), hi);
  }
  {
    const float lo = fminf(fminf(q1.y, q2.y), d.y), hi = fmaxf(fmaxf(q1.y, q2.y), d.y);
    q.y = fminf(fmaxf(c.y, lo), hi);
  }
  // Φ = fma(uf, q, -(ν·(uc - um1))): the product ν·∂ is the addend of the fma (no contraction possible)
  return fma2(uf, q, mul2(sub2(uc, um1), splat2(-nu)));
}
// uf = (a + b)·0.5
__device__ __forceinline__ float2 half_sum2(float2 a, float2 b) { return mul2(add2(a, b), splat2(0.5f)); }
