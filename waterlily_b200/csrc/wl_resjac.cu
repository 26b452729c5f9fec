// wl_resjac.cu — the head of solver! on the finest level as ONE pass over HBM.
//
// solver!(ml) (MultiLevelPoisson.jl:87-99) opens with residual!(fine) (Poisson.jl:91-97) and the first Vcycle!
// (MultiLevelPoisson.jl:70-82) opens with Jacobi!(fine) (Poisson.jl:110-113) and restrict! (:33):
//
//     r = z − A x ;  ϵ = r·iD ;  r −= A ϵ ;  x += ϵ ;  r_coarse = Σ_{2×2×2} r ;  x_coarse = 0
//
// As three kernels that is 13 words per cell (read x, z / write r, ϵ ; read ϵ, r, x / write r, x ; read r).  Here a
// block marches an (128 × 8)-cell tile along z with TWO stencils in flight: plane p gets its residual while plane p-1
// gets its Jacobi increment, the residual-derived ϵ of the three planes p-2 … p living in a shared-memory ring.  The
// second stencil needs ϵ one cell beyond the tile, so the tile's halo ring (two rows of packs, two columns of cells)
// gets its residual recomputed by two warps + sixteen threads per step (the x operand is staged with a halo of two).
// HBM traffic: read x, z — write r, x' — plus 1/8 word of coarse r: 4.1 words per cell.
//
// * x' cannot be stored in place (neighbouring tiles still read the old x as their halo): it is parked in the ϵ array
//   of the level — free until the V-cycle's prolongation, whose fused pass (k_prolong_increment_init) then takes x
//   from there and writes it back to x.
// * residual!'s mean subtraction (|Σr/n| > 2eps, Poisson.jl:95-96) is a global decision that is only known when the
//   pass is over.  The pass assumes "no subtraction" (the usual case: Σr/n is the net boundary flux per cell); when
//   the device-side flag says otherwise, k_resjac_fix applies the exact linear correction
//   r −= s·(1 − A·iD), x' −= s·iD and the coarse residual is restricted again.
// * Same arithmetic per cell, in the same order, as k_residual / k_jacobi_eps / k_increment / k_restrict.
// Single-GPU (or replicated) finest level, 3-D, interior rows a multiple of 128 cells × 8 rows.
#include <algorithm>
#include <cstdlib>

#include "wl_internal.h"
#include "wl_vec.cuh"

namespace {

constexpr int RJ_MX = 32;                       // packs per tile row (128 cells)
constexpr int RJ_MY = 8;                        // rows per tile
constexpr int RJ_RW = 4 * RJ_MX + 8;            // floats per staged row: [4 pad | 128 | 4 pad]
constexpr int RJ_RWB = RJ_RW * 4;               // … in bytes
constexpr int RJ_NXS = 5;                       // x ring: planes p-1, p, p+1 + two in flight
constexpr int RJ_NZS = 3;                       // z ring: plane p + two in flight;  ϵ ring: planes p-2, p-1, p
constexpr int RJ_XPLB = (RJ_MY + 4) * RJ_RWB;   // bytes per x stage (halo of two rows)
constexpr int RJ_ZPLB = (RJ_MY + 2) * RJ_RWB;   // bytes per z / ϵ stage (halo of one row)
constexpr int RJ_XBB = (RJ_MY / 2) * 4 * RJ_MX * 4;  // restrict!: r of the odd rows handed to the even rows
constexpr int RJ_SMEM = RJ_NXS * RJ_XPLB + 2 * RJ_NZS * RJ_ZPLB + RJ_XBB;
constexpr int RJ_SMEM_DIVU = RJ_SMEM - RJ_NZS * RJ_ZPLB;  // no z ring
#ifndef RJ_PF
#define RJ_PF 1                                 // planes ahead of the L1 prefetch of the velocity field
#endif

__device__ __forceinline__ void sts4(unsigned int a, const Pack<4>& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.v[0]), "f"(v.v[1]), "f"(v.v[2]), "f"(v.v[3]) : "memory");
}
__device__ __forceinline__ void sts1(unsigned int a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <int V> __device__ __forceinline__ Pack<V> ldsp(unsigned int a);
template <> __device__ __forceinline__ Pack<4> ldsp<4>(unsigned int a) { return lds4(a); }
template <> __device__ __forceinline__ Pack<1> ldsp<1>(unsigned int a) { Pack<1> r; r.v[0] = lds1(a); return r; }

// coordinate v of a cell that may lie up to two cells outside the interior 1 … n-2: its periodic image, or clamped
// into the array (such a cell is then "dead": its residual is zero by definition, the data read is never used)
__device__ __forceinline__ int rj_wrap(int v, int n, int per) {
  if (per) return v < 1 ? v + (n - 2) : (v > n - 2 ? v - (n - 2) : v);
  return v < 0 ? 0 : (v > n - 1 ? n - 1 : v);
}
__device__ __forceinline__ bool rj_dead(int v, int n, int per) { return !per && (v < 1 || v > n - 2); }

// residual of V cells at tile position (c, rr) of plane p: r = iD == 0 ? 0 : z − A x (Poisson.jl:93), e = r·iD (:111).
// xm/xc/xp: byte addresses of the x stages of planes p-1, p, p+1; zc: the z stage of plane p; Iw: (wrapped) cell index
// for the coefficient loads of a non-uniform pack.
// zc: the z stage of plane p, or — divu given — z = div u (Flow.jl:11-17, 138) is evaluated here from the velocity field
// (its lines were prefetched into L1 one plane earlier), with the operations of k_project_pre.
template <int V>
__device__ __forceinline__ void rj_residual(const Lvl& l, unsigned int xm, unsigned int xc, unsigned int xp, unsigned int zc, const float* __restrict__ divu, int c,
                                            int rr, int64_t Iw, bool uni, bool dead, Pack<V>& r, Pack<V>& e) {
  const unsigned int xo = (unsigned int)(((rr + 2) * RJ_RW + 4 + c) * 4), zo = (unsigned int)(((rr + 1) * RJ_RW + 4 + c) * 4);
  Pack<V> zz;
  if (divu) {
    if (!dead) zz = div_pack<3, V>(l.g, divu, Iw);
  } else zz = ldsp<V>(zc + zo);
  const Pack<V> x0 = ldsp<V>(xc + xo);
  const float xm1 = lds1(xc + xo - 4), xp4 = lds1(xc + xo + 4 * V);
  const Pack<V> ym = ldsp<V>(xc + xo - RJ_RWB), yp = ldsp<V>(xc + xo + RJ_RWB), zm = ldsp<V>(xm + xo), zp = ldsp<V>(xp + xo);
  if (dead) {
#pragma unroll
    for (int q = 0; q < V; ++q) { r.v[q] = 0.f; e.v[q] = 0.f; }
    return;
  }
  const Pack<V> ax = mult_vals<V>(l, Iw, x0, xm1, xp4, ym, yp, zm, zp, uni);
  Pack<V> id;
  if (uni) id = splat<V>(l.uiD); else id = ldp<V>(l.iD + Iw);
#pragma unroll
  for (int q = 0; q < V; ++q) {
    r.v[q] = id.v[q] == 0.f ? 0.f : zz.v[q] - ax.v[q];
    e.v[q] = r.v[q] * id.v[q];
  }
}

// project!'s x .*= dt (Flow.jl:139) applied to a staged value: dt is Float32 for w = 1, Float64 for w = 0.5
__device__ __forceinline__ float rj_mul(float v, int op, float m32, double m64) { return op == 1 ? v * m32 : (float)((double)v * m64); }

// xin: where the pressure is read from — the x array, or (mul_op != 0, inside mom_step!) the array that holds the divided
// pressure of the last project!, which gets the new project!'s x .*= dt while it is staged.  xout ≠ xin.
template <bool DIVU, int MULOP>
__global__ void __launch_bounds__(256, 3) k_resjac(Lvl l, Lvl cl, const float* __restrict__ xin, float* __restrict__ xout, float mul32, double mul64,
                                                    const float* __restrict__ divu_, MarchGeom mg, ReduceWS ws, FinArgs fa,
                                                    const unsigned int* __restrict__ nonuni) {
  extern __shared__ __align__(16) float rj_smem[];
  const float* __restrict__ const divu = DIVU ? divu_ : nullptr;  // compile-time switches: no registers spent on them
  constexpr int mul_op = MULOP;
  const GridD& g = l.g;
  const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
  const int64_t s1 = g.s[1], s2 = g.s[2];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  // (no z ring when z is evaluated from divu: the launch then asks for less shared memory, which stays with the L1)
  const unsigned int XR = wl_smem_u32(rj_smem), ZR = XR + RJ_NXS * RJ_XPLB, ER = divu ? ZR : ZR + RJ_NZS * RJ_ZPLB, XB = ER + RJ_NZS * RJ_ZPLB;
  // every pack of the level uniform (no body, no wall: the periodic benchmark): no mask traffic at all
  const bool has_mask = l.umask != nullptr;
  const bool all_uni = has_mask && *nonuni == 0u;
  double acc[2] = {0.0, 0.0};
  const int ntiles = mg.ntx * mg.nty * mg.ntz;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int ttx = t % mg.ntx, tq = t / mg.ntx, tty = tq % mg.nty, ttz = tq / mg.nty;
    const int i0 = 1 + 4 * RJ_MX * ttx, j0 = 1 + RJ_MY * tty;
    const int k0 = 1 + ttz * mg.zchunk, k1 = min(k0 + mg.zchunk - 1, n2 - 2);
    const int i = i0 + 4 * tx, j = j0 + ty;
    // ---- staging geometry: offsets inside a plane (wrapped / clamped) and destinations inside a stage ----
    const int o_own = i + j * (int)s1;
    int o_xh = 0, o_zh = 0;
    unsigned int d_xh = 0, d_zh = 0;
    if (tid < 128) {         // x: halo rows -2, -1, 8, 9 (32 packs each)
      const int h = tid >> 5, rr = h < 2 ? h - 2 : RJ_MY + h - 2;
      o_xh = (i0 + 4 * tx) + rj_wrap(j0 + rr, n1, g.per[1]) * (int)s1;
      d_xh = (unsigned int)(((rr + 2) * RJ_RW + 4 + 4 * tx) * 4);
    } else if (tid < 168) {  // x: halo columns -2, -1, 128, 129 of rows -1 … 8
      const int q = tid - 128, h = q / 10, rr = q - h * 10 - 1, c = h < 2 ? h - 2 : 4 * RJ_MX + h - 2;
      o_xh = rj_wrap(i0 + c, n0, g.per[0]) + rj_wrap(j0 + rr, n1, g.per[1]) * (int)s1;
      d_xh = (unsigned int)(((rr + 2) * RJ_RW + 4 + c) * 4);
    }
    if (tid < 64) {          // z: halo rows -1, 8
      const int h = tid >> 5, rr = h == 0 ? -1 : RJ_MY;
      o_zh = (i0 + 4 * tx) + rj_wrap(j0 + rr, n1, g.per[1]) * (int)s1;
      d_zh = (unsigned int)(((rr + 1) * RJ_RW + 4 + 4 * tx) * 4);
    } else if (tid < 80) {   // z: halo columns -1, 128 of rows 0 … 7
      const int q = tid - 64, h = q >> 3, rr = q & 7, c = h == 0 ? -1 : 4 * RJ_MX;
      o_zh = rj_wrap(i0 + c, n0, g.per[0]) + (j0 + rr) * (int)s1;
      d_zh = (unsigned int)(((rr + 1) * RJ_RW + 4 + c) * 4);
    }
    const unsigned int d_xown = (unsigned int)(((ty + 2) * RJ_RW + 4 + 4 * tx) * 4), d_zown = (unsigned int)(((ty + 1) * RJ_RW + 4 + 4 * tx) * 4);
    auto issue_x = [&](int P, unsigned int stage) {
      const float* b = xin + (int64_t)rj_wrap(P, n2, g.per[2]) * s2;
      cp_async16s(stage + d_xown, b + o_own);
      if (tid < 128) cp_async16s(stage + d_xh, b + o_xh);
      else if (tid < 168) cp_async4s(stage + d_xh, b + o_xh);
    };
    // x .*= dt on what this thread itself copied into a stage (its cp.async group has completed: visible to the thread)
    auto scale_x = [&](unsigned int stage) {
      Pack<4> v = lds4(stage + d_xown);
#pragma unroll
      for (int q = 0; q < 4; ++q) v.v[q] = rj_mul(v.v[q], mul_op, mul32, mul64);
      sts4(stage + d_xown, v);
      if (tid < 128) {
        Pack<4> h = lds4(stage + d_xh);
#pragma unroll
        for (int q = 0; q < 4; ++q) h.v[q] = rj_mul(h.v[q], mul_op, mul32, mul64);
        sts4(stage + d_xh, h);
      } else if (tid < 168) {
        sts1(stage + d_xh, rj_mul(lds1(stage + d_xh), mul_op, mul32, mul64));
      }
    };
    auto issue_z = [&](int P, unsigned int stage) {
      const float* b = l.z + (int64_t)rj_wrap(P, n2, g.per[2]) * s2;
      cp_async16s(stage + d_zown, b + o_own);
      if (tid < 64) cp_async16s(stage + d_zh, b + o_zh);
      else if (tid < 80) cp_async4s(stage + d_zh, b + o_zh);
    };
    // ---- halo work of the R phase: two warps take a halo row of packs each, sixteen threads the column cells ----
    // (which warps: rotated with the plane, so that the extra work is spread over the block)
    // row packs: lane = pack; column cells: lanes 0-7 left column, 8-15 right column
    const int lane = tx;
    // ---- prologue: x planes k0-2 … k0+1 (stages 0 … 3), z planes k0-1, k0 (stages 0, 1) ----
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      issue_x(k0 - 2 + q, XR + q * RJ_XPLB);
      if (q >= 2 && !divu) issue_z(k0 - 3 + q, ZR + (q - 2) * RJ_ZPLB);
      cp_async_commit_group();
    }
    unsigned int xa_m = XR, xa_c = XR + RJ_XPLB, xa_p = XR + 2 * RJ_XPLB, xa_i = XR + 4 * RJ_XPLB;  // x stages of planes s-1, s, s+1, s+3
    int zs = 0;                                                                                     // z / ϵ stage of plane s
    Pack<4> r_prev = splat<4>(0.f), rj = splat<4>(0.f);     // residual of the own pack on plane s-1; own r' waiting for the partner row (restrict!)
    float ca0 = 0.f, ca1 = 0.f;
    int pend = 0;           // plane whose r' is waiting in rj / XB (0: none)
    bool uni_prev = all_uni;
    unsigned int m_next = 1u;
    const unsigned char* mbase = has_mask ? l.umask + ((o_own + g.lead) >> 2) : nullptr;  // + plane·(s2/4)
    const int64_t mstp = s2 >> 2;
    if (has_mask && !all_uni) m_next = mbase[(int64_t)rj_wrap(k0 - 1, n2, g.per[2]) * mstp];
#pragma unroll 1
    for (int s = k0 - 1; s <= k1 + 1; ++s) {
      const unsigned int m_cur = m_next;
      if (has_mask && !all_uni && s <= k1) m_next = mbase[(int64_t)rj_wrap(s + 1, n2, g.per[2]) * mstp];
      cp_async_wait_group<1>();   // x planes ≤ s+1 and z planes ≤ s have landed
      if (mul_op) {
        if (s == k0 - 1) { scale_x(xa_m); scale_x(xa_c); }
        scale_x(xa_p);            // plane s+1, complete since this wait
      }
      __syncthreads();            // … for every thread; everyone is done with the stages about to be refilled
      if (s + 3 <= k1 + 2) issue_x(s + 3, xa_i);
      if (s + 2 <= k1 + 1 && !divu) issue_z(s + 2, ZR + (zs == 0 ? 2 : zs - 1) * RJ_ZPLB);
      cp_async_commit_group();
      if (divu && s + RJ_PF <= k1 + 1) {
        // the velocity lines the R phase of plane s+RJ_PF will read: own pack of each component; the tile's edge threads
        // add the halo ring (rows -1, 8 and u₂'s row 9; the packs left and right of the tile).  u₃ of the plane above
        // is the own pack of the following step.  Addresses of cells outside the array are clamped: never used.
        const int kp = rj_wrap(s + RJ_PF, n2, g.per[2]);
        const float* q = divu + o_own + (int64_t)kp * s2;
        prefetch_l1(q); prefetch_l1(q + g.comp); prefetch_l1(q + 2 * g.comp);
        if (s + RJ_PF == k1 + 1) prefetch_l1(q + 2 * g.comp + s2);
        if (ty == 0 || ty == RJ_MY - 1) {
          const int jw = rj_wrap(ty == 0 ? j0 - 1 : j0 + RJ_MY, n1, g.per[1]);
          const float* qr = divu + (i0 + 4 * tx) + jw * s1 + (int64_t)kp * s2;
          prefetch_l1(qr); prefetch_l1(qr + g.comp); prefetch_l1(qr + 2 * g.comp);
          if (ty == RJ_MY - 1) prefetch_l1(qr + g.comp + s1);
        }
        if (tx == 0 || tx == 31) {
          const float* qc = divu + rj_wrap(tx == 0 ? i0 - 1 : i0 + 4 * RJ_MX, n0, g.per[0]) + j * s1 + (int64_t)kp * s2;
          prefetch_l1(qc); prefetch_l1(qc + g.comp); prefetch_l1(qc + 2 * g.comp);
        }
      }
      // ---- restrict! of the plane finished in the previous step (MultiLevelPoisson.jl:33: Σ over the 2×2×2 block, x fastest) ----
      if (pend && (ty & 1) == 0) {
        const Pack<4> pr = lds4(XB + (unsigned int)(((ty >> 1) * 4 * RJ_MX + 4 * tx) * 4));
        if (((pend - k0) & 1) == 0) { ca0 = 0.f; ca1 = 0.f; }
        ca0 = (((ca0 + rj.v[0]) + rj.v[1]) + pr.v[0]) + pr.v[1];
        ca1 = (((ca1 + rj.v[2]) + rj.v[3]) + pr.v[2]) + pr.v[3];
        if (((pend - k0) & 1) == 1) {
          const int64_t Ic = ((i + 1) >> 1) + (int64_t)((j + 1) >> 1) * cl.g.s[1] + (int64_t)((pend + 1) >> 1) * cl.g.s[2];
          *reinterpret_cast<float2*>(cl.r + Ic) = make_float2(ca0, ca1);
          *reinterpret_cast<float2*>(cl.x + Ic) = make_float2(0.f, 0.f);
        }
      }
      // ---- R phase, plane p = s: residual and ϵ of the tile and of its halo ring → ϵ ring ----
      const int kw = rj_wrap(s, n2, g.per[2]);
      const bool pdead = rj_dead(s, n2, g.per[2]);
      const unsigned int zst = ZR + zs * RJ_ZPLB, est = ER + zs * RJ_ZPLB;
      Pack<4> r_cur;
      bool uni_cur;
      {
        uni_cur = all_uni || (has_mask && m_cur == 0u);
        Pack<4> e;
        rj_residual<4>(l, xa_m, xa_c, xa_p, zst, divu, 4 * tx, ty, o_own + (int64_t)kw * s2, uni_cur, pdead, r_cur, e);
        sts4(est + d_zown, e);
        if (s >= k0 && s <= k1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) { acc[0] += (double)r_cur.v[q]; acc[1] = dfma(r_cur.v[q], r_cur.v[q], acc[1]); }
        }
      }
      const int hw = (s & 3) * 2;  // the two warps that take the halo rows of this plane
      if (ty == hw || ty == hw + 1) {
        const int rr = ty == hw ? -1 : RJ_MY, jw = rj_wrap(j0 + rr, n1, g.per[1]);
        const int64_t Iw = (i0 + 4 * lane) + jw * s1 + (int64_t)kw * s2;
        const bool dead = pdead || rj_dead(j0 + rr, n1, g.per[1]);
        const bool uni = all_uni || (has_mask && !dead && l.umask[(Iw + g.lead) >> 2] == 0);
        Pack<4> r, e;
        rj_residual<4>(l, xa_m, xa_c, xa_p, zst, divu, 4 * lane, rr, Iw, uni, dead, r, e);
        sts4(est + (unsigned int)(((rr + 1) * RJ_RW + 4 + 4 * lane) * 4), e);
      } else if (ty == ((hw + 2) & 7) && lane < 16) {
        const int h = lane >> 3, rr = lane & 7, c = h == 0 ? -1 : 4 * RJ_MX, iw = rj_wrap(i0 + c, n0, g.per[0]);
        const int64_t Iw = iw + (j0 + rr) * s1 + (int64_t)kw * s2;
        const bool dead = pdead || rj_dead(i0 + c, n0, g.per[0]);
        const bool uni = all_uni || (has_mask && !dead && l.umask[(Iw - ((iw - 1) & 3) + g.lead) >> 2] == 0);
        Pack<1> r, e;
        rj_residual<1>(l, xa_m, xa_c, xa_p, zst, divu, c, rr, Iw, uni, dead, r, e);
        sts1(est + (unsigned int)(((rr + 1) * RJ_RW + 4 + c) * 4), e.v[0]);
      }
      __syncthreads();  // ϵ of plane s complete
      // ---- J phase, plane q = s-1: r −= Aϵ, x += ϵ (Poisson.jl:101-102) ----
      if (s - 1 >= k0) {
        const int q = s - 1;
        const unsigned int e_c = ER + (zs == 0 ? 2 : zs - 1) * RJ_ZPLB, e_m = ER + (zs == 2 ? 0 : zs + 1) * RJ_ZPLB, e_p = est;  // planes q, q-1, q+1
        const unsigned int eo = d_zown;
        const Pack<4> ec = lds4(e_c + eo);
        const float em1 = lds1(e_c + eo - 4), ep4 = lds1(e_c + eo + 16);
        const Pack<4> eym = lds4(e_c + eo - RJ_RWB), eyp = lds4(e_c + eo + RJ_RWB), ezm = lds4(e_m + eo), ezp = lds4(e_p + eo);
        const Pack<4> xq = lds4(xa_m + d_xown);
        const int64_t Iq = o_own + (int64_t)q * s2;
        const Pack<4> ae = mult_vals<4>(l, Iq, ec, em1, ep4, eym, eyp, ezm, ezp, uni_prev);
        Pack<4> rn, xn;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          rn.v[w] = r_prev.v[w] - ae.v[w];
          xn.v[w] = xq.v[w] + ec.v[w];
        }
        stp<4>(l.r + Iq, rn);
        stp<4>(xout + Iq, xn);
        if (ty & 1) sts4(XB + (unsigned int)(((ty >> 1) * 4 * RJ_MX + 4 * tx) * 4), rn);
        else rj = rn;
        pend = q;
      }
      r_prev = r_cur;
      uni_prev = uni_cur;
      // rotate the rings
      xa_m = xa_c; xa_c = xa_p;
      xa_p += RJ_XPLB; if (xa_p == XR + RJ_NXS * RJ_XPLB) xa_p = XR;
      xa_i += RJ_XPLB; if (xa_i == XR + RJ_NXS * RJ_XPLB) xa_i = XR;
      zs = zs == 2 ? 0 : zs + 1;
    }
    cp_async_wait_group<0>();
    __syncthreads();  // the partner rows of the last plane are in XB; the rings are free for the next tile
    if (pend && (ty & 1) == 0) {
      const Pack<4> pr = lds4(XB + (unsigned int)(((ty >> 1) * 4 * RJ_MX + 4 * tx) * 4));
      if (((pend - k0) & 1) == 0) { ca0 = 0.f; ca1 = 0.f; }
      ca0 = (((ca0 + rj.v[0]) + rj.v[1]) + pr.v[0]) + pr.v[1];
      ca1 = (((ca1 + rj.v[2]) + rj.v[3]) + pr.v[2]) + pr.v[3];
      if (((pend - k0) & 1) == 1) {
        const int64_t Ic = ((i + 1) >> 1) + (int64_t)((j + 1) >> 1) * cl.g.s[1] + (int64_t)((pend + 1) >> 1) * cl.g.s[2];
        *reinterpret_cast<float2*>(cl.r + Ic) = make_float2(ca0, ca1);
        *reinterpret_cast<float2*>(cl.x + Ic) = make_float2(0.f, 0.f);
      }
    }
    __syncthreads();  // XB is rewritten by the next tile
  }
  double tot[2];
  if (grid_reduce<2>(acc, ws, tot)) finish_reduction(FIN_RESIDUAL, fa, tot[0], tot[1]);
}

// residual! found |Σr/n| > 2eps after the fused pass ran without the subtraction: with s = Σr/n the unfused sequence
// gives ϵ = (r−s)·iD, r' = (r−s) − A((r−s)·iD), x' = x + (r−s)·iD, i.e. the fused results corrected by
// r' −= s·(1 − A·iD), x' −= s·iD (A·iD = 0 wherever the coefficients are uniform).
template <int D, int V>
__global__ void __launch_bounds__(256) k_resjac_fix(Lvl l, float* __restrict__ xpark) {
  if (!l.sc->need_sub) return;
  const float s = l.sc->rsum;
  for_rows<V>(l.g, [&](int i, int j, int k, int64_t I) {
    const bool uni = pack_uniform<V>(l, I);
    const Pack<V> idc = ldp<V>(l.iD + I), a = mult_w<D, V>(l, l.iD, i, j, k, I, idc, uni);
    Pack<V> r = ldp<V>(l.r + I), x = ldp<V>(xpark + I);
#pragma unroll
    for (int q = 0; q < V; ++q) {
      r.v[q] = (r.v[q] - s) + s * a.v[q];
      x.v[q] = x.v[q] - s * idc.v[q];
    }
    stp<V>(l.r + I, r);
    stp<V>(xpark + I, x);
  });
}
// restrict! again after the correction (MultiLevelPoisson.jl:33); predicated on the same flag
__global__ void __launch_bounds__(256) k_resjac_rerestrict(Lvl c, Lvl f) {
  if (!f.sc->need_sub) return;
  for_rows<1>(c.g, [&](int i0, int j0, int k0, int64_t I) {
    float s = 0.f;
    for (int z = 2 * k0 - 1; z <= 2 * k0; ++z)
      for (int y = 2 * j0 - 1; y <= 2 * j0; ++y)
        for (int x = 2 * i0 - 1; x <= 2 * i0; ++x) s += f.r[x + y * f.g.s[1] + z * f.g.s[2]];
    c.r[I] = s;
  });
}

template <int D, int V>
__global__ void __launch_bounds__(256) k_resjac_unpark(Lvl l) {
  for_rows<V>(l.g, [&](int, int, int, int64_t I) { stp<V>(l.x + I, ldp<V>(l.eps + I)); });
}

}  // namespace

int resjac_unpark(Sim& s) {
  Lvl& l = s.levels[0];
  RLAUNCH(s, k_resjac_unpark, l.g, l);
  return wl_cuda_check(s, cudaGetLastError(), "k_resjac_unpark");
}

// can the finest level of this handle run the fused head?  (geometry only; the caller checks the solver state)
bool resjac_eligible(const Sim& s) {
  if (!s.resjac_on || !s.march_on || !s.stage_on || s.D != 3 || s.levels.size() < 2) return false;
  if (s.dist.nranks > 1 && s.lvl_dist[0]) return false;   // the second stencil would need a pressure halo of depth two
  const GridD& g = s.levels[0].g;
  const int nxi = g.n[0] - 2, nyi = g.n[1] - 2, nzi = g.n[2] - 2;
  if (nxi % (4 * RJ_MX) || nyi % RJ_MY || (nzi & 1) || nzi < 8) return false;
  const GridD& c = s.levels[1].g;
  return c.n[0] - 2 == nxi / 2 && c.n[1] - 2 == nyi / 2 && c.n[2] - 2 == nzi / 2;
}

// residual!(fine) + Jacobi!(fine) + restrict!(fine → coarse) + fill!(coarse.x, 0); z holds the source term (interior), or —
// divu given (project! inside mom_step!) — it is div(divu), evaluated on the fly and never stored.
// Leaves r' in fine.r, the coarse residual in levels[1].r, and x' PARKED in fine.ϵ (Sim::x_parked) — or, when project! asked
// for its x .*= dt to be applied on the way (Sim::xmul_op: x is then read from fine.ϵ, see Sim::xdiv_valid), in fine.x.
int resjac_run(Sim& s, const float* divu) {
  Lvl& l = s.levels[0];
  Lvl& c = s.levels[1];
  const GridD& g = l.g;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_resjac<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resjac<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resjac<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resjac<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SMEM_DIVU);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resjac<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SMEM_DIVU);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resjac<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RJ_SMEM_DIVU);
    if (e != cudaSuccess) return wl_cuda_check(s, e, "cudaFuncSetAttribute(k_resjac)");
    attr_done = true;
  }
  MarchGeom mg;
  mg.mx = RJ_MX; mg.my = RJ_MY;
  mg.ntx = (g.n[0] - 2) / (4 * RJ_MX);
  mg.nty = (g.n[1] - 2) / RJ_MY;
  const int nzi = g.n[2] - 2;
  static int zc_env = -1;
  if (zc_env < 0) { const char* e = getenv("WL_RJ_ZCHUNK"); zc_env = e ? std::max(8, atoi(e)) : 64; }
  int zc = zc_env;
  while (zc > 8 && (int64_t)mg.ntx * mg.nty * ((nzi + zc - 1) / zc) < (int64_t)s.nsm * 6) zc >>= 1;
  mg.zchunk = zc;
  mg.ntz = (nzi + zc - 1) / zc;
  const int64_t ntiles = (int64_t)mg.ntx * mg.nty * mg.ntz;
  const int grid = (int)std::min<int64_t>(ntiles, WL_MAXBLOCKS);
  const FinArgs fa = fin_level(s, 0);
  wl_prof_begin(s, WL_PROF_RESIDUAL);
  const bool from_div = s.xmul_op != 0;
  const float* xin = from_div ? l.eps : l.x;
  float* xout = from_div ? l.x : l.eps;
#define RJ_LAUNCH(DV, MO) k_resjac<DV, MO><<<grid, 256, DV ? RJ_SMEM_DIVU : RJ_SMEM, s.stream>>>(l, c, xin, xout, s.xmul32, s.xmul64, divu, mg, s.ws, fa, s.d_nonuni)
  if (divu) { if (s.xmul_op == 0) RJ_LAUNCH(true, 0); else if (s.xmul_op == 1) RJ_LAUNCH(true, 1); else RJ_LAUNCH(true, 2); }
  else { if (s.xmul_op == 0) RJ_LAUNCH(false, 0); else if (s.xmul_op == 1) RJ_LAUNCH(false, 1); else RJ_LAUNCH(false, 2); }
#undef RJ_LAUNCH
  s.xmul_op = 0;
  s.x_parked = !from_div;
  ++s.launches;
  int rc = wl_cuda_check(s, cudaGetLastError(), "k_resjac");
  if (rc) return rc;
  if ((rc = post_reduce(s, FIN_RESIDUAL, fa))) return rc;
  RLAUNCH(s, k_resjac_fix, l.g, l, xout);
  {
    const RowLaunch rl = row_launch(c.g);
    k_resjac_rerestrict<<<wl_grid_for(s, rl.nunits), rl.blk, 0, s.stream>>>(c, l);
    ++s.launches;
  }
  // algorithmic words per cell: read x, z (or the three velocity components), write r, x', ⅛ coarse r
  wl_prof_end(s, WL_PROF_RESIDUAL, 4.0 * (double)g.interior * (divu ? 6.125 : 4.125));
  return wl_cuda_check(s, cudaGetLastError(), "k_resjac_fix");
}
