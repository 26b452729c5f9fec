// wl_convdiff_tma.cu — conv_diff! (+ accelerate! + BDIM! stage 1) for the 3-D core, TMA-staged.
// Reference: src/Flow.jl:36-60 (conv_diff!), :68-73 (accelerate!), :133 (BDIM! stage 1).
//
// The kernel is bound by instruction issue on the ALU pipe (min/max/select and integer adds run
// there at half rate), not by HBM: the first marching kernel (wl_momentum.cu) spent ~40 % of its
// ALU slots on 64-bit address arithmetic for its ~45 neighbour loads per cell.  Here a block owns a
// TX×TY (x,y) tile and marches along z, and the three velocity components of every plane it needs
// arrive — with their 2-cell halo — through ONE TMA tensor copy per plane (cp.async.bulk.tensor on
// a 4-D map x,y,z,component) into a 4-stage shared-memory ring.  All neighbour reads are then LDS
// with compile-time offsets, there is no halo logic in the threads, and the copy of plane z+4 is in
// flight while plane z is being differenced.
//
//  * each thread keeps its own column's z-window (4 samples × 3 components) in registers;
//  * every face flux is evaluated once: the z-flux of the upper face is carried to the next step,
//    x-fluxes travel between lanes by shuffle (between warps through shared memory), y-fluxes
//    between warps through shared memory (double-buffered, one barrier per step);
//  * the z-loop is unrolled by the ring depth so that every shared-memory offset is an immediate.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "wl_internal.h"
#include "wl_flux.cuh"

namespace {

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WL_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WL_MBAR_DONE;\n"
      "bra WL_MBAR_WAIT;\n"
      "WL_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one plane (box PX × PY × 1 × 3) of the velocity field → shared memory
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(0), "r"(smem_u32(bar))
      : "memory");
}

template <int TX, int TY>
struct Tile {
  // a TMA box must start on a 16-byte boundary in x (measured on B200: any other innermost coordinate
  // faults with "illegal instruction"): the box starts WL_OX cells left of the 2-cell halo, where the
  // rows of the padded layout are aligned, and is TX+8 wide
  static constexpr int PX = TX + 8, PY = TY + 4, PP = PX * PY, STAGE = 3 * PP;  // floats
  static constexpr unsigned int BYTES = STAGE * sizeof(float);
};

template <int TX, int TY, int MINB, bool HASG, bool CORR>
__global__ void __launch_bounds__(TX * TY, MINB)
    k_conv_diff3_tma(const __grid_constant__ CUtensorMap tmU, CDArgs a, int zchunk, int ntx, int nty) {
  using T = Tile<TX, TY>;
  constexpr int PX = T::PX, PP = T::PP;
  __shared__ __align__(128) float tile[4][T::STAGE];
  __shared__ float Fx[2][3][TY][TX / 32];
  __shared__ float Fy[2][3][TY][TX];
  __shared__ float FxE[2][3][TY];   // fluxes through the x-face beyond the tile's last column (see below)
  __shared__ float UxE[2][TY];      // u_x at that face on the previous plane
  __shared__ __align__(8) unsigned long long full[4];
  const GridD& g = a.g;
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  const int64_t s1 = g.s[1], s2 = g.s[2];
  const int tix = threadIdx.x, tiy = threadIdx.y, lane = tix & 31, wx = tix >> 5;
  const bool leader = (tix | tiy) == 0;
  const int bt = blockIdx.x;
  const int tx = bt % ntx, ty = (bt / ntx) % nty, tz = bt / (ntx * nty);
  const int xb = tx * TX, yb = ty * TY;  // tile origin (cell index of the first halo cell)
  const int x = 2 + xb + tix, y = 2 + yb + tiy;
  const int zlo = g.zlo_halo ? 1 : 2, zhi = g.zhi_halo ? nz - 2 : nz - 3;  // slab faces are plain inner cells
  const int z0 = zlo + tz * zchunk, z1 = min(z0 + zchunk, zhi + 1);
  const bool valid = x <= nx - 3 && y <= ny - 3;
  const int xc = min(x, nx - 3), yc = min(y, ny - 3);  // out-of-range threads shadow the last cell
  const bool extra_x = (xc == nx - 3) || (tix == TX - 1);
  const bool extra_y = (yc == ny - 3) || (tiy == TY - 1);
  const float nu = a.nu;
  const int toff = (yc - yb) * PX + (xc - xb) + WL_OX;  // this thread's cell inside a staged plane
  const int zb = z0 - 2;                        // plane held by ring stage 0 first
  // The x-face beyond the tile's last column belongs to a cell no thread of the block owns.  Its 3·TY
  // fluxes per plane are evaluated together by ONE warp (lane = 3·row + component) instead of three
  // serial evaluations in every warp that holds a last-column thread.
  const bool edge_warp = tiy == 0 && wx == 0;
  const int erow = min(lane / 3, TY - 1), ecmp = lane - 3 * (lane / 3);
  const int xe = min(2 + xb + TX, nx - 2);                                         // cell whose lower x-face it is
  const int eoff = (min(2 + yb + erow, ny - 3) - yb) * PX + (xe - xb) + WL_OX;   // that cell in a staged plane (component 0)
  // tensor coordinates of cell (i,j,k): (i + WL_OX, j, k + 1) — the map starts at the allocation.
  // Box origin: cell (xb - WL_OX, yb) ↔ tensor x = xb, a multiple of 4 floats.
  const int cx = xb, cy = yb;
  if (leader) {
#pragma unroll
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (leader) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      mbar_expect_tx(&full[s], T::BYTES);
      tma_load_plane(tile[s], &tmU, &full[s], cx, cy, zb + s + 1);
    }
  }
  // byte offset of this thread's cell in f, V, u⁰ (plain coalesced accesses; the per-component bases are uniform)
  int64_t off = (xc + yc * s1 + (int64_t)z0 * s2) * (int64_t)sizeof(float);
  const int64_t doff = s2 * (int64_t)sizeof(float);
  const char* const fb[3] = {(const char*)a.f, (const char*)(a.f + g.comp), (const char*)(a.f + 2 * g.comp)};
  const char* const vb[3] = {(const char*)a.V, (const char*)(a.V + g.comp), (const char*)(a.V + 2 * g.comp)};
  const char* const ub[3] = {(const char*)a.u0, (const char*)(a.u0 + g.comp), (const char*)(a.u0 + 2 * g.comp)};
  // z-window of the own column: plane P sits in slot (P - zb) & 3, the slot number of its smem stage, so
  // the window never has to be shifted (the loop is unrolled by 4 and every slot index is static)
  float R[3][4];
  float Fz[2][3];  // z-flux carried from one step to the next, double-buffered by step parity
  float cyp;  // u_y(y+1) of the previous plane (upper-edge fluxes only)
  // ---- pre-step at z0-1 (stage 1): the z-flux through the lower face of the first cell ----
#pragma unroll
  for (int s = 0; s < 4; ++s) mbar_wait(&full[s], 0);
  {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 4; ++k) R[c][k] = tile[k][c * PP + toff];
    }
    const float* C = tile[1] + toff;  // plane z0-1
    const float* N = tile[2] + toff;  // plane z0
    if (edge_warp && lane < 3 * TY && ecmp == 2) UxE[0][erow] = tile[1][eoff];
    cyp = C[PP + PX];
    const float uz = R[2][2];
    const float ufx = (uz + N[2 * PP - 1]) * 0.5f, ufy = (uz + N[2 * PP - PX]) * 0.5f, ufz = (uz + R[2][1]) * 0.5f;
    Fz[0][0] = flux1(ufx, R[0][0], R[0][1], R[0][2], R[0][3], nu);
    Fz[0][1] = flux1(ufy, R[1][0], R[1][1], R[1][2], R[1][3], nu);
    Fz[0][2] = flux1(ufz, R[2][0], R[2][1], R[2][2], R[2][3], nu);
  }
  __syncthreads();
  if (leader) {  // planes z0-2 and z0-1 are spent: refill their stages with z0+2 and z0+3
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (zb + s + 4 <= z1 + 1) {
        mbar_expect_tx(&full[s], T::BYTES);
        tma_load_plane(tile[s], &tmU, &full[s], cx, cy, zb + s + 4 + 1);
      }
    }
  }
  int buf = 0;
  // one step of the march: cell plane z lives in ring stage S, z+1 in S+1, z+2 in S+2
  auto step = [&](auto stage_tag, int z) {
    constexpr int S = decltype(stage_tag)::value;
    constexpr int SM1 = (S + 3) & 3, SP1 = (S + 1) & 3, SP2 = (S + 2) & 3;  // window slots of z-1, z+1, z+2
    const float* C = tile[S] + toff;
    const float* N = tile[SP1] + toff;
    const float* NN = tile[SP2] + toff;
    float xm2[3], xm1[3], xp1[3], ym2[3], ym1[3], yp1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xm2[c] = C[c * PP - 2]; xm1[c] = C[c * PP - 1]; xp1[c] = C[c * PP + 1];
      ym2[c] = C[c * PP - 2 * PX]; ym1[c] = C[c * PP - PX]; yp1[c] = C[c * PP + PX];
    }
    float fx_lo[3], fy_lo[3], fx_hi[3], fy_hi[3];
    {
      const float ux = R[0][S];
      fx_lo[0] = flux1((ux + xm1[0]) * 0.5f, xm2[0], xm1[0], R[0][S], xp1[0], nu);
      fx_lo[1] = flux1((ux + ym1[0]) * 0.5f, xm2[1], xm1[1], R[1][S], xp1[1], nu);
      fx_lo[2] = flux1((ux + R[0][SM1]) * 0.5f, xm2[2], xm1[2], R[2][S], xp1[2], nu);
      const float uy = R[1][S];
      fy_lo[0] = flux1((uy + xm1[1]) * 0.5f, ym2[0], ym1[0], R[0][S], yp1[0], nu);
      fy_lo[1] = flux1((uy + ym1[1]) * 0.5f, ym2[1], ym1[1], R[1][S], yp1[1], nu);
      fy_lo[2] = flux1((uy + R[1][SM1]) * 0.5f, ym2[2], ym1[2], R[2][S], yp1[2], nu);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (lane == 0) Fx[buf][i][tiy][wx] = fx_lo[i];
      Fy[buf][i][tiy][tix] = fy_lo[i];
    }
    // lower faces of the cells x+1 / y+1 that no thread of the block owns
    float ey[3] = {0.f, 0.f, 0.f};
    if (edge_warp) {
      const float* E = tile[S] + eoff;   // u_x at the edge cell; component c of the same cell is E[c*PP]
      const float* Ei = E + ecmp * PP;
      const float ux = E[0];
      const float other = ecmp == 0 ? E[-1] : (ecmp == 1 ? E[-PX] : UxE[buf][erow]);
      const float fe = flux1((ux + other) * 0.5f, Ei[-2], Ei[-1], Ei[0], Ei[1], nu);
      if (lane < 3 * TY) {
        FxE[buf][ecmp][erow] = fe;
        if (ecmp == 2) UxE[buf ^ 1][erow] = ux;
      }
    }
    if (extra_y) {
      const float uy = yp1[1];
      ey[0] = flux1((uy + C[PP + PX - 1]) * 0.5f, ym1[0], R[0][S], yp1[0], C[2 * PX], nu);
      ey[1] = flux1((uy + R[1][S]) * 0.5f, ym1[1], R[1][S], yp1[1], C[PP + 2 * PX], nu);
      ey[2] = flux1((uy + cyp) * 0.5f, ym1[2], R[2][S], yp1[2], C[2 * PP + 2 * PX], nu);
    }
    cyp = yp1[1];
    __syncthreads();
    // every read of plane z is done: its stage takes plane z+4
    if (leader && z + 4 <= z1 + 1) {
      mbar_expect_tx(&full[S], T::BYTES);
      tma_load_plane(tile[S], &tmU, &full[S], cx, cy, z + 4 + 1);
    }
    // upper z-face = lower z-face of cell z+1: uf = (u_z[J] + u_z[J-δi])/2, J = I+δz
    mbar_wait(&full[SP2], ((unsigned int)(z + 2 - zb) >> 2) & 1u);
#pragma unroll
    for (int c = 0; c < 3; ++c) R[c][SP2] = NN[c * PP];
    {
      const float uz = R[2][SP1];
      const float ufx = (uz + N[2 * PP - 1]) * 0.5f, ufy = (uz + N[2 * PP - PX]) * 0.5f, ufz = (uz + R[2][S]) * 0.5f;
      Fz[(S + 1) & 1][0] = flux1(ufx, R[0][SM1], R[0][S], R[0][SP1], R[0][SP2], nu);
      Fz[(S + 1) & 1][1] = flux1(ufy, R[1][SM1], R[1][S], R[1][SP1], R[1][SP2], nu);
      Fz[(S + 1) & 1][2] = flux1(ufz, R[2][SM1], R[2][S], R[2][SP1], R[2][SP2], nu);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      fx_hi[i] = __shfl_down_sync(0xffffffffu, fx_lo[i], 1);
      if (lane == 31 && wx < TX / 32 - 1) fx_hi[i] = Fx[buf][i][tiy][wx + 1];
      fy_hi[i] = tiy < TY - 1 ? Fy[buf][i][tiy + 1][tix] : 0.f;
      if (extra_x) fx_hi[i] = FxE[buf][i][tiy];
      if (extra_y) fy_hi[i] = ey[i];
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float v = 0.f;
        v += fx_lo[i]; v -= fx_hi[i];
        v += fy_lo[i]; v -= fy_hi[i];
        v += Fz[S & 1][i]; v -= Fz[(S + 1) & 1][i];
        if (HASG) v = (float)((double)v + a.gacc[i]);
        if (a.fuse_bdim1) {
          const float u0v = CORR ? *reinterpret_cast<const float*>(ub[i] + off) : R[i][S];
          v = u0v + a.dt * v - *reinterpret_cast<const float*>(vb[i] + off);
        }
        *reinterpret_cast<float*>(const_cast<char*>(fb[i]) + off) = v;
      }
    }
    buf ^= 1;
    off += doff;
  };
  // plane z0 sits in stage 2
  for (int z = z0;;) {
    step(std::integral_constant<int, 2>{}, z); if (++z >= z1) break;
    step(std::integral_constant<int, 3>{}, z); if (++z >= z1) break;
    step(std::integral_constant<int, 0>{}, z); if (++z >= z1) break;
    step(std::integral_constant<int, 1>{}, z); if (++z >= z1) break;
  }
}

// ---- two cells per thread ------------------------------------------------------------------------
// Same staging, but a lane owns the x-adjacent cells (x, x+1), x odd, so that the pair is 8-byte aligned in
// the staged planes and in global memory: neighbour reads become LDS.64 (21 per two cells instead of 62),
// the x-flux between the two cells never leaves the thread, a warp is one 64-cell row (the x-exchange is a
// single shuffle), and the loop/addressing/control overhead is shared by two cells.  The core therefore
// starts at x = 3; the column x = 2 goes to the boundary-shell kernel.
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
// 8-byte asynchronous global → shared copy (LDGSTS): no destination registers while the miss is outstanding
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int TY, int MINB, bool HASG, bool CORR>
__global__ void __launch_bounds__(32 * TY, MINB)
    k_conv_diff3_tma2(const __grid_constant__ CUtensorMap tmU, CDArgs a, int zchunk, int ntx, int nty) {
  constexpr int TX = 64;
  using T = Tile<TX, TY>;
  constexpr int PX = T::PX, PP = T::PP, STAGE = T::STAGE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const tile0 = reinterpret_cast<float*>(smem_raw);                                  // [4][STAGE]
  float* const Fy = tile0 + 4 * STAGE;                                                      // [2][3][TY][TX]
  float* const FxE = Fy + 2 * 3 * TY * TX;                                                  // [2][3][TY]
  float* const UxE = FxE + 2 * 3 * TY;                                                      // [2][TY]
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(UxE + 2 * TY);    // [4]
  float* const VU = reinterpret_cast<float*>(full + 4);                                     // [2][6][TY][TX]: V, u⁰ of the own cells
  const GridD& g = a.g;
  const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
  const int64_t s1 = g.s[1], s2 = g.s[2];
  const int lane = threadIdx.x, tiy = threadIdx.y;
  const bool leader = (lane | tiy) == 0;
  const int bt = blockIdx.x;
  const int tx = bt % ntx, ty = (bt / ntx) % nty, tz = bt / (ntx * nty);
  const int xb = tx * TX, yb = ty * TY;
  const int x = 3 + xb + 2 * lane, y = 2 + yb + tiy;  // cell A; cell B = x+1
  const int zlo = g.zlo_halo ? 1 : 2, zhi = g.zhi_halo ? nz - 2 : nz - 3;
  const int z0 = zlo + tz * zchunk, z1 = min(z0 + zchunk, zhi + 1);
  const int xlast = (nx - 3) - (((nx - 3) & 1) ^ 1);  // last odd cell index of the core
  const int xc = min(x, xlast), yc = min(y, ny - 3);   // out-of-range threads shadow the last pair / row
  const bool row_ok = y <= ny - 3;
  const bool valid_a = row_ok && x <= xlast, valid_b = row_ok && x + 1 <= nx - 3;
  const bool extra_y = (yc == ny - 3) || (tiy == TY - 1);
  const float nu = a.nu;
  const int toff = (yc - yb) * PX + (xc - xb) + WL_OX;  // cell A inside a staged plane (even: float2-aligned)
  const int zb = z0 - 2;
  const bool edge_warp = tiy == 0;
  const int erow = min(lane / 3, TY - 1), ecmp = lane - 3 * (lane / 3);
  const int xe = min(3 + xb + TX, nx - 2);  // cell whose lower x-face lies beyond the tile's last valid cell
  const int eoff = (min(2 + yb + erow, ny - 3) - yb) * PX + (xe - xb) + WL_OX;
  const bool a_edge = xc == xe - 1, b_edge = xc + 1 == xe - 1;
  const int cx = xb, cy = yb;
  if (leader) {
#pragma unroll
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (leader) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      mbar_expect_tx(&full[s], T::BYTES);
      tma_load_plane(tile0 + s * STAGE, &tmU, &full[s], cx, cy, zb + s + 1);
    }
  }
  int64_t off = (xc + yc * s1 + (int64_t)z0 * s2) * (int64_t)sizeof(float);
  const int64_t doff = s2 * (int64_t)sizeof(float);
  const char* const fb[3] = {(const char*)a.f, (const char*)(a.f + g.comp), (const char*)(a.f + 2 * g.comp)};
  const char* const vb[3] = {(const char*)a.V, (const char*)(a.V + g.comp), (const char*)(a.V + 2 * g.comp)};
  const char* const ub[3] = {(const char*)a.u0, (const char*)(a.u0 + g.comp), (const char*)(a.u0 + 2 * g.comp)};
  float2 R[3][4];   // own two columns, plane P in slot (P - zb) & 3
  float2 Fz[2][3];  // carried z-fluxes, double-buffered by step parity
  float2 cyp;       // u_y(y+1) of the previous plane
  const float* const tb = tile0 + toff;
#pragma unroll
  for (int s = 0; s < 4; ++s) mbar_wait(&full[s], 0);
  {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 4; ++k) R[c][k] = ld2(tb + k * STAGE + c * PP);
    }
    const float* C = tb + 1 * STAGE;
    const float* N = tb + 2 * STAGE;
    if (edge_warp && lane < 3 * TY && ecmp == 2) UxE[erow] = tile0[1 * STAGE + eoff];
    cyp = ld2(C + PP + PX);
    const float2 uz = R[2][2];
    const float2 nzy = ld2(N + 2 * PP - PX);
    const float nzx = N[2 * PP - 1];
    Fz[0][0] = flux2(half_sum2(uz, make_float2(nzx, uz.x)), R[0][0], R[0][1], R[0][2], R[0][3], nu);
    Fz[0][1] = flux2(half_sum2(uz, nzy), R[1][0], R[1][1], R[1][2], R[1][3], nu);
    Fz[0][2] = flux2(half_sum2(uz, R[2][1]), R[2][0], R[2][1], R[2][2], R[2][3], nu);
  }
  __syncthreads();
  if (leader) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (zb + s + 4 <= z1 + 1) {
        mbar_expect_tx(&full[s], T::BYTES);
        tma_load_plane(tile0 + s * STAGE, &tmU, &full[s], cx, cy, zb + s + 4 + 1);
      }
    }
  }
  int buf = 0;
  // V and u⁰ are read once, at the very end of a step, with few warps per SM to hide the miss: each thread
  // copies the values of its own two cells asynchronously into shared memory one plane ahead
  // (cp.async, no registers held) and picks them up after cp.async.wait_group
  float* const vu = VU + tiy * TX + 2 * lane;
  // clean tile-planes (CDArgs::clean): V is zero there and is not fetched.  The flags of the block's z-chunk are
  // copied to shared memory up front: a global load per step sat on the critical path of the march (+0.3 ms per launch)
  __shared__ unsigned char cflag[128];
  if (a.ufused) {
    const int t = tiy * 32 + lane;
    if (t < 128) cflag[t] = (t < z1 - z0) ? a.clean[((size_t)(z0 + t) * nty + ty) * ntx + tx] : 0;
  }
  __syncthreads();
  auto is_clean = [&](int z) -> bool { return a.ufused != nullptr && cflag[z - z0] != 0; };
  auto fetch_plane = [&](int st, int64_t o, bool cl) {
    if (a.fuse_bdim1) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (!cl) cp_async8(vu + (st * 6 + i) * TY * TX, vb[i] + o);
        if (CORR) cp_async8(vu + (st * 6 + 3 + i) * TY * TX, ub[i] + o);
      }
    }
    cp_async_commit();
  };
  bool cl_cur = is_clean(z0);
  fetch_plane(0, off, cl_cur);
  const char* const qb[3] = {(const char*)a.ufused, (const char*)(a.ufused + g.comp), (const char*)(a.ufused + 2 * g.comp)};
  auto step = [&](auto stage_tag, int z) {
    constexpr int S = decltype(stage_tag)::value;
    constexpr int SM1 = (S + 3) & 3, SP1 = (S + 1) & 3, SP2 = (S + 2) & 3;
    const float* C = tb + S * STAGE;
    const float* N = tb + SP1 * STAGE;
    const float* NN = tb + SP2 * STAGE;
    const bool cl = cl_cur;
    cl_cur = z + 1 < z1 ? is_clean(z + 1) : false;
    fetch_plane(buf ^ 1, off + doff, cl_cur);
    float2 fx[3], fxh[3], fy[3], fyh[3];
    float2 ym1[3], yp1[3];
    float2 xm[3];
    {
      float2 xq[3], ym2[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        xm[c] = ld2(C + c * PP - 2);
        xq[c] = ld2(C + c * PP + 2);
        ym2[c] = ld2(C + c * PP - 2 * PX);
        ym1[c] = ld2(C + c * PP - PX);
        yp1[c] = ld2(C + c * PP + PX);
      }
      const float2 ux = R[0][S], uy = R[1][S];
      // lower x-faces of A and B (cell pair = one packed evaluation): stencil (x-2, x-1, x, x+1) for A, shifted by one for B
      fx[0] = flux2(half_sum2(ux, make_float2(xm[0].y, ux.x)), xm[0], make_float2(xm[0].y, R[0][S].x), R[0][S], make_float2(R[0][S].y, xq[0].x), nu);
      fx[1] = flux2(half_sum2(ux, ym1[0]), xm[1], make_float2(xm[1].y, R[1][S].x), R[1][S], make_float2(R[1][S].y, xq[1].x), nu);
      fx[2] = flux2(half_sum2(ux, R[0][SM1]), xm[2], make_float2(xm[2].y, R[2][S].x), R[2][S], make_float2(R[2][S].y, xq[2].x), nu);
      // lower y-faces
      fy[0] = flux2(half_sum2(uy, make_float2(xm[1].y, uy.x)), ym2[0], ym1[0], R[0][S], yp1[0], nu);
      fy[1] = flux2(half_sum2(uy, ym1[1]), ym2[1], ym1[1], R[1][S], yp1[1], nu);
      fy[2] = flux2(half_sum2(uy, R[1][SM1]), ym2[2], ym1[2], R[2][S], yp1[2], nu);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) *reinterpret_cast<float2*>(Fy + ((buf * 3 + i) * TY + tiy) * TX + 2 * lane) = fy[i];
    if (edge_warp) {  // the x-face beyond the tile: 3·TY fluxes in one evaluation
      const float* E = tile0 + S * STAGE + eoff;
      const float* Ei = E + ecmp * PP;
      const float uxe = E[0];
      const float other = ecmp == 0 ? E[-1] : (ecmp == 1 ? E[-PX] : UxE[buf * TY + erow]);
      const float fe = flux1((uxe + other) * 0.5f, Ei[-2], Ei[-1], Ei[0], Ei[1], nu);
      if (lane < 3 * TY) {
        FxE[(buf * 3 + ecmp) * TY + erow] = fe;
        if (ecmp == 2) UxE[(buf ^ 1) * TY + erow] = uxe;
      }
    }
    if (extra_y) {  // lower y-faces of the cells y+1 that no thread of the block owns
      const float uyxm = C[PP + PX - 1];  // u_y(x-1, y+1)
      const float2 uyp = yp1[1];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float2 yp2 = ld2(C + c * PP + 2 * PX);
        const float2 other = c == 0 ? make_float2(uyxm, uyp.x) : (c == 1 ? R[1][S] : cyp);
        fyh[c] = flux2(half_sum2(uyp, other), ym1[c], R[c][S], yp1[c], yp2, nu);
      }
    }
    cyp = yp1[1];
    __syncthreads();
    if (leader && z + 4 <= z1 + 1) {
      mbar_expect_tx(&full[S], T::BYTES);
      tma_load_plane(tile0 + S * STAGE, &tmU, &full[S], cx, cy, z + 4 + 1);
    }
    mbar_wait(&full[SP2], ((unsigned int)(z + 2 - zb) >> 2) & 1u);
#pragma unroll
    for (int c = 0; c < 3; ++c) R[c][SP2] = ld2(NN + c * PP);
    {
      const float2 uz = R[2][SP1];
      const float2 nzy = ld2(N + 2 * PP - PX);
      const float nzx = N[2 * PP - 1];
      constexpr int Q = (S + 1) & 1;
      Fz[Q][0] = flux2(half_sum2(uz, make_float2(nzx, uz.x)), R[0][SM1], R[0][S], R[0][SP1], R[0][SP2], nu);
      Fz[Q][1] = flux2(half_sum2(uz, nzy), R[1][SM1], R[1][S], R[1][SP1], R[1][SP2], nu);
      Fz[Q][2] = flux2(half_sum2(uz, R[2][S]), R[2][SM1], R[2][S], R[2][SP1], R[2][SP2], nu);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      fxh[i].x = fx[i].y;                                          // B's lower face is A's upper face
      fxh[i].y = __shfl_down_sync(0xffffffffu, fx[i].x, 1);      // the next lane's A
      const float fe = FxE[(buf * 3 + i) * TY + tiy];
      if (a_edge) fxh[i].x = fe;
      if (b_edge) fxh[i].y = fe;
      if (!extra_y) fyh[i] = *reinterpret_cast<const float2*>(Fy + ((buf * 3 + i) * TY + tiy + 1) * TX + 2 * lane);
    }
    cp_async_wait1();  // this plane's V/u⁰ (the group committed one step ago) have landed
    if (valid_a) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float2 v = add2(splat2(0.f), fx[i]);
        v = sub2(v, fxh[i]); v = add2(v, fy[i]); v = sub2(v, fyh[i]); v = add2(v, Fz[S & 1][i]); v = sub2(v, Fz[(S + 1) & 1][i]);
        if (HASG) { v.x = (float)((double)v.x + a.gacc[i]); v.y = (float)((double)v.y + a.gacc[i]); }
        if (a.fuse_bdim1) {
          const float2 u0v = CORR ? ld2(vu + (buf * 6 + 3 + i) * TY * TX) : R[i][S];
          const float2 Vv = cl ? make_float2(0.f, 0.f) : ld2(vu + (buf * 6 + i) * TY * TX);
          v = sub2(add2(u0v, mul2_sep(v, a.dt)), Vv);   // u⁰ + Δt·f - V: the product is rounded on its own
        }
        float* dst = reinterpret_cast<float*>(const_cast<char*>(fb[i]) + off);
        if (cl) {  // BDIM! stage 2 away from bodies: u = (u_acc + f)·scale with u_acc = 0 (predictor) or u (corrector)
          // the reference adds in Float64 and rounds once; for the sum of two Float32 that is the Float32 sum (53 ≥ 2·24+2 bits:
          // the double rounding is innocuous), so the Float32 add below has the same bits — and no FP64 conversions to issue
          v = mul2(add2(CORR ? R[i][S] : splat2(0.f), v), splat2(a.bscale));
          dst = reinterpret_cast<float*>(const_cast<char*>(qb[i]) + off);
        }
        if (valid_b) *reinterpret_cast<float2*>(dst) = v;
        else dst[0] = v.x;
      }
    }
    buf ^= 1;
    off += doff;
  };
  for (int z = z0;;) {
    step(std::integral_constant<int, 2>{}, z); if (++z >= z1) break;
    step(std::integral_constant<int, 3>{}, z); if (++z >= z1) break;
    step(std::integral_constant<int, 0>{}, z); if (++z >= z1) break;
    step(std::integral_constant<int, 1>{}, z); if (++z >= z1) break;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kTX = 64, kTY = 4, kMINB = 4;
constexpr int kTY2 = 8, kMINB2 = 2;  // two-cells-per-thread variant: 32 x 8 threads, 64 x 8 cells
constexpr size_t kSmem2 = sizeof(float) * (4 * Tile<64, kTY2>::STAGE + 2 * 3 * kTY2 * 64 + 2 * 3 * kTY2 + 2 * kTY2 + 2 * 6 * kTY2 * 64) + 4 * sizeof(unsigned long long);

}  // namespace

struct ConvDiffTma {
  CUtensorMap map[4];         // the velocity-sized buffers u, u⁰, f, w (mom_step rotates the pointers, not the data)
  const float* ptr[4];
  int cells_per_thread;       // 1: k_conv_diff3_tma (core from x = 2), 2: k_conv_diff3_tma2 (core from x = 3)
};

// Builds the tensor maps of the two velocity buffers.  Leaves s.cd_tma null (the first marching
// kernel stays in use) if the driver entry point is missing or the grid is too small.
int convdiff_tma_setup(Sim& s) {
  s.cd_tma = nullptr;
  const char* env = getenv("WL_CD_TMA");
  if (env && atoi(env) == 0) return 0;
  const GridD& g = s.g;
  if (s.D != 3 || g.n[0] < 8 || g.n[1] < 8 || g.n[2] < 8) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return 0;
  }
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  ConvDiffTma* t = new ConvDiffTma;
  using T = Tile<kTX, kTY>;
  const char* e2 = getenv("WL_CD_CELLS");
  t->cells_per_thread = e2 ? atoi(e2) : 2;
  if (t->cells_per_thread != 1) t->cells_per_thread = 2;
  const cuuint32_t box_y = t->cells_per_thread == 2 ? (cuuint32_t)Tile<64, kTY2>::PY : (cuuint32_t)T::PY;
  auto make = [&](CUtensorMap* m, const float* field) -> bool {
    void* base = (void*)(field - g.lead);  // start of the allocation: 256-byte aligned
    const cuuint64_t dims[4] = {(cuuint64_t)g.s[1], (cuuint64_t)g.n[1], (cuuint64_t)g.n[2] + 2, 3};
    const cuuint64_t strides[3] = {(cuuint64_t)g.s[1] * 4, (cuuint64_t)g.s[2] * 4, (cuuint64_t)g.comp * 4};
    const cuuint32_t box[4] = {(cuuint32_t)T::PX, box_y, 1, 3};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  t->ptr[0] = s.u; t->ptr[1] = s.u0; t->ptr[2] = s.f; t->ptr[3] = s.w;
  if (!make(&t->map[0], s.u) || !make(&t->map[1], s.u0) || !make(&t->map[2], s.f) || !make(&t->map[3], s.w)) {
    if (getenv("WL_VERBOSE")) fprintf(stderr, "[waterlily_b200] cuTensorMapEncodeTiled failed: conv_diff! uses the untiled kernel\n");
    delete t;
    return 0;
  }
  if (t->cells_per_thread == 2) {
    cudaError_t e = cudaSuccess;
    auto optin = [&](const void* k) { if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2); };
    optin((const void*)k_conv_diff3_tma2<kTY2, kMINB2, false, false>);
    optin((const void*)k_conv_diff3_tma2<kTY2, kMINB2, false, true>);
    optin((const void*)k_conv_diff3_tma2<kTY2, kMINB2, true, false>);
    optin((const void*)k_conv_diff3_tma2<kTY2, kMINB2, true, true>);
    if (e != cudaSuccess) { cudaGetLastError(); delete t; return 0; }
  }
  if (getenv("WL_VERBOSE")) fprintf(stderr, "[waterlily_b200] conv_diff!: TMA-staged kernel, %d cell(s) per thread\n", t->cells_per_thread);
  s.cd_tma = t;
  return 0;
}
void convdiff_tma_destroy(Sim& s) {
  delete (ConvDiffTma*)s.cd_tma;
  s.cd_tma = nullptr;
}

template <bool HASG, bool CORR>
static void launch_tma2(Sim& s, const CUtensorMap& m, const CDArgs& a, int zchunk) {
  const GridD& g = s.g;
  const int nzc = g.n[2] - 4 + (g.zlo_halo ? 1 : 0) + (g.zhi_halo ? 1 : 0);
  const int ntx = (g.n[0] - 5 + 63) / 64, nty = (g.n[1] - 4 + kTY2 - 1) / kTY2, ntz = (nzc + zchunk - 1) / zchunk;
  k_conv_diff3_tma2<kTY2, kMINB2, HASG, CORR><<<ntx * nty * ntz, dim3(32, kTY2, 1), kSmem2, s.stream>>>(m, a, zchunk, ntx, nty);
}
template <bool HASG, bool CORR>
static void launch_tma(Sim& s, const CUtensorMap& m, const CDArgs& a, int zchunk) {
  const GridD& g = s.g;
  const int nzc = g.n[2] - 4 + (g.zlo_halo ? 1 : 0) + (g.zhi_halo ? 1 : 0);
  const int ntx = (g.n[0] - 4 + kTX - 1) / kTX, nty = (g.n[1] - 4 + kTY - 1) / kTY, ntz = (nzc + zchunk - 1) / zchunk;
  k_conv_diff3_tma<kTX, kTY, kMINB, HASG, CORR><<<ntx * nty * ntz, dim3(kTX, kTY, 1), 0, s.stream>>>(m, a, zchunk, ntx, nty);
}
// the 3-D core [2,n-3]³ (slab faces included); false if this source buffer has no tensor map
bool convdiff_tma_launch(Sim& s, const CDArgs& a, int zchunk, bool corr) {
  ConvDiffTma* t = (ConvDiffTma*)s.cd_tma;
  if (!t) return false;
  const CUtensorMap* m = nullptr;
  for (int q = 0; q < 4; ++q) if (a.u == t->ptr[q]) m = &t->map[q];
  if (!m) return false;
  if (t->cells_per_thread == 2) {
    if (a.has_g) {
      if (corr) launch_tma2<true, true>(s, *m, a, zchunk);
      else launch_tma2<true, false>(s, *m, a, zchunk);
    } else {
      if (corr) launch_tma2<false, true>(s, *m, a, zchunk);
      else launch_tma2<false, false>(s, *m, a, zchunk);
    }
  } else if (a.has_g) {
    if (corr) launch_tma<true, true>(s, *m, a, zchunk);
    else launch_tma<true, false>(s, *m, a, zchunk);
  } else {
    if (corr) launch_tma<false, true>(s, *m, a, zchunk);
    else launch_tma<false, false>(s, *m, a, zchunk);
  }
  ++s.launches;
  return true;
}
// tile grid of the two-cells-per-thread kernel (the `clean` map of CDArgs is indexed with it); false if another kernel runs
bool convdiff_tma_tiles(const Sim& s, int* ntx, int* nty) {
  const ConvDiffTma* t = (const ConvDiffTma*)s.cd_tma;
  if (!t || t->cells_per_thread != 2) return false;
  *ntx = (s.g.n[0] - 5 + 63) / 64;
  *nty = (s.g.n[1] - 4 + kTY2 - 1) / kTY2;
  return true;
}
// first x index of the core the TMA kernel covers (the boundary-shell kernel takes everything below it)
int convdiff_tma_core_x0(const Sim& s) {
  const ConvDiffTma* t = (const ConvDiffTma*)s.cd_tma;
  return (t && t->cells_per_thread == 2) ? 3 : 2;
}
