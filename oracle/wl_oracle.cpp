// wl_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A literal, un-fused C++/OpenMP restatement of the reference's `mom_step!` hot path
// (marinlauber/WaterLily v1.3.0: src/Flow.jl, src/Poisson.jl, src/MultiLevelPoisson.jl and the
// BC functions of src/util.jl).  One loop here per `@loop` there, one reduction per `⋅`/`sum`/
// `maximum`, the same aliasing (σ ≡ Φ ≡ levels[1].z, μ₀ ≡ levels[1].L, p ≡ levels[1].x), padded
// full-array reductions, and Julia's Float32/Float64 promotion rules, so that the reference's
// ghost-cell quirks (SURVEY.md §8a Q0–Q2) come out by construction.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (libwaterlily_b200.so, waterlily_b200/) never does.
//
// PARITY STATUS: the reference is Julia and cannot run in this environment (no julia binary,
// no network).  The oracle is pinned against every known-answer assertion the reference's own
// test-suite holds for this path (test/maintests.jl, see tests/test_oracle_pins.py); the
// reference stores no golden arrays.  Summation order of `⋅`/`sum` is library-defined in the
// reference (OpenBLAS / Base pairwise) and not pinned by any reference test: bitwise parity is
// unpinned, tolerance parity is pinned.
//
// It exports the same C ABI as include/waterlily_b200.h with the prefix `wlo_`.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/waterlily_b200.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& m) { g_err = m; return code; }

constexpr float EPS32 = 1.1920929e-07f;  // eps(Float32)

// ------------------------------------------------------------------------------------------
// grid + index helpers (src/util.jl:26-61,180-182).  1-based indices like the reference.
// ------------------------------------------------------------------------------------------
struct Grid {
  int D = 0;
  int n[3] = {1, 1, 1};       // padded size Ng per dim (1 for dims >= D)
  int64_t s[3] = {1, 1, 1};   // strides
  int64_t cells = 1;
  void set(int D_, const int* Ng) {
    D = D_; cells = 1;
    for (int k = 0; k < 3; ++k) { n[k] = k < D ? Ng[k] : 1; s[k] = cells; cells *= n[k]; }
  }
  inline int64_t at(int i, int j, int k) const { return (i - 1) + (j - 1) * s[1] + (k - 1) * s[2]; }
  int64_t interior() const { int64_t c = 1; for (int k = 0; k < D; ++k) c *= n[k] - 2; return c; }
};

struct Range { int lo[3], hi[3]; };

// inside(a) (util.jl:47)
Range r_inside(const Grid& g) {
  Range R; for (int k = 0; k < 3; ++k) { R.lo[k] = k < g.D ? 2 : 1; R.hi[k] = k < g.D ? g.n[k] - 1 : 1; } return R;
}
// inside_u(dims,j) (util.jl:55-57): 3:Ng-1 in dim j, 2:Ng elsewhere
Range r_inside_u(const Grid& g, int j) {
  Range R; for (int k = 0; k < 3; ++k) {
    if (k >= g.D) { R.lo[k] = R.hi[k] = 1; }
    else if (k == j) { R.lo[k] = 3; R.hi[k] = g.n[k] - 1; }
    else { R.lo[k] = 2; R.hi[k] = g.n[k]; }
  } return R;
}
// slice(dims,i,j,low) (util.jl:180-182)
Range r_slice(const Grid& g, int i, int j, int low = 1) {
  Range R; for (int k = 0; k < 3; ++k) {
    if (k >= g.D) { R.lo[k] = R.hi[k] = 1; }
    else if (k == j) { R.lo[k] = R.hi[k] = i; }
    else { R.lo[k] = low; R.hi[k] = g.n[k]; }
  } return R;
}

// @loop <body> over I ∈ R (util.jl:119-141): one parallel loop per reference loop.
template <class F>
inline void loop(const Grid& g, const Range& R, F&& f) {
  const int ni = R.hi[0] - R.lo[0] + 1, nj = R.hi[1] - R.lo[1] + 1, nk = R.hi[2] - R.lo[2] + 1;
  if (ni <= 0 || nj <= 0 || nk <= 0) return;
  const int64_t total = (int64_t)ni * nj * nk;
#pragma omp parallel for collapse(2) schedule(static) if (total > 16384)
  for (int k = R.lo[2]; k <= R.hi[2]; ++k)
    for (int j = R.lo[1]; j <= R.hi[1]; ++j) {
      int64_t I = g.at(R.lo[0], j, k);
      for (int i = R.lo[0]; i <= R.hi[0]; ++i, ++I) f(I, i, j, k);
    }
}

// ------------------------------------------------------------------------------------------
// reductions (`⋅` → BLAS sdot, `sum` → Base pairwise sum in the reference).
//
// The reference's summation ORDER is library-defined (OpenBLAS kernel / Base.mapreduce blocking)
// and unpinned, and it matters: z⋅ϵ = ϵᵀAϵ in pcg! (Poisson.jl:131) cancels heavily, so two legal
// Float32 orders give α's that differ by 1e-6…1e-5.  The oracle therefore defines every reduction
// canonically as the correctly rounded Float32 of the exact sum (products and accumulation in
// Float64), which is order-independent; the CUDA path does the same.  Mode 1 (wlo_set_sum_mode)
// switches to a BLAS-like Float32 accumulation (8 interleaved lanes per 4096-element chunk, chunks
// combined pairwise) and is used by tests/ to measure that reproducibility floor of the reference.
// Both modes use a fixed chunking, so results do not depend on the OpenMP thread count.
// ------------------------------------------------------------------------------------------
constexpr int64_t CHUNK = 4096;
int g_sum_mode = 0;  // 0: canonical (exact, rounded once)   1: Float32 BLAS-like accumulation
template <class F>
float chunked_sum_f32(int64_t n, F&& term) {
  const int64_t nch = (n + CHUNK - 1) / CHUNK;
  std::vector<float> part(nch);
#pragma omp parallel for schedule(static) if (n > 65536)
  for (int64_t c = 0; c < nch; ++c) {
    const int64_t a = c * CHUNK, b = std::min(n, a + CHUNK);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = a;
    for (; i + 8 <= b; i += 8)
      for (int l = 0; l < 8; ++l) acc[l] += (float)term(i + l);
    float t = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < b; ++i) t += (float)term(i);
    part[c] = t;
  }
  int64_t m = nch;
  while (m > 1) {  // pairwise combine of chunk partials
    const int64_t h = (m + 1) / 2;
    for (int64_t c = 0; c + h < m; ++c) part[c] += part[c + h];
    m = h;
  }
  return nch ? part[0] : 0.f;
}
template <class F>
float chunked_sum_f64(int64_t n, F&& term) {
  const int64_t nch = (n + CHUNK - 1) / CHUNK;
  std::vector<double> part(nch);
#pragma omp parallel for schedule(static) if (n > 65536)
  for (int64_t c = 0; c < nch; ++c) {
    const int64_t a = c * CHUNK, b = std::min(n, a + CHUNK);
    double t = 0.0;
    for (int64_t i = a; i < b; ++i) t += term(i);
    part[c] = t;
  }
  double t = 0.0;
  for (int64_t c = 0; c < nch; ++c) t += part[c];
  return (float)t;
}
float dot(const float* a, const float* b, int64_t n) {
  if (g_sum_mode == 1) return chunked_sum_f32(n, [=](int64_t i) { return a[i] * b[i]; });
  return chunked_sum_f64(n, [=](int64_t i) { return (double)a[i] * (double)b[i]; });
}
float sum(const float* a, int64_t n) {
  if (g_sum_mode == 1) return chunked_sum_f32(n, [=](int64_t i) { return a[i]; });
  return chunked_sum_f64(n, [=](int64_t i) { return (double)a[i]; });
}
float maximum(const float* a, int64_t n) {
  float m = a[0];
#pragma omp parallel for reduction(max : m) schedule(static) if (n > 65536)
  for (int64_t i = 0; i < n; ++i) m = std::max(m, a[i]);
  return m;
}

// ------------------------------------------------------------------------------------------
// BC functions (src/util.jl:192-231)
// ------------------------------------------------------------------------------------------
// BC!(a,A,saveexit,perdir) util.jl:192-210
void BC(const Grid& g, float* a, const double* A, bool saveexit, const bool* per) {
  const int n = g.D;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
    float* ai = a + i * g.cells; const int64_t sj = g.s[j]; const int Nj = g.n[j];
    if (per[j]) {
      loop(g, r_slice(g, 1, j), [&](int64_t I, int, int, int) { ai[I] = ai[I + (int64_t)(Nj - 2) * sj]; });   // a[CIj(j,I,N[j]-1)]
      loop(g, r_slice(g, Nj, j), [&](int64_t I, int, int, int) { ai[I] = ai[I - (int64_t)(Nj - 2) * sj]; });  // a[CIj(j,I,2)]
    } else if (i == j) {
      const float Ai = (float)A[i];
      for (int sidx = 1; sidx <= 2; ++sidx) loop(g, r_slice(g, sidx, j), [&](int64_t I, int, int, int) { ai[I] = Ai; });
      if (!saveexit || i > 0) loop(g, r_slice(g, Nj, j), [&](int64_t I, int, int, int) { ai[I] = Ai; });
    } else {
      loop(g, r_slice(g, 1, j), [&](int64_t I, int, int, int) { ai[I] = ai[I + sj]; });
      loop(g, r_slice(g, Nj, j), [&](int64_t I, int, int, int) { ai[I] = ai[I - sj]; });
    }
  }
}
// perBC!(a,perdir) util.jl:227-231
void perBC(const Grid& g, float* a, const bool* per) {
  for (int j = 0; j < g.D; ++j) if (per[j]) {
    const int64_t sj = g.s[j]; const int Nj = g.n[j];
    loop(g, r_slice(g, 1, j), [&](int64_t I, int, int, int) { a[I] = a[I + (int64_t)(Nj - 2) * sj]; });
    loop(g, r_slice(g, Nj, j), [&](int64_t I, int, int, int) { a[I] = a[I - (int64_t)(Nj - 2) * sj]; });
  }
}
// exitBC!(u,u⁰,U,Δt) util.jl:216-222.  U arrives as Float64 (a Float64 tuple in the reference
// promotes the update to Float64; an Int tuple with U[1]∈{0,1} gives the same values).
void exitBC(const Grid& g, float* u, const float* u0, const double* U, float dt) {
  // exitR = slice(N.-1, N[1], 1, 2): I_1 = Ng, others 2:Ng-1
  Range R; for (int k = 0; k < 3; ++k) { R.lo[k] = k < g.D ? 2 : 1; R.hi[k] = k < g.D ? g.n[k] - 1 : 1; }
  R.lo[0] = R.hi[0] = g.n[0];
  const double Udt = U[0] * (double)dt;
  loop(g, R, [&](int64_t I, int, int, int) { u[I] = (float)((double)u0[I] - Udt * (double)(u0[I] - u0[I - 1])); });
  // ∮u = sum(u[exitR,1])/length(exitR) - U[1]   (the slice is copied to a dense array first)
  std::vector<float> plane; plane.reserve((size_t)(g.cells / g.n[0]));
  for (int k = R.lo[2]; k <= R.hi[2]; ++k) for (int j = R.lo[1]; j <= R.hi[1]; ++j) plane.push_back(u[g.at(g.n[0], j, k)]);
  const float mean = sum(plane.data(), (int64_t)plane.size()) / (float)plane.size();
  const double du = (double)mean - U[0];
  loop(g, R, [&](int64_t I, int, int, int) { u[I] = (float)((double)u[I] - du); });
}

// ------------------------------------------------------------------------------------------
// Flow.jl kernels
// ------------------------------------------------------------------------------------------
// median (Flow.jl:25-34)
inline float median(float a, float b, float c) {
  if (a > b) { if (b >= c) return b; if (a > c) return c; }
  else { if (b <= c) return b; if (a < c) return c; }
  return a;
}
// quick(u,c,d) (Flow.jl:4): all-Float32
inline float quick(float u, float c, float d) { return median((5 * c + 2 * d - u) / 6, c, median(10 * c - 9 * u, c, d)); }
// ϕ(a,I,f) (Flow.jl:3): Float32 add, then ×0.5 promotes to Float64
inline double phi(const float* f, int64_t I, int64_t sa) { return (double)(f[I] + f[I - sa]) * 0.5; }
// ϕu, ϕuP, ϕuL, ϕuR (Flow.jl:6-9) with λ = quick; `u` is the Float64 face velocity ϕ(i,CI(I,j),u), sa the stride of direction a,
// Ip the linear index of the upwind-upwind sample of the periodic variant
inline double phiu(const float* f, int64_t I, int64_t sa, double u) { return u > 0 ? u * quick(f[I - 2 * sa], f[I - sa], f[I]) : u * quick(f[I + sa], f[I], f[I - sa]); }
inline double phiuP(const float* f, int64_t Ip, int64_t I, int64_t sa, double u) { return u > 0 ? u * quick(f[Ip], f[I - sa], f[I]) : u * quick(f[I + sa], f[I], f[I - sa]); }
inline double phiuL(const float* f, int64_t I, int64_t sa, double u) { return u > 0 ? u * phi(f, I, sa) : u * quick(f[I + sa], f[I], f[I - sa]); }
inline double phiuR(const float* f, int64_t I, int64_t sa, double u) { return u < 0 ? u * phi(f, I, sa) : u * quick(f[I - 2 * sa], f[I - sa], f[I]); }

// conv_diff!(r,u,Φ;ν,perdir) Flow.jl:36-60
void conv_diff(const Grid& g, float* r, const float* u, float* Phi, float nu, const bool* per) {
  const int n = g.D;
  std::memset(r, 0, sizeof(float) * g.cells * n);  // r .= 0.
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
    float* ri = r + i * g.cells; const float* ui = u + i * g.cells; const float* uj = u + j * g.cells;
    const int64_t si = g.s[i], sj = g.s[j]; const int Nj = g.n[j];
    // lowerBoundary! (Flow.jl:54,58-59) over slice(N,2,j,2)
    if (per[j]) {
      const int64_t wrap = (int64_t)(Nj - 2 - 2) * sj;  // CIj(j,I,N[j]-2) from I_j=2
      loop(g, r_slice(g, 2, j, 2), [&](int64_t I, int, int, int) {
        const double uf = phi(uj, I, si);
        const double flux = phiuP(ui, I + wrap, I, sj, uf);
        Phi[I] = (float)(flux - (double)(nu * (ui[I] - ui[I - sj])));
        ri[I] += Phi[I];
      });
    } else {
      loop(g, r_slice(g, 2, j, 2), [&](int64_t I, int, int, int) {
        const double uf = phi(uj, I, si);
        const double flux = phiuL(ui, I, sj, uf);
        ri[I] = (float)((double)ri[I] + (flux - (double)(nu * (ui[I] - ui[I - sj]))));
      });
    }
    // inner cells (Flow.jl:45-47)
    loop(g, r_inside_u(g, j), [&](int64_t I, int, int, int) {
      const double uf = phi(uj, I, si);
      const double flux = phiu(ui, I, sj, uf);
      Phi[I] = (float)(flux - (double)(nu * (ui[I] - ui[I - sj])));
      ri[I] += Phi[I];
    });
    loop(g, r_inside_u(g, j), [&](int64_t I, int, int, int) { ri[I - sj] -= Phi[I]; });
    // upperBoundary! (Flow.jl:55,60) over slice(N,N[j],j,2)
    if (per[j]) {
      const int64_t to2 = (int64_t)(2 - Nj) * sj;  // CIj(j,I,2)
      loop(g, r_slice(g, Nj, j, 2), [&](int64_t I, int, int, int) { ri[I - sj] -= Phi[I + to2]; });
    } else {
      loop(g, r_slice(g, Nj, j, 2), [&](int64_t I, int, int, int) {
        const double uf = phi(uj, I, si);
        const double flux = phiuR(ui, I, sj, uf);
        ri[I - sj] = (float)((double)ri[I - sj] + (-flux + (double)(nu * (ui[I] - ui[I - sj]))));
      });
    }
  }
}

// ------------------------------------------------------------------------------------------
// Poisson (src/Poisson.jl)
// ------------------------------------------------------------------------------------------
struct Pois {
  Grid g;
  float *L = nullptr, *Dg = nullptr, *iD = nullptr, *x = nullptr, *eps = nullptr, *r = nullptr, *z = nullptr;
  bool own_xLz = false;
  bool per[3] = {false, false, false};
  std::vector<float> store;  // backing store of owned arrays
};

// diag (Poisson.jl:48-54) / set_diag! (:42-45)
void set_diag(Pois& p) {
  const Grid& g = p.g;
  loop(g, r_inside(g), [&](int64_t I, int, int, int) {
    float s = 0.f;
    for (int i = 0; i < g.D; ++i) s -= (p.L[I + i * g.cells] + p.L[I + g.s[i] + i * g.cells]);
    p.Dg[I] = s;
  });
  loop(g, r_inside(g), [&](int64_t I, int, int, int) {
    const float d = p.Dg[I];
    p.iD[I] = (d * d < 2 * EPS32) ? 0.f : 1.f / d;
  });
}
// mult (Poisson.jl:69-75)
inline float mult(const Pois& p, int64_t I, const float* x) {
  const Grid& g = p.g;
  float s = x[I] * p.Dg[I];
  for (int i = 0; i < g.D; ++i)
    s += (x[I - g.s[i]] * p.L[I + i * g.cells] + x[I + g.s[i]] * p.L[I + g.s[i] + i * g.cells]);
  return s;
}
// mult!(p,x) (Poisson.jl:62-68)
void mult_bang(Pois& p, float* x) {
  perBC(p.g, x, p.per);
  std::memset(p.z, 0, sizeof(float) * p.g.cells);
  loop(p.g, r_inside(p.g), [&](int64_t I, int, int, int) { p.z[I] = mult(p, I, x); });
}
// residual! (Poisson.jl:91-97)
void residual(Pois& p) {
  const Grid& g = p.g;
  perBC(g, p.x, p.per);
  loop(g, r_inside(g), [&](int64_t I, int, int, int) { p.r[I] = p.iD[I] == 0 ? 0.f : p.z[I] - mult(p, I, p.x); });
  const float s = sum(p.r, g.cells) / (float)g.interior();
  if (std::fabs(s) <= 2 * EPS32) return;
  loop(g, r_inside(g), [&](int64_t I, int, int, int) { p.r[I] = p.r[I] - s; });
}
// increment! (Poisson.jl:99-103)
void increment(Pois& p) {
  const Grid& g = p.g;
  perBC(g, p.eps, p.per);
  loop(g, r_inside(g), [&](int64_t I, int, int, int) {
    p.r[I] = p.r[I] - mult(p, I, p.eps);
    p.x[I] = p.x[I] + p.eps[I];
  });
}
// Jacobi! (Poisson.jl:110-113)
void jacobi(Pois& p, int it = 1) {
  for (int t = 0; t < it; ++t) {
    loop(p.g, r_inside(p.g), [&](int64_t I, int, int, int) { p.eps[I] = p.r[I] * p.iD[I]; });
    increment(p);
  }
}
// pcg! (Poisson.jl:123-144)
void pcg(Pois& p, int it = 6) {
  const Grid& g = p.g; const Range in = r_inside(g);
  float *x = p.x, *r = p.r, *e = p.eps, *z = p.z;
  loop(g, in, [&](int64_t I, int, int, int) { z[I] = e[I] = r[I] * p.iD[I]; });
  float rho = dot(r, z, g.cells);
  if (std::fabs(rho) < 10 * EPS32) return;
  for (int i = 1; i <= it; ++i) {
    perBC(g, e, p.per);
    loop(g, in, [&](int64_t I, int, int, int) { z[I] = mult(p, I, e); });
    const float alpha = rho / dot(z, e, g.cells);
    if (std::fabs((double)alpha) < 1e-2 || std::fabs((double)alpha) > 1e2) return;
    loop(g, in, [&](int64_t I, int, int, int) { x[I] += alpha * e[I]; r[I] -= alpha * z[I]; });
    if (i == it) return;
    loop(g, in, [&](int64_t I, int, int, int) { z[I] = r[I] * p.iD[I]; });
    const float rho2 = dot(r, z, g.cells);
    if (std::fabs(rho2) < 10 * EPS32) return;
    const float beta = rho2 / rho;
    loop(g, in, [&](int64_t I, int, int, int) { e[I] = beta * e[I] + z[I]; });
    rho = rho2;
  }
}
inline float L2(const Pois& p) { return dot(p.r, p.r, p.g.cells); }  // Poisson.jl:146

// ------------------------------------------------------------------------------------------
// MultiLevelPoisson (src/MultiLevelPoisson.jl)
// ------------------------------------------------------------------------------------------
// restrictL! (:26-32) with restrictL (:10-16), up(I,i) (:1)
void restrictL(const Grid& ga, float* a, const Grid& gb, const float* b, const bool* per) {
  const int n = ga.D;
  for (int i = 0; i < n; ++i) {
    float* ai = a + i * ga.cells; const float* bi = b + i * gb.cells;
    loop(ga, r_inside(ga), [&](int64_t I, int ii, int jj, int kk) {
      const int c[3] = {ii, jj, kk};
      int lo[3], hi[3];
      for (int k = 0; k < 3; ++k) {
        if (k < n) { lo[k] = 2 * c[k] - 2; hi[k] = 2 * c[k] - 1 - (k == i ? 1 : 0); } else { lo[k] = hi[k] = 1; }
      }
      float s = 0.f;
      for (int z = lo[2]; z <= hi[2]; ++z) for (int y = lo[1]; y <= hi[1]; ++y) for (int x = lo[0]; x <= hi[0]; ++x) s += bi[gb.at(x, y, z)];
      ai[I] = (float)(0.5 * (double)s);
    });
  }
  const double zero[3] = {0, 0, 0};
  BC(ga, a, zero, false, per);
}
// restrict! (:33) with restrict (:3-9)
void restrict_r(const Grid& ga, float* a, const Grid& gb, const float* b) {
  const int n = ga.D;
  loop(ga, r_inside(ga), [&](int64_t I, int ii, int jj, int kk) {
    const int c[3] = {ii, jj, kk};
    int lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { if (k < n) { lo[k] = 2 * c[k] - 2; hi[k] = 2 * c[k] - 1; } else { lo[k] = hi[k] = 1; } }
    float s = 0.f;
    for (int z = lo[2]; z <= hi[2]; ++z) for (int y = lo[1]; y <= hi[1]; ++y) for (int x = lo[0]; x <= hi[0]; ++x) s += b[gb.at(x, y, z)];
    a[I] = s;
  });
}
// prolongate! (:34) with down(I) (:2)
void prolongate(const Grid& ga, float* a, const Grid& gb, const float* b) {
  const int n = ga.D;
  loop(ga, r_inside(ga), [&](int64_t I, int ii, int jj, int kk) {
    const int c[3] = {ii, jj, kk}; int d[3] = {1, 1, 1};
    for (int k = 0; k < n; ++k) d[k] = (c[k] + 2) / 2;
    a[I] = b[gb.at(d[0], d[1], d[2])];
  });
}
inline bool divisible(int N) { return N % 2 == 0 && N > 4; }  // :36

struct Sim {
  int D = 0; Grid g;
  std::vector<float> u, u0, f, p, sigma, V, mu0, mu1;
  float nu = 0.f; std::vector<float> dt; bool exitBC = false; bool per[3] = {false, false, false};
  int pois_kind = WL_POIS_MULTILEVEL;
  std::vector<Pois> levels;
  std::vector<int> n_log;
  int64_t launches = 0;
  std::vector<double> nds;  // Metrics: nds(body, loc(0,I,Float64), t), (Ng...,D)
};

void make_level1(Sim& s) {
  Pois p; p.g = s.g; p.own_xLz = false;
  for (int k = 0; k < 3; ++k) p.per[k] = s.per[k];
  p.x = s.p.data(); p.L = s.mu0.data(); p.z = s.sigma.data();
  p.store.assign((size_t)s.g.cells * 4, 0.f);  // r, ϵ, D, iD  (Poisson.jl:33-34)
  s.levels.push_back(std::move(p));
  Pois& q = s.levels.back();
  q.r = q.store.data(); q.eps = q.r + s.g.cells; q.Dg = q.eps + s.g.cells; q.iD = q.Dg + s.g.cells;
  set_diag(q);
}
// restrictML (:18-25)
void push_coarse(Sim& s) {
  const Pois& b = s.levels.back();
  int Na[3]; for (int k = 0; k < s.D; ++k) Na[k] = 1 + b.g.n[k] / 2;
  Pois p; p.g.set(s.D, Na); p.own_xLz = true;
  for (int k = 0; k < 3; ++k) p.per[k] = b.per[k];
  const int64_t c = p.g.cells;
  p.store.assign((size_t)c * (s.D + 6), 0.f);
  s.levels.push_back(std::move(p));
  Pois& q = s.levels.back(); const Pois& fine = s.levels[s.levels.size() - 2];
  float* base = q.store.data();
  q.L = base; base += c * s.D; q.x = base; base += c; q.z = base; base += c;
  q.r = base; base += c; q.eps = base; base += c; q.Dg = base; base += c; q.iD = base;
  restrictL(q.g, q.L, fine.g, fine.L, fine.per);
  set_diag(q);
}
// update!(ml) (:62-68)
void update(Sim& s) {
  set_diag(s.levels[0]);
  for (size_t l = 1; l < s.levels.size(); ++l) {
    restrictL(s.levels[l].g, s.levels[l].L, s.levels[l - 1].g, s.levels[l - 1].L, s.levels[l - 1].per);
    set_diag(s.levels[l]);
  }
}
// Vcycle! (:70-82)
void vcycle(Sim& s, size_t l) {
  Pois& fine = s.levels[l]; Pois& coarse = s.levels[l + 1];
  jacobi(fine);
  restrict_r(coarse.g, coarse.r, fine.g, fine.r);
  std::memset(coarse.x, 0, sizeof(float) * coarse.g.cells);
  if (l + 2 < s.levels.size()) vcycle(s, l + 1);
  pcg(coarse);
  prolongate(fine.g, fine.eps, coarse.g, coarse.x);
  increment(fine);
}
// solver!(ml;tol=1e-4,itmx=32) (:87-99) and solver!(p::Poisson;tol=1e-4,itmx=1e3) (Poisson.jl:162-172)
int solver(Sim& s, double tol, int itmx) {
  Pois& p = s.levels[0];
  residual(p); float r2 = L2(p);
  int n = 0;
  while (n < itmx) {
    if (s.pois_kind == WL_POIS_MULTILEVEL) vcycle(s, 0);
    pcg(p); r2 = L2(p); ++n;
    if ((double)r2 < tol) break;
  }
  perBC(p.g, p.x, p.per);
  s.n_log.push_back(n);
  return n;
}

// BDIM! (Flow.jl:131-135) with μddn (:18-24)
void BDIM(Sim& a) {
  const Grid& g = a.g; const int n = a.D; const float dt = a.dt.back();
  float* f = a.f.data(); const float* u0 = a.u0.data(); const float* V = a.V.data();
  const int64_t tot = g.cells * n;
#pragma omp parallel for schedule(static)
  for (int64_t I = 0; I < tot; ++I) f[I] = u0[I] + dt * f[I] - V[I];
  float* u = a.u.data(); const float* mu0 = a.mu0.data(); const float* mu1 = a.mu1.data();
  for (int i = 0; i < n; ++i) {
    loop(g, r_inside(g), [&](int64_t I, int, int, int) {
      const int64_t Ii = I + i * g.cells;
      float s = 0.f;
      for (int j = 0; j < n; ++j) s += mu1[I + (i + (int64_t)n * j) * g.cells] * (f[Ii + g.s[j]] - f[Ii - g.s[j]]);
      const double add = (0.5 * (double)s + (double)V[Ii]) + (double)(mu0[Ii] * f[Ii]);
      u[Ii] = (float)((double)u[Ii] + add);
    });
  }
}
// scale_u! (Flow.jl:170)
void scale_u(Sim& a, double scale) {
  float* u = a.u.data();
  for (int i = 0; i < a.D; ++i)
    loop(a.g, r_inside(a.g), [&](int64_t I, int, int, int) { u[I + i * a.g.cells] = (float)((double)u[I + i * a.g.cells] * scale); });
}
// accelerate! (Flow.jl:68-73)
void accelerate(Sim& a, const double* gacc) {
  float* f = a.f.data();
  for (int i = 0; i < a.D; ++i) {
    const double gi = gacc[i];
#pragma omp parallel for schedule(static)
    for (int64_t I = 0; I < a.g.cells; ++I) f[I + i * a.g.cells] = (float)((double)f[I + i * a.g.cells] + gi);
  }
}
// project! (Flow.jl:137-145) with div (:11-17)
int project(Sim& a, double w) {
  const Grid& g = a.g; Pois& b = a.levels[0]; float* u = a.u.data();
  // dt = w*Δt[end]: Float32 for the Int w=1, Float64 for w=0.5
  const float dt32 = a.dt.back(); const double dtw = w * (double)dt32;
  const bool f32 = (w == 1.0);
  loop(g, r_inside(g), [&](int64_t I, int, int, int) {
    float s = 0.f;
    for (int i = 0; i < g.D; ++i) s += u[I + g.s[i] + i * g.cells] - u[I + i * g.cells];
    b.z[I] = s;
  });
#pragma omp parallel for schedule(static)
  for (int64_t I = 0; I < g.cells; ++I) b.x[I] = f32 ? b.x[I] * dt32 : (float)((double)b.x[I] * dtw);
  const int n = solver(a, 1e-4, a.pois_kind == WL_POIS_MULTILEVEL ? 32 : 1000);
  for (int i = 0; i < g.D; ++i)
    loop(g, r_inside(g), [&](int64_t I, int, int, int) { u[I + i * g.cells] -= b.L[I + i * g.cells] * (b.x[I] - b.x[I - g.s[i]]); });
#pragma omp parallel for schedule(static)
  for (int64_t I = 0; I < g.cells; ++I) b.x[I] = f32 ? b.x[I] / dt32 : (float)((double)b.x[I] / dtw);
  return n;
}
// CFL (Flow.jl:172-175) with flux_out (:176-182)
float CFL(Sim& a) {
  const Grid& g = a.g; const float* u = a.u.data(); float* sg = a.sigma.data();
  loop(g, r_inside(g), [&](int64_t I, int, int, int) {
    double s = 0.0;
    for (int i = 0; i < g.D; ++i) s += std::max(0., (double)u[I + g.s[i] + i * g.cells]) + std::max(0., (double)-u[I + i * g.cells]);
    sg[I] = (float)s;
  });
  const float m = maximum(sg, g.cells) + 5 * a.nu;
  return std::min(10.f, 1.f / m);
}
// mom_step! (Flow.jl:153-169)
void mom_step(Sim& a, const double* U, const double* gp, const double* gc, int* n_pred, int* n_corr) {
  a.u0 = a.u; scale_u(a, 0);
  // predictor u → u'
  conv_diff(a.g, a.f.data(), a.u0.data(), a.sigma.data(), a.nu, a.per);
  if (gp) accelerate(a, gp);
  BDIM(a); BC(a.g, a.u.data(), U, a.exitBC, a.per);
  if (a.exitBC) exitBC(a.g, a.u.data(), a.u0.data(), U, a.dt.back());
  *n_pred = project(a, 1.0); BC(a.g, a.u.data(), U, a.exitBC, a.per);
  // corrector u → u¹
  conv_diff(a.g, a.f.data(), a.u.data(), a.sigma.data(), a.nu, a.per);
  if (gc) accelerate(a, gc);
  BDIM(a); scale_u(a, 0.5); BC(a.g, a.u.data(), U, a.exitBC, a.per);
  *n_corr = project(a, 0.5); BC(a.g, a.u.data(), U, a.exitBC, a.per);
  a.dt.push_back(CFL(a));
}

float* field_ptr(Sim& s, int field, int64_t* count) {
  const int64_t c = s.g.cells;
  switch (field) {
    case WL_U: *count = c * s.D; return s.u.data();
    case WL_U0: *count = c * s.D; return s.u0.data();
    case WL_F: *count = c * s.D; return s.f.data();
    case WL_P: *count = c; return s.p.data();
    case WL_SIGMA: *count = c; return s.sigma.data();
    case WL_V: *count = c * s.D; return s.V.data();
    case WL_MU0: *count = c * s.D; return s.mu0.data();
    case WL_MU1: *count = c * s.D * s.D; return s.mu1.data();
  }
  return nullptr;
}
float* level_ptr(Sim& s, int level, int lf, int64_t* count) {
  if (level < 0 || level >= (int)s.levels.size()) return nullptr;
  Pois& p = s.levels[level]; const int64_t c = p.g.cells; *count = c;
  switch (lf) {
    case WL_LVL_L: *count = c * s.D; return p.L;
    case WL_LVL_D: return p.Dg;
    case WL_LVL_ID: return p.iD;
    case WL_LVL_X: return p.x;
    case WL_LVL_EPS: return p.eps;
    case WL_LVL_R: return p.r;
    case WL_LVL_Z: return p.z;
  }
  return nullptr;
}
void per_flags(const Sim& s, const int* perdir, bool* out) {
  for (int k = 0; k < 3; ++k) out[k] = perdir ? (k < s.D && perdir[k] != 0) : s.per[k];
}

// ------------------------------------------------------------------------------------------
// Metrics: surface-integrated loads (src/Metrics.jl:23-31,94-141).  T = promote_type(Float64,
// Float32) = Float64: nds and loc are Float64, the per-cell product is rounded to Float32 when it
// is stored into df (≡ flow.f), and `sum(T, df)` accumulates those Float32 values in Float64.
// ------------------------------------------------------------------------------------------
int set_nds(Sim& s, const double* nds, int64_t count) {
  if (!nds) return s.nds.empty() ? fail(WL_ERR_STATE, "no nds array has been passed yet") : WL_OK;
  if (count != s.g.cells * s.D) return fail(WL_ERR_ARG, "nds: bad size");
  s.nds.assign(nds, nds + count);
  return WL_OK;
}
void sum_df(const Sim& s, double* out) {
  for (int i = 0; i < s.D; ++i) {
    const float* d = s.f.data() + i * s.g.cells;
    double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t I = 0; I < s.g.cells; ++I) acc += (double)d[I];
    out[i] = acc;
  }
}
// pressure_force (Metrics.jl:96-100)
void pressure_force(Sim& s, double* out) {
  std::fill(s.f.begin(), s.f.end(), 0.f);
  loop(s.g, r_inside(s.g), [&](int64_t I, int, int, int) {
    for (int i = 0; i < s.D; ++i) s.f[I + i * s.g.cells] = (float)((double)s.p[I] * s.nds[I + i * s.g.cells]);
  });
  sum_df(s, out);
}
// ∂(i,j,I,u) (Metrics.jl:27-31): ∂uᵢ/∂xⱼ at the centre of cell I
inline float dudx(const Sim& s, int i, int j, int64_t I) {
  const float* u = s.u.data() + i * s.g.cells;
  const int64_t si = s.g.s[i], sj = s.g.s[j];
  if (i == j) return u[I + si] - u[I];
  return (u[I + sj] + u[I + sj + si] - u[I - sj] - u[I - sj + si]) / 4;
}
// viscous_force (Metrics.jl:114-118): df[I,:] = -ν*∇²u(I,u)*nds, ∇²u[i,j] = ∂(i,j)+∂(j,i) (:107-108)
void viscous_force(Sim& s, double* out) {
  std::fill(s.f.begin(), s.f.end(), 0.f);
  const float mnu = -s.nu;
  loop(s.g, r_inside(s.g), [&](int64_t I, int, int, int) {
    for (int i = 0; i < s.D; ++i) {
      double acc = 0.0;
      for (int j = 0; j < s.D; ++j) {
        const float m = mnu * (dudx(s, i, j, I) + dudx(s, j, i, I));
        const double term = (double)m * s.nds[I + j * s.g.cells];
        acc = j == 0 ? term : acc + term;
      }
      s.f[I + i * s.g.cells] = (float)acc;
    }
  });
  sum_df(s, out);
}
// pressure_moment (Metrics.jl:137-141): df[I,:] = p[I]*cross(loc(0,I)-x₀, nds)
void pressure_moment(Sim& s, const double* x0, double* out) {
  std::fill(s.f.begin(), s.f.end(), 0.f);
  loop(s.g, r_inside(s.g), [&](int64_t I, int i1, int j1, int k1) {
    const double r[3] = {(i1 - 1.5) - x0[0], (j1 - 1.5) - x0[1], s.D > 2 ? (k1 - 1.5) - x0[2] : 0.0};
    double n[3] = {0, 0, 0};
    for (int i = 0; i < s.D; ++i) n[i] = s.nds[I + i * s.g.cells];
    const double p = (double)s.p[I];
    if (s.D == 2) {
      const float v = (float)(p * (r[0] * n[1] - r[1] * n[0]));
      s.f[I] = v; s.f[I + s.g.cells] = v;
    } else {
      s.f[I] = (float)(p * (r[1] * n[2] - r[2] * n[1]));
      s.f[I + s.g.cells] = (float)(p * (r[2] * n[0] - r[0] * n[2]));
      s.f[I + 2 * s.g.cells] = (float)(p * (r[0] * n[1] - r[1] * n[0]));
    }
  });
  sum_df(s, out);
}

// ------------------------------------------------------------------------------------------
// Metrics: derived fields, `@inside σ[I] = f(I,u)` (src/Metrics.jl:14-77)
// ------------------------------------------------------------------------------------------
// eigvals(Hermitian(A))[2] of a symmetric 3×3 matrix (the reference calls LAPACK): trigonometric closed form
double mid_eig_sym3(const double A[3][3]) {
  const double p1 = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
  const double q = (A[0][0] + A[1][1] + A[2][2]) / 3.0;
  if (p1 == 0.0) {
    double d[3] = {A[0][0], A[1][1], A[2][2]};
    std::sort(d, d + 3);
    return d[1];
  }
  double B[3][3];
  double p2 = 2.0 * p1;
  for (int i = 0; i < 3; ++i) { const double t = A[i][i] - q; p2 += t * t; }
  const double pp = std::sqrt(p2 / 6.0);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = (A[i][j] - (i == j ? q : 0.0)) / pp;
  const double det = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[1][2]) - B[0][1] * (B[0][1] * B[2][2] - B[1][2] * B[0][2]) +
                     B[0][2] * (B[0][1] * B[1][2] - B[1][1] * B[0][2]);
  const double r = std::min(1.0, std::max(-1.0, det / 2.0));
  const double phi = std::acos(r) / 3.0;
  const double e1 = q + 2.0 * pp * std::cos(phi), e3 = q + 2.0 * pp * std::cos(phi + 2.0943951023931954923084289221863);
  return 3.0 * q - e1 - e3;
}
int derived(Sim& s, int kind, int comp, const double* U, const double* axis, const double* center) {
  const Grid& g = s.g;
  const int D = s.D;
  if (kind < WL_DF_KE || kind > WL_DF_OMEGA_THETA) return fail(WL_ERR_ARG, "wl_derived: bad kind");
  if (kind == WL_DF_CURL && (comp < 1 || comp > 3 || (D == 2 && comp != 3))) return fail(WL_ERR_ARG, "wl_derived: curl component must be 1..3 (3 in 2-D)");
  if (kind >= WL_DF_OMEGA_MAG && D != 3) return fail(WL_ERR_ARG, "wl_derived: this diagnostic needs a 3-D flow");
  if (kind == WL_DF_OMEGA_THETA && (!axis || !center)) return fail(WL_ERR_ARG, "wl_derived: ω_θ needs an axis and a centre");
  loop(g, r_inside(g), [&](int64_t I, int i1, int j1, int k1) {
    float out = 0.f;
    if (kind == WL_DF_KE) {  // ke(I,u,U) (:20-22)
      if (U) {
        double acc = 0.0;
        for (int i = 0; i < D; ++i) { const double t = (double)(s.u[I + i * g.cells] + s.u[I + g.s[i] + i * g.cells]) - 2.0 * U[i]; acc += t * t; }
        out = (float)(0.125 * acc);
      } else {
        float acc = 0.f;
        for (int i = 0; i < D; ++i) { const float t = s.u[I + i * g.cells] + s.u[I + g.s[i] + i * g.cells]; acc += t * t; }
        out = (float)(0.125 * (double)acc);
      }
    } else if (kind == WL_DF_CURL) {  // curl(i,I,u) (:55) with ∂(a,I,f) = f[I]-f[I-δa] (Flow.jl:1)
      const int i = comp - 1, j = (i + 1) % 3, k = (i + 2) % 3;
      const float fjk = s.u[I + k * g.cells] - s.u[I - g.s[j] + k * g.cells];
      const float fkj = s.u[I + j * g.cells] - s.u[I - g.s[k] + j * g.cells];
      out = fjk - fkj;
    } else {
      float J[3][3];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[i][j] = dudx(s, i, j, I);
      const float w[3] = {J[2][1] - J[1][2], J[0][2] - J[2][0], J[1][0] - J[0][1]};  // ω(I,u) (:61)
      if (kind == WL_DF_OMEGA_MAG) {
        out = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      } else if (kind == WL_DF_LAMBDA2) {  // λ₂ (:40-44)
        float Sm[3][3], Om[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Sm[i][j] = (J[i][j] + J[j][i]) / 2; Om[i][j] = (J[i][j] - J[j][i]) / 2; }
        double M[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
          float s2 = 0.f, o2 = 0.f;
          for (int k = 0; k < 3; ++k) { s2 += Sm[i][k] * Sm[k][j]; o2 += Om[i][k] * Om[k][j]; }
          M[i][j] = (double)(s2 + o2);
        }
        out = (float)mid_eig_sym3(M);
      } else {  // ω_θ (:74-78)
        const int c1[3] = {i1, j1, k1};
        double d[3], th[3];
        for (int i = 0; i < 3; ++i) d[i] = (double)(float)(c1[i] - 1.5) - center[i];
        th[0] = axis[1] * d[2] - axis[2] * d[1]; th[1] = axis[2] * d[0] - axis[0] * d[2]; th[2] = axis[0] * d[1] - axis[1] * d[0];
        const double n = std::sqrt(th[0] * th[0] + th[1] * th[1] + th[2] * th[2]);
        out = n == 0.0 ? 0.f : (float)((th[0] * (double)w[0] + th[1] * (double)w[1] + th[2] * (double)w[2]) / n);
      }
    }
    s.sigma[I] = out;
  });
  return WL_OK;
}

}  // namespace

#define S ((Sim*)s)
extern "C" {

int wlo_create(wl_sim** out, int D, const int* N, double nu, double dt0, const int* perdir, int exitBC, int pois_kind, int) {
  if (!out || !N || (D != 2 && D != 3)) return fail(WL_ERR_ARG, "wl_create: D must be 2 or 3");
  for (int k = 0; k < D; ++k) if (N[k] < 1) return fail(WL_ERR_ARG, "wl_create: N must be positive");
  Sim* a = new Sim; a->D = D;
  int Ng[3]; for (int k = 0; k < D; ++k) Ng[k] = N[k] + 2;
  a->g.set(D, Ng);
  const int64_t c = a->g.cells;
  a->u.assign(c * D, 0.f); a->u0.assign(c * D, 0.f); a->f.assign(c * D, 0.f);
  a->p.assign(c, 0.f); a->sigma.assign(c, 0.f); a->V.assign(c * D, 0.f);
  a->mu0.assign(c * D, 1.f); a->mu1.assign(c * D * D, 0.f);
  a->nu = (float)nu; a->dt.push_back((float)dt0); a->exitBC = exitBC != 0; a->pois_kind = pois_kind;
  for (int k = 0; k < D; ++k) a->per[k] = perdir && perdir[k];
  const double zero[3] = {0, 0, 0};
  BC(a->g, a->mu0.data(), zero, false, a->per);  // Flow.jl:119
  make_level1(*a);
  if (pois_kind == WL_POIS_MULTILEVEL) {
    auto div_ok = [&](const Pois& p) { for (int k = 0; k < D; ++k) if (!divisible(p.g.n[k])) return false; return true; };
    while (div_ok(a->levels.back()) && (int)a->levels.size() <= 10) {
      // keep pointers of existing levels valid: vector<Pois> moves, store vectors keep their heap blocks
      push_coarse(*a);
    }
    if (a->levels.size() <= 2) { delete a; return fail(WL_ERR_ARG, "MultiLevelPoisson requires size=a2ⁿ, where n>2"); }
  }
  *out = (wl_sim*)a; return WL_OK;
}
// the oracle is single-process: the dist entry points exist only to keep the ABI identical
int wlo_dist_unique_id(char* out128) { if (out128) std::memset(out128, 0, 128); return WL_OK; }
int wlo_create_dist(wl_sim** out, int D, const int* N, double nu, double dt0, const int* perdir, int exitBC, int pois_kind, int device,
                    int rank, int nranks, const char*) {
  if (nranks != 1 || rank != 0) return fail(WL_ERR_ARG, "the oracle is not distributed");
  return wlo_create(out, D, N, nu, dt0, perdir, exitBC, pois_kind, device);
}
int wlo_dist_info(wl_sim* s, int* rank, int* nranks, int* zoff, int* nz_local, int64_t* collectives) {
  if (rank) *rank = 0;
  if (nranks) *nranks = 1;
  if (zoff) *zoff = 0;
  if (nz_local) *nz_local = S->D == 3 ? S->g.n[2] - 2 : 1;
  if (collectives) *collectives = 0;
  return WL_OK;
}
int wlo_dist_finalize(void) { return WL_OK; }
int wlo_dist_plan(int, const int*, const int*, int, int, int, int, int*, int*, int*, int*, int*, int*) {
  return fail(WL_ERR_STATE, "the oracle is not distributed");
}
int wlo_destroy(wl_sim* s) { delete S; return WL_OK; }
const char* wlo_last_error(void) { return g_err.c_str(); }
int wlo_set_stream(wl_sim*, void*) { return WL_OK; }
int wlo_sync(wl_sim*) { return WL_OK; }

int wlo_field_count(wl_sim* s, int field, int64_t* count) { return field_ptr(*S, field, count) ? WL_OK : fail(WL_ERR_ARG, "bad field"); }
int wlo_upload(wl_sim* s, int field, const float* host, int64_t count) {
  int64_t c; float* p = field_ptr(*S, field, &c);
  if (!p || c != count) return fail(WL_ERR_ARG, "wl_upload: bad field or size");
  std::memcpy(p, host, sizeof(float) * c); return WL_OK;
}
int wlo_download(wl_sim* s, int field, float* host, int64_t count) {
  int64_t c; float* p = field_ptr(*S, field, &c);
  if (!p || c != count) return fail(WL_ERR_ARG, "wl_download: bad field or size");
  std::memcpy(host, p, sizeof(float) * c); return WL_OK;
}
int wlo_nlevels(wl_sim* s) { return (int)S->levels.size(); }
int wlo_level_dims(wl_sim* s, int level, int* Ng) {
  if (level < 0 || level >= (int)S->levels.size()) return fail(WL_ERR_ARG, "bad level");
  for (int k = 0; k < S->D; ++k) Ng[k] = S->levels[level].g.n[k];
  return WL_OK;
}
int wlo_level_upload(wl_sim* s, int level, int lf, const float* host, int64_t count) {
  int64_t c; float* p = level_ptr(*S, level, lf, &c);
  if (!p || c != count) return fail(WL_ERR_ARG, "wl_level_upload: bad level/field/size");
  std::memcpy(p, host, sizeof(float) * c); return WL_OK;
}
int wlo_level_download(wl_sim* s, int level, int lf, float* host, int64_t count) {
  int64_t c; float* p = level_ptr(*S, level, lf, &c);
  if (!p || c != count) return fail(WL_ERR_ARG, "wl_level_download: bad level/field/size");
  std::memcpy(host, p, sizeof(float) * c); return WL_OK;
}

int wlo_bc(wl_sim* s, int field, const double* A, int saveexit, const int* perdir) {
  int64_t c; float* p = field_ptr(*S, field, &c);
  if (!p || c != S->g.cells * S->D) return fail(WL_ERR_ARG, "wl_bc: not a vector field");
  bool per[3]; per_flags(*S, perdir, per);
  BC(S->g, p, A, saveexit != 0, per); return WL_OK;
}
int wlo_exit_bc(wl_sim* s, const double* U, double dt) { exitBC(S->g, S->u.data(), S->u0.data(), U, (float)dt); return WL_OK; }
int wlo_per_bc(wl_sim* s, int field, const int* perdir) {
  int64_t c; float* p = field_ptr(*S, field, &c);
  if (!p || c != S->g.cells) return fail(WL_ERR_ARG, "wl_per_bc: not a scalar field");
  bool per[3]; per_flags(*S, perdir, per);
  perBC(S->g, p, per); return WL_OK;
}
int wlo_conv_diff(wl_sim* s, int src) {
  if (src != WL_U && src != WL_U0) return fail(WL_ERR_ARG, "wl_conv_diff: src must be WL_U or WL_U0");
  conv_diff(S->g, S->f.data(), src == WL_U ? S->u.data() : S->u0.data(), S->sigma.data(), S->nu, S->per); return WL_OK;
}
int wlo_accelerate(wl_sim* s, const double* g) { if (g) accelerate(*S, g); return WL_OK; }
int wlo_bdim(wl_sim* s) { BDIM(*S); return WL_OK; }
int wlo_scale_u(wl_sim* s, double scale) { scale_u(*S, scale); return WL_OK; }
int wlo_project(wl_sim* s, double w, int* n) { const int k = project(*S, w); if (n) *n = k; return WL_OK; }
int wlo_cfl(wl_sim* s, double* dt) { *dt = (double)CFL(*S); return WL_OK; }

int wlo_pois_update(wl_sim* s) { update(*S); return WL_OK; }
int wlo_pois_mult(wl_sim* s) { mult_bang(S->levels[0], S->levels[0].x); return WL_OK; }
int wlo_pois_residual(wl_sim* s, int level, double* r2) {
  if (level < 0 || level >= (int)S->levels.size()) return fail(WL_ERR_ARG, "bad level");
  residual(S->levels[level]); if (r2) *r2 = (double)L2(S->levels[level]); return WL_OK;
}
int wlo_pois_jacobi(wl_sim* s, int level, int it) {
  if (level < 0 || level >= (int)S->levels.size()) return fail(WL_ERR_ARG, "bad level");
  jacobi(S->levels[level], it); return WL_OK;
}
int wlo_pois_pcg(wl_sim* s, int level, int it) {
  if (level < 0 || level >= (int)S->levels.size()) return fail(WL_ERR_ARG, "bad level");
  pcg(S->levels[level], it); return WL_OK;
}
int wlo_pois_vcycle(wl_sim* s, int level) {
  if (level < 0 || level + 1 >= (int)S->levels.size()) return fail(WL_ERR_ARG, "bad level");
  vcycle(*S, (size_t)level); return WL_OK;
}
int wlo_pois_head(wl_sim* s) {  // MultiLevelPoisson.jl:88 (residual!) and :72-75 (Jacobi!, restrict!, fill!)
  if (S->pois_kind != WL_POIS_MULTILEVEL || S->levels.size() < 2) return fail(WL_ERR_ARG, "MultiLevelPoisson with at least two levels required");
  Pois& fine = S->levels[0]; Pois& coarse = S->levels[1];
  residual(fine);
  jacobi(fine);
  restrict_r(coarse.g, coarse.r, fine.g, fine.r);
  std::memset(coarse.x, 0, sizeof(float) * coarse.g.cells);
  return WL_OK;
}
int wlo_pois_solve(wl_sim* s, double tol, int itmx, int* n) {
  if (tol < 0) tol = 1e-4;
  if (itmx < 0) itmx = S->pois_kind == WL_POIS_MULTILEVEL ? 32 : 1000;
  const int k = solver(*S, tol, itmx); if (n) *n = k; return WL_OK;
}
int wlo_pressure_force(wl_sim* s, const double* nds, int64_t count, double* force) {
  const int rc = set_nds(*S, nds, count); if (rc) return rc;
  pressure_force(*S, force); return WL_OK;
}
int wlo_viscous_force(wl_sim* s, const double* nds, int64_t count, double* force) {
  const int rc = set_nds(*S, nds, count); if (rc) return rc;
  viscous_force(*S, force); return WL_OK;
}
int wlo_pressure_moment(wl_sim* s, const double* x0, const double* nds, int64_t count, double* moment) {
  const int rc = set_nds(*S, nds, count); if (rc) return rc;
  pressure_moment(*S, x0, moment); return WL_OK;
}
int wlo_derived(wl_sim* s, int kind, int comp, const double* U, const double* axis, const double* center) {
  return derived(*S, kind, comp, U, axis, center);
}
int wlo_mom_step(wl_sim* s, const double* U, const double* gp, const double* gc, double* dt_next, int* n_pred, int* n_corr) {
  int a = 0, b = 0; mom_step(*S, U, gp, gc, &a, &b);
  if (dt_next) *dt_next = (double)S->dt.back();
  if (n_pred) *n_pred = a;
  if (n_corr) *n_corr = b;
  return WL_OK;
}
int wlo_get_dt(wl_sim* s, double* dt) { *dt = (double)S->dt.back(); return WL_OK; }
int wlo_set_dt(wl_sim* s, double dt) { S->dt.back() = (float)dt; return WL_OK; }
int64_t wlo_launch_count(wl_sim*) { return 0; }
int wlo_profile(wl_sim*, int) { return WL_OK; }
int wlo_profile_read(wl_sim*, int, double* ms, int64_t* launches, double* bytes) {
  if (ms) *ms = 0;
  if (launches) *launches = 0;
  if (bytes) *bytes = 0;
  return WL_OK;
}
int wlo_set_sum_mode(int mode) { g_sum_mode = mode; return WL_OK; }
// the flux helpers on a 1-D Float32 vector (0-based index I, unit stride) — the reference pins them directly (maintests.jl:120-155).
// which: 0 quick(f[0],f[1],f[2]), 1 ϕ(1,I,f), 2 ϕu(1,I,f,u), 3 ϕuP(1,Ip,I,f,u), 4 ϕuL(1,I,f,u), 5 ϕuR(1,I,f,u)
double wlo_debug_flux(int which, const float* f, int I, int Ip, double u) {
  switch (which) {
    case 0: return (double)quick(f[0], f[1], f[2]);
    case 1: return phi(f, I, 1);
    case 2: return phiu(f, I, 1, u);
    case 3: return phiuP(f, Ip, I, 1, u);
    case 4: return phiuL(f, I, 1, u);
    default: return phiuR(f, I, 1, u);
  }
}
int wlo_profile_count(void) { return 0; }
const char* wlo_profile_name(int) { return ""; }
// torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py --impl reference asks for all host cores explicitly
int wlo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
  return WL_OK;
}
int wlo_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
#pragma omp master
  n = omp_get_num_threads();
#endif
  return n;
}
}  // extern "C"
