// wl_internal.h — host-side state of one wl_sim handle and the launch functions shared by the
// translation units of libwaterlily_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/waterlily_b200.h"
#include "wl_common.cuh"
#include "wl_p2p.cuh"

// profiled kernel classes (wl_profile_read `which`)
enum {
  WL_PROF_CONVDIFF = 0,   // conv_diff!+accelerate!+BDIM! stage 1 (one kernel)
  WL_PROF_BDIM2,          // BDIM! stage 2 (+scale_u!)
  WL_PROF_PCG_MULT,       // fine level: z = Aϵ, z⋅ϵ
  WL_PROF_PCG_UPDATE,     // fine level: x += αϵ, r -= αz, r⋅z, r⋅r
  WL_PROF_PCG_DIR,        // fine level: ϵ = βϵ + r·iD
  WL_PROF_PCG_INIT,       // fine level: ϵ = r·iD, r⋅z
  WL_PROF_INCREMENT,      // fine level: increment! (r -= Aϵ, x += ϵ)
  WL_PROF_RESIDUAL,       // fine level: residual!
  WL_PROF_PROJECT,        // div, x scaling, pressure correction
  WL_PROF_VCTAIL,         // fused coarse V-cycle kernel (wl_vcycle.cu)
  WL_PROF_SOLVER,         // whole solver! call (all levels)
  WL_PROF_STEP,           // whole mom_step!
  WL_PROF_N
};

struct ProfClass {
  std::vector<cudaEvent_t> ev;  // start/stop pairs
  size_t used = 0;
  double ms = 0;
  int64_t launches = 0;
  double bytes = 0;
  const int* d_counter = nullptr;  // device-side count of launches that were not predicated off
  double bytes_per_active = 0;     // algorithmic bytes of one active launch (with d_counter)
  int counter_base = 0;
};

struct DistState {
  int rank = 0, nranks = 1;
  void* comm = nullptr;  // ncclComm_t (bootstrap, and the data path when peer mapping is off)
  // peer-memory path (wl_p2p.cuh): arenas of all ranks mapped over CUDA IPC
  bool p2p = false;
  bool ready = false;            // wl_create_dist completed (the collective teardown handshake is safe)
  P2PDev host{};                 // host copy of the device context
  P2PDev* dev = nullptr;         // the same in device memory (kernels take the pointer)
  int64_t pushes = 0;            // peer-push kernels launched
};

struct Sim {
  int D = 0;
  GridD g{};
  int device = 0;
  int nsm = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // flow fields (pointers already offset by WL_OX); u/u0 are swapped by mom_step
  float* w = nullptr;  // fourth velocity-sized buffer: the corrector writes the new u here (BDIM! stage 2 folded into conv_diff!), then u ↔ w
  float *u = nullptr, *u0 = nullptr, *f = nullptr, *p = nullptr, *sigma = nullptr, *V = nullptr, *mu0 = nullptr, *mu1 = nullptr;
  std::vector<void*> allocs;
  char* arena = nullptr;       // all fields live in one allocation (symmetric across ranks, wl_p2p.cuh)
  size_t arena_bytes = 0, arena_used = 0;
  double* nds = nullptr;       // Metrics: nds(body, loc(0,I), t) in Float64, D components, field indexing
  void* stage = nullptr;       // device staging buffer of the host <-> device field copies
  size_t stage_bytes = 0;
  unsigned char* body_mask = nullptr;  // 1 byte per cell, same indexing as a scalar field (see k_body_mask)
  bool mask_dirty = true;              // V, μ₀ or μ₁ changed since the mask was built
  unsigned char* tile_clean = nullptr; // [plane][tile row][tile column] of the conv_diff! core kernel: 1 = BDIM! stage 2 folds into its store
  bool bdim_fuse = true;               // WL_BDIM_FUSE=0: conv_diff! writes f everywhere and BDIM! stage 2 runs on every cell
  // flow.f after a fused step: the clean tiles never received the corrector's f.  It is recomputed on demand
  // (download of f, wl_bdim, …) from the inputs the corrector used, which are all still in memory: u' in w, u⁰, V.
  bool f_stale = false;
  float f_dt = 0.f;
  double f_g[3] = {0, 0, 0};
  bool f_has_g = false;
  // flow.p between project! calls of mom_step!: still scaled by dt, the division is pending (k_scale_alloc)
  int p_pend = 0;                      // 0 none, 1 divide by p_pend32 (Float32 dt), 2 divide by p_pend64 (Float64 dt)
  float p_pend32 = 0.f;
  double p_pend64 = 0.0;
  // … and, when the next solver! opens with the fused head (wl_resjac.cu), the divided pressure of the interior cells
  // already sits in levels[0].eps (written by the pressure correction): the head reads x from there, applies the next
  // project!'s x .*= dt while staging and writes the smoothed x straight into the x array — no scaling pass at all
  bool xdiv_valid = false;
  bool xdiv_on = false;                // WL_XDIV=1 switches this on.  Measured at 512³ (profiles/r02_ab_fusions.txt): the divisions cost the two pressure
                                       // corrections more (+0.24 ms each) than the scaling pass they replace (0.21 ms) — off by default
  int xmul_op = 0;                     // set by project! for the head it is about to run: 0 x is read as is, 1 ·xmul32 (Float32 dt), 2 ·xmul64 (Float64 dt)
  float xmul32 = 0.f;
  double xmul64 = 0.0;
  bool cflfuse_on = true;              // WL_CFL_FUSE=0: the corrector's pressure correction and CFL stay separate kernels
  bool cfl_fused = false;              // the corrector's project! already produced σ = flux_out and its maximum away from the upper boundaries
  bool last_cd_fused = false;          // the last conv_diff! folded BDIM! stage 2 into its store on the clean tile-planes
  float nu = 0.f;
  std::vector<float> dt;
  bool exitBC = false;
  int pois_kind = WL_POIS_MULTILEVEL;
  std::vector<Lvl> levels;
  LevelScalars* d_scalars = nullptr;  // one per level
  MiscScalars* d_misc = nullptr;
  ReduceWS ws{};
  float* h_pinned = nullptr;  // pinned readback buffer (64 floats)
  std::vector<int> n_log;
  std::vector<int> pcg_cnt;    // host mirror of LevelScalars::cnt_mult at the last check
  std::vector<int> pcg_guess;  // sweeps the last pcg! needed, per level (speculative enqueue)
  int64_t launches = 0;
  int64_t collectives = 0;
  DistState dist;
  std::vector<int> lvl_dist;   // 1: level is z-distributed, 0: replicated on every rank (coarse tail)
  int Nz_global = 0;           // interior z-planes of the whole domain (3-D)
  int per_z_global = 0;
  void* cd_tma = nullptr;      // TMA tensor maps of u/u⁰ for the staged conv_diff! kernel (wl_convdiff_tma.cu), or null
  bool fuse_prolong = false;   // inside solver!: Vcycle! ends with the fused prolongate!/increment!/pcg!-init pass
  int pcg_init_done = -1;      // level whose pcg! head (ϵ = r·iD, ρ) that pass already produced
  bool eps_is_jacobi = false;  // levels[1].ϵ already holds r·iD of the current residual (written by k_residual)
  void* vc_fused = nullptr;    // fused coarse V-cycle (wl_vcycle.cu), or null
  int vc_first = -1;           // first level the fused kernel covers (-1: off)
  int last_solve_n[2] = {0, 0};
  bool vc_solve = false;       // the fused kernel runs the whole solver! (one launch, no host read)
  // uniform-coefficient packs (Lvl::umask): per-level mask storage, non-uniform pack counters, switch
  std::vector<unsigned char*> umask_store;
  unsigned int* d_nonuni = nullptr;  // [levels] packs that are NOT uniform, counted by k_set_diag
  bool uniform_path = true;          // WL_UNIFORM=0 switches the fast path off (every stencil loads L, D, iD)
  bool march_on = true;              // WL_MARCH=0: the stencil kernels iterate (row-group, plane) units instead of marching in z
  bool stage_on = true;              // WL_STAGE=0: marching kernels keep their z-window in registers instead of a shared-memory ring
  bool resjac_on = true;             // WL_RESJAC=0: residual!, the first Jacobi! and restrict! of solver! stay separate kernels (wl_resjac.cu)
  bool resjac_div = false;           // WL_RESJAC_DIV=1: the fused head evaluates project!'s z = div u itself instead of reading the z a pass of its own stored.
                                     // Measured at 512³: the head slows down by 0.52 ms, the pass it replaces takes 0.42 ms — off by default
  bool head_fused = false;           // solver!: the fused head already ran Jacobi! + restrict! of the first Vcycle! on the finest level
  bool x_parked = false;             // … and the finest x lives in levels[0].eps until the prolongation pass moves it back
  bool nonuni_known = false;
  unsigned int nonuni0 = 0;
  double uni_frac0 = 0.0;            // fraction of uniform packs on the finest level (profiling: algorithmic bytes)
  int sticky = 0;  // first CUDA error, sticky per handle
  bool profiling = false;
  ProfClass prof[WL_PROF_N];
};

int wl_fail(int code, const std::string& msg);
int wl_cuda_check(Sim& s, cudaError_t e, const char* what);
#define WL_CUDA(s, call) do { int _rc = wl_cuda_check((s), (call), #call); if (_rc) return _rc; } while (0)

// launch geometry
dim3 wl_pick_block(const GridD& g);
int wl_grid_for(const Sim& s, int64_t ntiles);

// profiling brackets (no-ops unless wl_profile(s,1))
void wl_prof_begin(Sim& s, int cls);
void wl_prof_end(Sim& s, int cls, double bytes);

// ---- wl_dist.cu ----
int dist_unique_id(char* out128);
int dist_init(Sim& s, int rank, int nranks, const char* id128);
int dist_p2p_setup(Sim& s);   // after the arena exists: exchange IPC handles, map the peers
int dist_check_p2p_error(Sim& s);  // a peer wait timed out → sticky failure
void dist_destroy(Sim& s);
int dist_finalize();
int dist_halo(Sim& s, const GridD& g, float* base, int ncomp, int depth);
int dist_allreduce(Sim& s, double* dptr, int count, bool is_max);
int dist_allgather(Sim& s, float* base, size_t count);
// runs finalize(stage) after the all-reduce of fa.raw when the producing kernel deferred it
int post_reduce(Sim& s, int stage, const FinArgs& fa, bool is_max = false);
FinArgs fin_level(const Sim& s, int level);
FinArgs fin_misc(const Sim& s);

// ---- wl_convdiff_tma.cu ----
struct CDArgs;
int convdiff_tma_setup(Sim& s);
void convdiff_tma_destroy(Sim& s);
bool convdiff_tma_launch(Sim& s, const CDArgs& a, int zchunk, bool corr);
int convdiff_tma_core_x0(const Sim& s);
bool convdiff_tma_tiles(const Sim& s, int* ntx, int* nty);

// ---- wl_metrics.cu ----
// which: 0 pressure_force, 1 viscous_force, 2 pressure_moment (x0 required); out has D entries
int metrics_run(Sim& s, int which, const double* x0, double* out);
int derived_run(Sim& s, int kind, int comp, const double* U, const double* axis, const double* center);

// ---- wl_vcycle.cu ----
int vcycle_fused_setup(Sim& s);
void vcycle_fused_destroy(Sim& s);
int vcycle_fused_run(Sim& s);
int vcycle_fused_trace(Sim& s, unsigned long long* out, int max_pairs, int* n);
int vcycle_fused_solve(Sim& s, double tol, int itmx, bool ghosts_valid, int slot, const float* divu = nullptr);

// ---- wl_bc.cu ----
int bc_vec(Sim& s, const GridD& g, float* a, const double* A, bool saveexit, const int* per, int halo_depth = 1);
int bc_per(Sim& s, const GridD& g, float* a, const int* per, const int* active = nullptr);
int bc_exit(Sim& s, const double* U, float dt);

// ---- wl_resjac.cu ----
bool resjac_eligible(const Sim& s);
bool pois_head_fusable(const Sim& s);  // solver!(ml) on this handle opens with the fused head
int resjac_run(Sim& s, const float* divu = nullptr);
int resjac_unpark(Sim& s);  // moves the parked x' from levels[0].eps back into levels[0].x (API inspection only)

// ---- wl_poisson.cu ----
int pois_set_diag(Sim& s, Lvl& l);
void pois_invalidate_uniform(Sim& s, int level);
int pois_uniform_fraction(Sim& s);  // refreshes s.uni_frac0 (one scalar read)
// algorithmic words per cell of a finest-level kernel: `generic` when every pack loads its coefficients,
// `saved` words less on uniform packs, plus the 1/16 word of the pack mask
inline double wl_words(const Sim& s, double generic, double saved) {
  const bool on = !s.levels.empty() && s.levels[0].umask != nullptr;
  return on ? generic - saved * s.uni_frac0 + 0.0625 : generic;
}
int pois_restrictL(Sim& s, Lvl& coarse, const Lvl& fine);
int pois_update(Sim& s);
int pois_mult(Sim& s);
int pois_residual(Sim& s, Lvl& l, bool need_r2 = true, bool ghosts_valid = false, bool jacobi_eps = false, const float* divu = nullptr);
int pois_jacobi(Sim& s, Lvl& l, int it);
int pois_pcg(Sim& s, int level, int it);
int pois_vcycle(Sim& s, int level);
int pois_head(Sim& s);  // residual! + Jacobi! + restrict! + fill!(coarse.x,0) on the finest level
int pois_solve(Sim& s, double tol, int itmx, int* n, bool ghosts_valid = false, int slot = 0, bool defer_count = false, const float* divu = nullptr);
int pois_materialize_eps(Sim& s, int level);
int pois_dot_rr(Sim& s, Lvl& l, double* out);

// ---- wl_momentum.cu ----
int mom_conv_diff(Sim& s, const float* src, bool fuse_bdim1, const double* g, int bdim2_mode = 0, float dt_override = 0.f);
int mom_materialize_f(Sim& s);  // flow.f of the last fused step, on demand (see Sim::f_stale)
int mom_accelerate(Sim& s, const double* g);
int mom_bdim1(Sim& s);
int mom_bdim2(Sim& s, bool accumulate, float scale, bool fused = false);
int mom_scale_u(Sim& s, double scale);
int mom_project(Sim& s, double w, int* n, int slot = 0, bool defer_count = false, bool in_step = false, bool with_cfl = false);
int mom_div(Sim& s, const float* u, float* z);  // z = div u on the interior (Flow.jl:139)
int mom_materialize_p(Sim& s);  // applies a pending x ./= dt (see Sim::p_pend)
int mom_cfl(Sim& s, double* dt);
int mom_step(Sim& s, const double* U, const double* gp, const double* gc, double* dt_next, int* n_pred, int* n_corr);

// generic helpers (wl_api.cu)
int wl_fill(Sim& s, float* p, float v, int64_t n);
int wl_read_scalars(Sim& s, const void* dptr, size_t bytes);  // → h_pinned, synchronises the stream
