/*
 * waterlily_b200.h — C ABI of libwaterlily_b200.so
 *
 * Drop-in boundary for the per-timestep hot path of WaterLily.jl
 * (`mom_step!(::Flow, ::MultiLevelPoisson)`, reference src/Flow.jl:153-169 and everything it
 * calls).  The reference selects its backend by array type (`mem=Array|CuArray`,
 * src/WaterLily.jl:66-78) and has no FFI of its own; these are the entry points a Julia
 * extension module binds with `ccall` (see INTEGRATION.md and julia/WaterLilyB200.jl) and that
 * the Python host mirror in waterlily_b200/ binds with ctypes.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative wl_status code and
 *    never throws across the ABI; wl_last_error() returns the message of the last failure on
 *    the calling thread (the reference's own assertion texts are reproduced verbatim).
 *  - all fields are Float32 (`T=Float32`, src/WaterLily.jl:69).  Host layout is the reference's:
 *    padded (one ghost layer per side) column-major arrays, x fastest, components last:
 *    scalar (Ng...), vector (Ng...,D), tensor (Ng...,D,D).
 *  - the library owns all device memory behind the opaque handle.  Level-1 multigrid arrays
 *    alias the flow arrays exactly as in the reference: x ≡ p, L ≡ μ₀, z ≡ σ
 *    (src/WaterLily.jl:77).
 *  - perdir is passed as D 0/1 flags (flag[j]=1 ⇔ j+1 ∈ perdir of the reference).
 *  - one host thread per handle; calls are synchronous from the caller's view.
 *
 * The identical ABI, with the prefix `wlo_`, is exported by the CPU oracle (oracle/), which is
 * test infrastructure only and is never loaded by this library or the package.
 */
#ifndef WATERLILY_B200_H
#define WATERLILY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wl_sim wl_sim;

enum wl_status {
  WL_OK = 0,
  WL_ERR_ARG = -1,      /* bad argument (AssertionError in the reference) */
  WL_ERR_CUDA = -2,     /* CUDA runtime failure (sticky per handle) */
  WL_ERR_NCCL = -3,     /* NCCL failure (sticky per handle) */
  WL_ERR_STATE = -4     /* call not valid for this handle (e.g. wrong pois_kind) */
};

/* Flow fields, reference src/Flow.jl:92-109 */
enum wl_field {
  WL_U = 0,     /* u   (Ng...,D)   velocity                      */
  WL_U0 = 1,    /* u⁰  (Ng...,D)   previous velocity             */
  WL_F = 2,     /* f   (Ng...,D)   force / RHS                   */
  WL_P = 3,     /* p   (Ng...)     pressure  (≡ levels[1].x)     */
  WL_SIGMA = 4, /* σ   (Ng...)     divergence scratch (≡ levels[1].z, ≡ Φ of conv_diff!) */
  WL_V = 5,     /* V   (Ng...,D)   body velocity                 */
  WL_MU0 = 6,   /* μ₀  (Ng...,D)   zeroth moment (≡ levels[1].L) */
  WL_MU1 = 7    /* μ₁  (Ng...,D,D) first moment                  */
};

/* Poisson level fields, reference src/Poisson.jl:21-30 */
enum wl_level_field {
  WL_LVL_L = 0, WL_LVL_D = 1, WL_LVL_ID = 2, WL_LVL_X = 3,
  WL_LVL_EPS = 4, WL_LVL_R = 5, WL_LVL_Z = 6
};

/* which AbstractPoisson the handle owns (reference src/Poisson.jl:21, src/MultiLevelPoisson.jl:44) */
enum wl_pois_kind { WL_POIS_MULTILEVEL = 0, WL_POIS_SINGLE = 1 };

/* ---- lifetime ------------------------------------------------------------------------- */

/* Flow(N,U;Δt,ν,perdir,exitBC,T=Float32) + [MultiLevel]Poisson(p,μ₀,σ;perdir)
 * (src/Flow.jl:110-121, src/WaterLily.jl:75-77, src/MultiLevelPoisson.jl:51-59).
 * N = interior cells per dimension (D entries, D = 2 or 3).  Fields start as the reference's
 * constructor leaves them before apply!/measure!: u=u⁰=f=p=σ=V=μ₁=0, μ₀=1 with BC!(μ₀,0).
 * WL_POIS_MULTILEVEL fails with "MultiLevelPoisson requires size=a2ⁿ, where n>2"
 * (src/MultiLevelPoisson.jl:56-57) when fewer than 3 levels can be built.
 * device: CUDA ordinal (ignored by the oracle). */
int wl_create(wl_sim** out, int D, const int* N, double nu, double dt0,
              const int* perdir, int exitBC, int pois_kind, int device);
int wl_destroy(wl_sim* s);
const char* wl_last_error(void);
/* cudaStream_t on which all work of this handle is issued (NULL = handle's own stream). */
int wl_set_stream(wl_sim* s, void* cuda_stream);
int wl_sync(wl_sim* s);

/* ---- z-slab domain decomposition over the GPUs of one box (new; the reference is single-device,
 *      README.md:157).  One process per GPU: rank 0 makes an id with wl_dist_unique_id, the launcher
 *      broadcasts its 128 bytes (torch.distributed, MPI, a file …), every rank calls wl_create_dist
 *      with the GLOBAL N.  N[2] is split into N[2]/nranks planes per rank (must be even).  The id
 *      bootstraps a NCCL communicator that exchanges CUDA IPC handles: afterwards halo planes (at the
 *      reference's BC!/perBC! sites) and scalar reductions travel as direct stores into the peers'
 *      memory over NVLink from inside the kernels (WL_P2P=0 keeps NCCL send/recv + all-reduce);
 *      multigrid levels with fewer than 16 planes per rank are replicated on every rank
 *      (agglomeration).  wl_destroy of such a handle is collective too.  All other entry points then act on the rank's slab and are
 *      collective: every rank must make the same calls in the same order.  Host arrays passed to
 *      upload/download always have the GLOBAL shape; upload reads the rank's planes from it,
 *      download fills only the planes the rank owns. ------------------------------------------- */
int wl_dist_unique_id(char* out128);
int wl_create_dist(wl_sim** out, int D, const int* N, double nu, double dt0,
                   const int* perdir, int exitBC, int pois_kind, int device,
                   int rank, int nranks, const char* nccl_id128);
int wl_dist_info(wl_sim* s, int* rank, int* nranks, int* zoff, int* nz_local, int64_t* collectives);
/* Destroys the NCCL communicators of this process that no live handle uses (they are cached per id so
 * that several handles can share one).  Optional; call after the last wl_destroy. */
int wl_dist_finalize(void);
/* Host-only (no GPU touched): the slab plan of `rank` — per multigrid level whether it is distributed,
 * its local interior z-planes, the global z offset of local plane 0 and the halo flags of its faces. */
int wl_dist_plan(int D, const int* N, const int* perdir, int pois_kind, int rank, int nranks, int max_levels,
                 int* nlevels, int* distributed, int* nz_local, int* zoff, int* zlo_halo, int* zhi_halo);

/* ---- field transfer: B200Array(::Array) / Array(::B200Array) ---------------------------- */
int wl_field_count(wl_sim* s, int field, int64_t* count);
int wl_upload(wl_sim* s, int field, const float* host, int64_t count);
int wl_download(wl_sim* s, int field, float* host, int64_t count);
int wl_nlevels(wl_sim* s);
int wl_level_dims(wl_sim* s, int level, int* Ng);            /* padded size, D entries; level is 0-based */
int wl_level_upload(wl_sim* s, int level, int lfield, const float* host, int64_t count);
int wl_level_download(wl_sim* s, int level, int lfield, float* host, int64_t count);

/* ---- boundary conditions (src/util.jl:192-231) ----------------------------------------- */
/* BC!(a,A,saveexit,perdir) on a vector field of the handle; perdir==NULL → the handle's own. */
int wl_bc(wl_sim* s, int field, const double* A, int saveexit, const int* perdir);
/* exitBC!(u,u⁰,U,Δt) */
int wl_exit_bc(wl_sim* s, const double* U, double dt);
/* perBC!(a,perdir) on a scalar field (WL_P or WL_SIGMA) */
int wl_per_bc(wl_sim* s, int field, const int* perdir);

/* ---- momentum pieces (src/Flow.jl) — exported so that each reference function has a parity
 *      test of its own; wl_mom_step runs the fused sequence ---------------------------------- */
/* conv_diff!(f, src, σ; ν, perdir), src = WL_U or WL_U0 (src/Flow.jl:36-60) */
int wl_conv_diff(wl_sim* s, int src_field);
/* accelerate!: f[..,i] += g[i] (src/Flow.jl:68-73) */
int wl_accelerate(wl_sim* s, const double* g);
/* BDIM!(a) (src/Flow.jl:131-135) */
int wl_bdim(wl_sim* s);
/* scale_u!(a,scale) (src/Flow.jl:170) */
int wl_scale_u(wl_sim* s, double scale);
/* project!(a,b,w) (src/Flow.jl:137-145); n = V-cycles (or PCG sweeps) used */
int wl_project(wl_sim* s, double w, int* n);
/* CFL(a) (src/Flow.jl:172-175); does not push to Δt */
int wl_cfl(wl_sim* s, double* dt);

/* ---- pressure solver (src/Poisson.jl, src/MultiLevelPoisson.jl) ------------------------- */
/* update!(pois): set_diag! on level 1, restrictL!+set_diag! on the coarser levels */
int wl_pois_update(wl_sim* s);
/* mult!(pois,x): z = A x with perBC!(x) and zero ghosts; x is levels[1].x */
int wl_pois_mult(wl_sim* s);
/* residual!(levels[level]); r2 = L₂ = r⋅r afterwards */
int wl_pois_residual(wl_sim* s, int level, double* r2);
/* Jacobi!(levels[level];it) */
int wl_pois_jacobi(wl_sim* s, int level, int it);
/* pcg!(levels[level];it) */
int wl_pois_pcg(wl_sim* s, int level, int it);
/* Vcycle!(ml;l=level+1) */
int wl_pois_vcycle(wl_sim* s, int level);
/* The head of solver!(ml) on the finest level: residual!(levels[1]) (src/Poisson.jl:91-97) followed by the opening of
 * the first Vcycle! (src/MultiLevelPoisson.jl:72-75): Jacobi!(fine), restrict!(coarse.r, fine.r), fill!(coarse.x, 0).
 * The library runs these four reference calls as one pass where the grid allows (wl_resjac.cu); this entry exposes
 * exactly that group so that it has its own parity test.  MultiLevelPoisson only. */
int wl_pois_head(wl_sim* s);
/* solver!(pois;tol,itmx): MultiLevelPoisson (tol=1e-4,itmx=32) or Poisson (tol=1e-4,itmx=1e3);
 * pass tol<0 / itmx<0 for the reference defaults.  n = iterations, also appended to the log. */
int wl_pois_solve(wl_sim* s, double tol, int itmx, int* n);

/* ---- Metrics: surface-integrated loads (src/Metrics.jl:84-141) ---------------------------------
 * nds is the BDIM-masked surface normal nds(body, loc(0,I,Float64), t) (src/Metrics.jl:84-87) at every
 * cell of the padded grid, a Float64 array (Ng...,D) evaluated by the host — the geometry stays in
 * Julia, as for measure!.  NULL reuses the array of the previous call (static body).  The per-cell
 * products, the Float32 store into flow.f (the reference's df scratch, overwritten here too) and the
 * Float64 sums run on the device.  Collective on slab handles; every rank receives the global value. */
/* pressure_force(p,df,body,t): sum over inside(p) of p[I]*nds (src/Metrics.jl:94-100); force has D entries */
int wl_pressure_force(wl_sim* s, const double* nds, int64_t count, double* force);
/* viscous_force(u,ν,df,body,t): sum of -ν*∇²u(I,u)*nds (src/Metrics.jl:102-118) */
int wl_viscous_force(wl_sim* s, const double* nds, int64_t count, double* force);
/* pressure_moment(x₀,p,df,body,t): sum of p[I]*cross(loc(0,I)-x₀, nds) (src/Metrics.jl:129-141);
 * D entries (in 2-D the scalar cross product is broadcast into both, as in the reference) */
int wl_pressure_moment(wl_sim* s, const double* x0, const double* nds, int64_t count, double* moment);

/* ---- Metrics: derived fields for output (src/Metrics.jl:14-77) --------------------------------
 * `@inside flow.σ[I] = f(I, flow.u)` for the reference's per-cell diagnostics: the interior of σ
 * is filled on the device (ghost cells untouched, as with @inside), the caller downloads WL_SIGMA.
 *   WL_DF_KE          ke(I,u,U)           ½‖u−U‖² at the cell centre; U = NULL for U=0        (:20-22)
 *   WL_DF_CURL        curl(comp,I,u)      component comp (1-based; 3 in 2-D) of ∇×u at the cell edge (:55)
 *   WL_DF_OMEGA_MAG   ω_mag(I,u)          ‖∇×u‖ at the cell centre (3-D)                       (:61-67)
 *   WL_DF_LAMBDA2     λ₂(I,u)             middle eigenvalue of S²+Ω² (3-D)                     (:40-44)
 *   WL_DF_OMEGA_THETA ω_θ(I,z,center,u)   azimuthal vorticity about the axis `axis` through `center` (3-D) (:74-78) */
enum wl_derived_kind { WL_DF_KE = 0, WL_DF_CURL = 1, WL_DF_OMEGA_MAG = 2, WL_DF_LAMBDA2 = 3, WL_DF_OMEGA_THETA = 4 };
int wl_derived(wl_sim* s, int kind, int comp, const double* U, const double* axis, const double* center);

/* ---- the hot path --------------------------------------------------------------------- */
/* mom_step!(a,b) (src/Flow.jl:153-169).
 * U       : BCTuple(a.U,a.Δt,N) evaluated by the host at t+Δt (D entries)
 * g_pred  : uniform acceleration gᵢ(+dUᵢ/dt) at t, or NULL for `nothing` (src/Flow.jl:158)
 * g_corr  : the same at t+Δt (src/Flow.jl:165)
 * dt_next : CFL(a), already pushed onto the handle's Δt history
 * n_pred, n_corr : the two entries pushed onto pois.n */
int wl_mom_step(wl_sim* s, const double* U, const double* g_pred, const double* g_corr,
                double* dt_next, int* n_pred, int* n_corr);
/* Δt[end] */
int wl_get_dt(wl_sim* s, double* dt);
int wl_set_dt(wl_sim* s, double dt);
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t wl_launch_count(wl_sim* s);
/* Per-kernel-class device time, measured with CUDA events on the handle's stream while profiling
 * is on.  Classes 0..wl_profile_count()-1 are named by wl_profile_name(); read returns the summed
 * milliseconds, the number of launches that did work (kernels of an exited pcg! are predicated off
 * on the device and not counted) and their algorithmic bytes. */
int wl_profile(wl_sim* s, int on);
int wl_profile_count(void);
const char* wl_profile_name(int which);
int wl_profile_read(wl_sim* s, int which, double* ms, int64_t* launches, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* WATERLILY_B200_H */
