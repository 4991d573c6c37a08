/* wl_b200.h — C ABI of the B200-native `mom_step!` library (libwl_b200.so).
 *
 * This is the drop-in boundary for WaterLily.jl's per-time-step hot path.  Host code
 * (Julia `ccall`, Python ctypes, C) owns configuration and set-up; the library owns all
 * device state (re-pitched fields, multigrid hierarchy, reduction cells, streams).
 * Every entry point cites the reference interface it replaces (paths relative to the
 * WaterLily.jl tree, v1.8.0).  The Julia-side binding is shown in INTEGRATION.md.
 *
 * Conventions
 *  - All functions return 0 on success, non-zero on error; wl_last_error() returns a
 *    thread-local message.  Nothing throws, longjmps or calls back into the host runtime.
 *  - Fields cross the boundary in the REFERENCE layout: Float32, ghost-padded
 *    (N_k = n_k+2), column-major with x fastest and the vector component slowest,
 *    i.e. u[x,y,z,c] at linear index x + N1*(y + N2*(z + N3*c)) (0-based).
 *  - `src_is_device`/`dst_is_device` != 0 means the pointer is a CUDA device pointer on
 *    the handle's device (e.g. a CuArray); otherwise host memory.
 *  - One host thread per handle.  Calls are asynchronous with respect to the GPU except
 *    the getters, wl_download to host memory and wl_sync.
 *  - There is no CPU fallback: every compute entry point fails if no CUDA device exists.
 */
#ifndef WL_B200_H
#define WL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wl_handle wl_handle;

/* convective scheme λ(u,c,d): src/Flow.jl:4-6, `λ` keyword of Simulation (src/WaterLily.jl:95) */
enum { WL_QUICK = 0, WL_CDS = 1, WL_VANLEER = 2 };
/* pressure solver: MultiLevelPoisson (default pois_ctor, src/WaterLily.jl:97) or Poisson */
enum { WL_POIS_MULTILEVEL = 0, WL_POIS_SINGLE = 1 };
/* `smooth!` (src/MultiLevelPoisson.jl:106): GaussSeidelRB! (default) or pcg! */
enum { WL_SMOOTH_GSRB = 0, WL_SMOOTH_PCG = 1 };
/* fields of Flow (src/Flow.jl:114-124) addressable through wl_upload / wl_download */
enum { WL_U = 0, WL_U0 = 1, WL_F = 2, WL_P = 3, WL_SIGMA = 4, WL_V = 5, WL_MU0 = 6, WL_MU1 = 7 };
/* arrays of one Poisson level (src/Poisson.jl:22-31) addressable through wl_download_level */
enum { WL_LVL_L = 0, WL_LVL_D = 1, WL_LVL_ID = 2, WL_LVL_X = 3, WL_LVL_EPS = 4, WL_LVL_R = 5, WL_LVL_Z = 6 };

/* Keyword arguments of Simulation / Flow (src/WaterLily.jl:93-98, src/Flow.jl:133-134) that
 * reach the hot path.  Function-valued uBC / g / udf / u0 are host closures and are not
 * representable here: the host evaluates u0 itself and uploads u (see wl_apply_bc). */
typedef struct wl_config {
  int32_t D;         /* 2 or 3 */
  int32_t n[3];      /* interior cells per dimension (`dims`) */
  float uBC[3];      /* constant boundary velocity tuple */
  int32_t perdir[3]; /* perdir[d] != 0: dimension d+1 is periodic */
  int32_t exitBC;    /* convective exit in x (src/core.jl:226) */
  int32_t lambda;    /* WL_QUICK | WL_CDS | WL_VANLEER */
  float nu;          /* kinematic viscosity ν */
  float dt0;         /* initial Δt (default 0.25) */
  int32_t pois_kind; /* WL_POIS_MULTILEVEL | WL_POIS_SINGLE */
  int32_t smoother;  /* WL_SMOOTH_GSRB | WL_SMOOTH_PCG */
  float tol;         /* solver tolerance on Σr² (reference default 1e-4) */
  int32_t itmx;      /* max solver iterations (32 for MultiLevelPoisson, 1000 for Poisson; 0 = default) */
  int32_t device;    /* CUDA device ordinal */
  int32_t flags;     /* WL_FLAG_* */
} wl_config;

enum {
  WL_FLAG_GENERAL_COEFF = 1, /* never use the constant-coefficient (NoBody) kernel variants */
  WL_FLAG_NO_PERSISTENT = 4, /* launch every coarse-level operation separately instead of the one cooperative coarse-level kernel */
  WL_FLAG_NCCL_HALO = 8,     /* multi-GPU: exchange halo planes with ncclSend/ncclRecv instead of the peer-to-peer NVLink kernel */
  WL_FLAG_UNFUSED_GS = 2,    /* run GaussSeidelRB! as six separate launches like the reference (debugging aid) */
  /* A/B switches between kernel variants that produce the same bits (used by the parity tests to compare them): */
  WL_FLAG_NO_VSMOOTH = 16,   /* constant-coefficient mode: separate prolongation / GaussSeidelRB! / increment launches instead of f_vsmooth */
  WL_FLAG_NO_CONV4 = 32,     /* constant-coefficient mode: the one-cell-per-thread flux kernel fm_conv instead of fm_conv4 */
  WL_FLAG_NO_FUSED_UNI = 64, /* constant-coefficient mode: f_div_residual / f_jacobi / f_correct + f_cfl instead of their fused forms */
  WL_FLAG_NO_TINY = 256,     /* run the coarsest levels (≤ 8192 cells) inside the cooperative coarse-level kernel instead of the one-block kernel */
  WL_FLAG_NO_SEMI = 128,     /* general mode: always read the face coefficients L (no semi-uniform march blocks, no body-free BDIM blocks) */
  WL_FLAG_NO_PREFETCH = 1024, /* z slabs: push the halo of r that f_vsmooth reads right before that kernel instead of on a side stream after Jacobi! */
  WL_FLAG_NCCL_ALLREDUCE = 2048, /* z slabs: ncclAllReduce for the solver's and CFL's scalars instead of the one-warp all-reduce over peer memory */
  WL_FLAG_NO_PDL = 4096, /* launch the uniform-mode step's kernels without programmatic dependent launch */
  WL_FLAG_NO_FAST_READ = 512, /* read the solver's residual norms with a copy + stream synchronisation instead of polling the mapped mirror the reduction writes */
};

const char* wl_last_error(void);
int wl_device_count(void);

/* Flow(N,uBC;…) with a tuple initial condition (src/Flow.jl:133-147): allocates u,u⁰,f,p,σ,V,μ₀,μ₁,
 * sets u=uBC, BC!, exitBC!(u,u,0), μ₀=1 with BC!(μ₀,0), Δt=[dt0]; then builds the pressure
 * solver like pois_ctor(flow) (src/WaterLily.jl:105; MultiLevelPoisson ctor src/MultiLevelPoisson.jl:68-76). */
int wl_create(const wl_config* cfg, wl_handle** out);
int wl_destroy(wl_handle* h);
/* The device memory of destroyed handles is kept for the next handle of the process (same device, same chunk sizes), so that a second
 * Simulation of the same shape starts without the cudaFree + cudaMalloc of the whole state; this returns it to the driver. */
int wl_release_pool(void);

/* Multi-GPU (new functionality; the reference has none, README.md:155): the domain is decomposed into z slabs, one process per
 * GPU.  Rank 0 obtains a 128-byte NCCL id with wl_dist_unique_id and the host broadcasts it (torch.distributed, MPI, …); every
 * rank then calls wl_create_dist with the GLOBAL configuration.  Each rank owns dims[3]/nranks planes (an even number >= 4);
 * wl_upload / wl_download move the rank's own slab, ghost planes included: shape (N1, N2, dims[3]/nranks + 2[, ncomp]).
 * Ghost planes move by peer-to-peer stores over NVLink (CUDA IPC mappings of the neighbours' field memory and a flag handshake;
 * grouped ncclSend/ncclRecv when IPC is unavailable), the scalar reductions by ncclAllReduce, and multigrid levels of at most
 * 3 M cells (or fewer than 8 planes per rank) are replicated on every rank (ncclAllGather of the restricted residual).
 * A P-rank run is bit-identical to the single-GPU run. */
int wl_dist_unique_id(void* id128);
int wl_create_dist(const wl_config* cfg, int rank, int nranks, const void* nccl_id128, wl_handle** out);

/* Array accessors replacing direct reads/writes of flow.u, flow.p, flow.σ, flow.f, flow.V, flow.μ₀, flow.μ₁
 * (src/Flow.jl:116-124; used by Metrics/JLD2/VTK extensions).  Buffers hold ncomp*ΠN floats in the reference layout. */
int wl_upload(wl_handle* h, int field, const float* src, int src_is_device);
int wl_download(wl_handle* h, int field, float* dst, int dst_is_device);
/* One component of a vector field (ΠN floats): what apply!(f,c) does per component when it fills u from a function initial
 * condition (src/util.jl `apply!`, called by Flow's constructor src/Flow.jl:140) — lets a host hand over components it holds
 * separately without assembling them first. */
int wl_upload_component(wl_handle* h, int field, int comp, const float* src, int src_is_device);

/* After uploading u from a function initial condition: BC!(u,uBC,exitBC,perdir); exitBC!(u,u,0); u⁰=copy(u)
 * (src/Flow.jl:141-142). */
int wl_apply_bc(wl_handle* h);

/* Tail of measure!(flow,body) (src/Body.jl:49-50) after the host uploaded μ₀ and V:
 * BC!(μ₀,0,false,perdir); BC!(V,0,exitBC,perdir). */
int wl_measure_bc(wl_handle* h);

/* measure!(flow, body; t, ϵ) ON THE DEVICE for bodies the library can evaluate itself (src/Body.jl:28-51,
 * src/AutoBody.jl:29-37, set operations src/Body.jl:88-107): primitive k is AutoBody(sdf_k, (x,t) -> x .- vel_k.*t) with sdf_k a
 * sphere (circle in 2-D) or a torus with its axis along x; the body is ((p₀ op₁ p₁) op₂ p₂) … with op ∈ {∪, ∩, −}.  Everything a
 * host closure would do for these shapes — signed distance at the cell centres (flow.σ), distance, normal and velocity at the
 * faces of the band |d| < 2+ϵ, μ₀, μ₁, V, then BC!(μ₀,0), BC!(V,0,exitBC) — runs in one kernel; the Poisson hierarchy is marked
 * stale (rebuilt by wl_update or by the next step).  This is what makes sim_step!(sim; remeasure=true) possible without a
 * per-step upload.  wl_set_body registers the body (nprims ≤ 8; nprims = 0 removes it); wl_measure measures it at time t;
 * wl_set_remeasure(h,1) makes every step of wl_mom_step / wl_sim_step_n / wl_sim_step_until start with
 * measure!(sim, t=sum(Δt)) + update!(pois) like sim_step!(sim; remeasure=true) (src/WaterLily.jl:136-149). */
enum { WL_BODY_SPHERE = 0, WL_BODY_TORUS = 1 };
enum { WL_OP_UNION = 0, WL_OP_INTERSECT = 1, WL_OP_MINUS = 2 };
typedef struct wl_body_prim {
  int32_t kind;    /* WL_BODY_* */
  int32_t op;      /* WL_OP_* combining this primitive with everything before it (ignored for the first) */
  float center[3]; /* at t = 0, in cell units like the reference's coordinates (src/core.jl:177) */
  float R;         /* sphere radius / torus major radius */
  float r;         /* torus minor radius */
  float vel[3];    /* translation velocity of the rigid map */
} wl_body_prim;
int wl_set_body(wl_handle* h, const wl_body_prim* prims, int nprims, float eps);
int wl_measure(wl_handle* h, float t);
int wl_set_remeasure(wl_handle* h, int enabled);
/* sum(flow.Δt): the time at the END of the next step, the default `t` of measure!(sim) (src/WaterLily.jl:146). */
int wl_time_next(wl_handle* h, double* t);

/* pressure_force, viscous_force, pressure_moment(x₀), viscous_moment(x₀) of the registered body (src/Metrics.jl:111-190) at
 * t = time(flow), as one fused device reduction over p and u (Float32 per-cell vectors, Float64 sums like sum(Float64, df)).
 * out[0:3] pressure force, out[3:6] viscous force, out[6:9] pressure moment, out[9:12] viscous moment about x0 (3 floats; may be
 * NULL = origin).  In 2-D the third components are 0 and both moment components hold the scalar a₁b₂−a₂b₁ like the reference's
 * broadcast.  total_force = out[0:3] + out[3:6]. */
int wl_body_forces(wl_handle* h, const float* x0, double* out12);

/* MeanFlow (src/Metrics.jl:205-257): running time averages P, U and (uu_stats) UU = u⊗u kept ON THE DEVICE in the library's
 * layout.  wl_meanflow_init = MeanFlow(flow; t_init=time(flow), uu_stats); wl_meanflow_update = update!(meanflow, flow) with the
 * reference's weights (ε = dt/(dt + time(meanflow) + eps(T)), ε = 1 on the first update), over all cells, ghosts included;
 * wl_meanflow_reset = reset!(meanflow; t_init); wl_meanflow_copy_to_flow = copy!(flow, meanflow) (u .= U; p .= P).
 * wl_meanflow_download / _upload move P (which = 0, ΠN floats), U (1, D·ΠN) or UU (2, D·D·ΠN; component i + D·j) in the reference
 * layout — upload is what load!(meanflow) of a checkpoint does (ext/WaterLilyJLD2Ext.jl:24-38); wl_meanflow_times reads / writes
 * meanflow.t (buf = NULL queries the length). */
int wl_meanflow_init(wl_handle* h, int uu_stats);
int wl_meanflow_update(wl_handle* h);
int wl_meanflow_reset(wl_handle* h, float t_init);
int wl_meanflow_copy_to_flow(wl_handle* h);
int wl_meanflow_download(wl_handle* h, int which, float* dst, int dst_is_device);
int wl_meanflow_upload(wl_handle* h, int which, const float* src, int src_is_device);
int wl_meanflow_get_times(wl_handle* h, float* buf, int* len);
int wl_meanflow_set_times(wl_handle* h, const float* buf, int len);

/* Enumerated forcings in place of the host closures g(i,x,t) and uBC(i,x,t) (accelerate!, src/Flow.jl:64-73, and
 * BC! with a function uBC, src/core.jl:201-219), uniform in space:
 *     g_i(t) = g0_i + g1_i·t                       (body force / reference-frame acceleration)
 *     U_i(t) = cfg.uBC_i + U1_i·t + ½·U2_i·t²      (boundary velocity; its dU/dt = U1 + U2·t is added to the momentum equation)
 * Every step adds g(t₀)+dU/dt(t₀) to the predictor's and g(t₁)+dU/dt(t₁) to the corrector's right-hand side (on every cell of r,
 * ghosts included, like the reference's loop over CartesianIndices(r)) and applies BC! with U(t₁), t₁ = sum(Δt), t₀ = t₁ − Δt[end].
 * Any pointer may be NULL (= zeros); all zeros switches the forcing off. */
int wl_set_forcing(wl_handle* h, const float* g0, const float* g1, const float* U1, const float* U2);

/* The reference's LES udf as a built-in: sim_step!(sim; udf=sgs!, νₜ=smagorinsky, S, Cs, Δ) (src/util.jl:46-76) with
 * the Smagorinsky–Lilly eddy viscosity of its docstring, νₜ(I) = (Cs·Δ)²·sqrt(S[I,:,:]⋅S[I,:,:]), S(I,u) the rate-of-strain tensor at the
 * cell centre (src/Metrics.jl:42-44,140).  In both phases of every step, between conv_diff! and accelerate! (src/Flow.jl:192,207), the
 * sub-grid fluxes −νₜ(I)·∂ⱼuᵢ are added to the right-hand side over inside_u(N,j), evaluated on the advecting field of the phase (u⁰ /
 * u).  Cs·Δ = 0 switches it off.  Runs on the general kernels (the udf reads ghost cells of u); not available on z slabs. */
int wl_set_sgs(wl_handle* h, float Cs, float Delta);

/* update!(pois) (src/WaterLily.jl:148, src/MultiLevelPoisson.jl:79-86, src/Poisson.jl:47): call after uploading μ₀
 * (measure!): set_diag! on level 1 and restrictL! + set_diag! on every coarse level. */
int wl_update(wl_handle* h);

/* mom_step!(flow,pois) without udf (src/Flow.jl:156-167): appends one Δt and two Poisson iteration counts. */
int wl_mom_step(wl_handle* h);
/* `for _ in 1:nsteps sim_step!(sim; remeasure=false) end` without returning to the host in between. */
int wl_sim_step_n(wl_handle* h, int nsteps);
/* sim_step!(sim,t_end; remeasure=false, max_steps) (src/WaterLily.jl:128-135): steps while time*U/L < t_end. */
int wl_sim_step_until(wl_handle* h, double t_end, double U, double L, int64_t max_steps, int64_t* steps_taken);

/* mom_project!(a,b,w,t) (src/Flow.jl:223-232) on the current u. */
int wl_project(wl_handle* h, float w);
/* Pieces of the step exposed for operator-level parity tests:
 * conv_diff!(f,u or u⁰,σ,λ) (src/Flow.jl:38-53) — without BDIM: f receives the raw flux sum r; */
int wl_conv_diff(wl_handle* h, int from_u0);
/* CFL(a) (src/Flow.jl:234-237) — returns the value, does not push it. */
int wl_cfl(wl_handle* h, float* dt_out);

/* Standalone operator API on the handle's pressure system (x≡p, L≡μ₀, z≡σ):
 * mult!(pois,x) (src/Poisson.jl:63-69): σ = A·p, ghosts 0.
 * solver!(pois) (src/Poisson.jl:204-214, src/MultiLevelPoisson.jl:108-127): solves A·p = σ in place, returns iterations. */
int wl_pois_mult(wl_handle* h);
int wl_pois_solve(wl_handle* h, int* iters_out);
/* residual!(p) then L₂(p) (src/Poisson.jl:92-98,189). */
int wl_pois_residual(wl_handle* h, float* r2_out);
/* One smoother application on a level: kind 0 GaussSeidelRB!(it=4,ω), 1 Jacobi!(ω), 2 pcg!  (src/Poisson.jl:111-186) */
int wl_pois_smooth(wl_handle* h, int level, int kind, float omega);
/* One Vcycle!(ml;ω) (src/MultiLevelPoisson.jl:88-101). */
int wl_pois_vcycle(wl_handle* h, float omega);
int wl_num_levels(wl_handle* h, int* nlevels);
/* Ghost-padded size of a level (N[3], N[2]=1 in 2-D) and its arrays in the reference layout. */
int wl_level_dims(wl_handle* h, int level, int32_t* N);
int wl_download_level(wl_handle* h, int level, int which, float* dst);
int wl_upload_level(wl_handle* h, int level, int which, const float* src);

/* Histories mirrored to the host: flow.Δt (src/Flow.jl:127) and pois.n (src/MultiLevelPoisson.jl:66).
 * Call with buf=NULL to query the length. */
int wl_get_dt(wl_handle* h, float* buf, int* len);
int wl_set_dt(wl_handle* h, const float* buf, int len); /* restart: load!(flow.Δt) (ext/WaterLilyJLD2Ext.jl:40-50) */
int wl_get_iters(wl_handle* h, int16_t* buf, int* len);
/* The reference's solver log (src/MultiLevelPoisson.jl:111,116): rows of (iter, r2, omega). */
int wl_get_solver_log(wl_handle* h, float* buf, int* rows);
int wl_set_logging(wl_handle* h, int enabled);
/* time(a) = sum(Δt[1:end-1]) (src/Flow.jl:174) */
int wl_time(wl_handle* h, double* t);

/* Per-kernel CUDA-event timing on the launching stream (the tracing hook the reference lacks).
 * wl_get_timings writes a text table "kernel launches total_ms\n" and clears the records; buf=NULL queries the size. */
int wl_set_profiling(wl_handle* h, int enabled);
int wl_get_timings(wl_handle* h, char* buf, int* len);

/* Device self-test: number of Float32 bit patterns (of all 2^32) for which the kernels' division-free x/6 differs from IEEE x/6. */
int wl_selftest_div6(uint64_t* nbad);

int wl_sync(wl_handle* h);
/* Number of CUDA kernels this handle has launched since creation (bench.py's gpu_launches). */
int wl_launch_count(wl_handle* h, int64_t* count);
/* Launch-geometry overrides for tests and tuning (results do not depend on them): "vs_nz" = number of z chunks of f_vsmooth
 * (0 = automatic), "conv4_zchunk" = planes per block of fm_conv4. */
int wl_set_tuning(wl_handle* h, const char* key, int value);
/* 1 if the constant-coefficient (NoBody) kernel variants are active. */
int wl_is_const_coeff(wl_handle* h, int* flag);
/* cudaStream_t the handle launches on (for CUDA-event timing by the caller). */
int wl_stream(wl_handle* h, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* WL_B200_H */
