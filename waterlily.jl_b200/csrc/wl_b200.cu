// wl_b200.cu — host side of libwl_b200.so: the C ABI declared in include/wl_b200.h and the
// orchestration of mom_step! (src/Flow.jl:156-167) and solver!(::MultiLevelPoisson)
// (src/MultiLevelPoisson.jl:108-127) over the kernels in wl_kernels.cuh.
//
// There is no CPU fallback anywhere in this file: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <future>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include <nvtx3/nvToolsExt.h>
#include <sys/mman.h>
#include <thread>
#include "../../include/wl_b200.h"
#include "wl_dist.h"
#include "wl_fast.cuh"
#include "wl_conv4.cuh"
#include "wl_vsmooth.cuh"

// ---------------------------------------------------------------------------------------
// NVTX ranges around the phases of a step (mom_step!, predictor, projection, V-cycle, corrector): header-only NVTX3 — a no-op costing
// a few nanoseconds unless a tool (nsys, ncu --nvtx) has injected itself.
struct Nvtx {
  explicit Nvtx(const char* name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
};
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define TRY(call)            \
  do {                       \
    int r_ = (call);         \
    if (r_) return r_;       \
  } while (0)

static NcclApi g_nccl;
static std::map<std::tuple<int, int, int>, ncclComm_t> g_comm_cache;
// Device memory of destroyed handles is kept for the next handle of this process (by device and size): a second Simulation of the
// same shape — a restart, a parameter sweep, the bench's end-to-end leg — starts without 17 GB of cudaFree + cudaMalloc
// (0.2–0.6 s at 512³).  wl_release_pool() returns it to the driver; a failed cudaMalloc empties the pool and retries.
static std::multimap<std::pair<int, size_t>, void*> g_chunk_pool;
static std::mutex g_pool_mu;
static void* pool_take(int device, size_t size) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_chunk_pool.find({device, size});
  if (it == g_chunk_pool.end()) return nullptr;
  void* q = it->second;
  g_chunk_pool.erase(it);
  return q;
}
static void pool_give(int device, size_t size, void* q) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_chunk_pool.insert({{device, size}, q});
}
static void pool_release(int device) {  // device < 0: all devices
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto it = g_chunk_pool.begin(); it != g_chunk_pool.end();) {
    if (device < 0 || it->first.first == device) {
      cudaSetDevice(it->first.first);
      cudaFree(it->second);
      it = g_chunk_pool.erase(it);
    } else
      ++it;
  }
}
// Pinned staging ring for transfers between the device and ORDINARY host memory (what a Julia Array or a NumPy array is): the DMA
// engine moves 16 MB pieces to / from pinned buffers while host threads copy the previous pieces to / from the caller's array —
// 2–3× the rate of a pageable cudaMemcpy, which bounces through one small driver buffer on one thread.
enum { PIN_NB = 12 };
static const size_t PIN_BYTES = (size_t)16 << 20;
static char* g_pin[PIN_NB] = {nullptr};
static cudaEvent_t g_pin_ev[PIN_NB] = {nullptr};
static bool pin_ring_init() {
  if (g_pin[0]) return true;
  char* base = nullptr;
  if (cudaHostAlloc((void**)&base, PIN_BYTES * PIN_NB, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  for (int k = 0; k < PIN_NB; k++) {
    g_pin[k] = base + PIN_BYTES * k;
    cudaEventCreateWithFlags(&g_pin_ev[k], cudaEventDisableTiming);
  }
  return true;
}
static bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}  // (device, rank, nranks) → communicator, see create_impl
#define NCK(call)                                                                               \
  do {                                                                                          \
    int r_ = (call);                                                                            \
    if (r_ != 0) return fail("%s:%d NCCL %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
  } while (0)

enum { SLOT_EXIT0 = 0, SLOT_EXIT1 = 1, SLOT_RSUM = 2, SLOT_R2 = 3, SLOT_CFL = 4, SLOT_RHO = 5, SLOT_SIG = 6, SLOT_LINF = 7, SLOT_UNI = 8, SLOT_PHIMAX = 9, SLOT_CFLINT = 10, NSLOTS = 16 };

struct Level {
  Grid g;
  float *L = nullptr, *Dg = nullptr, *iD = nullptr, *x = nullptr, *eps = nullptr, *r = nullptr, *r2 = nullptr, *z = nullptr;
  unsigned char* semi = nullptr;  // general mode: semi-uniform flag per march block (Coef::semi), nullptr = not used
  float *rext = nullptr, *xext = nullptr;  // z slabs, f_vsmooth: 2×4 planes of r and 2×2 planes of x beyond the ghost planes
  int c[3] = {0, 0, 0};  // coarsening mask from the previous (finer) level
  bool ownL = false, ownz = false;
  bool fast = false;     // march kernels apply (3-D, interior x size a multiple of 4)
  bool fullc = false;    // built from the finer level by coarsening all three directions
  float Lc[3] = {1.f, 1.f, 1.f};  // uniform face coefficients (valid when the handle is in uniform mode)
  // multi-GPU: slab = this level is z-decomposed like level 0; otherwise every rank holds the whole level (coarse levels)
  bool slab = false;
  int Ng2 = 0;   // global N[2] of the level
  int zoffc = 0; // first replicated level only: global coarse plane offset of this rank's restriction (rank·nzl_fine/2)
  Coef coef(bool uni) const {
    Coef k;
    k.L = L;
    k.Dg = Dg;
    k.iD = iD;
    float s = 0.f;
    for (int d = 0; d < 3; d++) {
      k.Lc[d] = Lc[d];
      if (d < g.D) s -= Lc[d] + Lc[d];
    }
    k.Dc = s;
    k.iDc = (s == 0.f) ? s : 1.f / s;
    k.semi = uni ? nullptr : semi;
    return k;
  }
  // planes a block marches over: 16 when that still gives every SM several blocks, fewer (even, ≥2) on small grids and thin slabs
  int zchunk() const {
    const int nz = g.N[2] - 2;
    int zc = std::max(2, std::min(16, (nz + 1) / 2 * 2));
    const long xy = (long)cdiv(g.N[0] - 2, 128) * cdiv(g.N[1] - 2, FTY);
    while (zc > 2 && xy * cdiv(nz, zc) < 4 * 148) zc = std::max(2, (zc / 2 + 1) / 2 * 2);
    return zc;
  }
  dim3 fgrid() const { return dim3(cdiv(g.N[0] - 2, 128), cdiv(g.N[1] - 2, FTY), cdiv(g.N[2] - 2, zchunk())); }
  Lvl dev() const {
    Lvl l;
    l.g = g;
    l.L = L;
    l.Dg = Dg;
    l.iD = iD;
    l.x = x;
    l.eps = eps;
    l.r = r;
    l.r2 = r2;
    l.z = z;
    return l;
  }
  Box inside() const {
    Box b;
    for (int d = 0; d < 3; d++) {
      b.lo[d] = d < g.D ? 1 : 0;
      b.n[d] = d < g.D ? g.N[d] - 2 : 1;
    }
    return b;
  }
  Box all() const {
    Box b;
    for (int d = 0; d < 3; d++) {
      b.lo[d] = 0;
      b.n[d] = g.N[d];
    }
    return b;
  }
  size_t cells() const { return (size_t)g.sc; }
};

struct Chunk {
  char* base;
  size_t size, used;
};

struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
  double cells = 0;
};

struct wl_handle {
  wl_config cfg;
  bool prof = false;
  std::vector<ProfRec> prof_recs;
  std::string prof_text;
  int D;
  Grid g;
  float *u = nullptr, *u0 = nullptr, *f = nullptr, *p = nullptr, *sigma = nullptr, *V = nullptr, *mu0 = nullptr, *mu1 = nullptr;
  std::vector<Level> levels;
  RedBuf red{nullptr, nullptr, nullptr, nullptr, nullptr, 0u};
  double* h_out = nullptr;  // pinned mirror of red.out
  // mapped pinned mirror the folding thread of a reduction writes itself (RedBuf::hout/hseq): slot_tag[s] = tag of the last launch
  // that reduces into slot s (0 = none pending), see red_for / read_slot
  unsigned int tag_ctr = 0;
  unsigned int slot_tag[NSLOTS] = {0};
  bool fast_read = true;
  float* d_dthist = nullptr;
  size_t dt_cap = 0;
  float* d_scal = nullptr;  // [0]=omega, [1]=one, [2]=alpha, [3]=beta, [4]=scratch dt
  float* h_scal = nullptr;  // pinned
  std::vector<float> dt;    // host mirror of flow.Δt
  double tsum = 0.0;        // Σ dt[0 .. tsum_n) in double (wl_time)
  size_t tsum_n = 0;
  size_t dt_dev_len = 0;    // entries valid on the device
  std::vector<int16_t> iters;
  std::vector<float> log;  // rows of (iter, rinf, r2, omega)
  bool logging = false;
  Dist dist;
  std::vector<Chunk> chunks;
  // peer-to-peer halo exchange over NVLink (CUDA IPC): mappings of the neighbours' chunks and a mailbox of flags
  bool p2p = false;
  std::vector<char*> peer_base[2];  // [0] lower neighbour, [1] upper neighbour
  int* mbox = nullptr;              // ready[2], arrived[2], block counter, error
  int halo_seq = 0;
  // second exchange lane (own stream, own mailbox words mbox+16…, own sequence): pushes whose data is final long before its reader runs
  // — the 4+1-plane halo of r that f_vsmooth reads is final after the level's Jacobi! — overlap the coarse part of the V-cycle
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_pre[4] = {nullptr, nullptr, nullptr, nullptr};  // per level: the prefetch push of that level's r halo is complete
  bool pre_done[4] = {false, false, false, false};
  int halo_seq2 = 0;
  bool prefetch = true;
  bool skip_ready = true;
  bool coop_pdl = true;  // k_small_levels: cooperative launch with the PDL attribute as well (cleared if the driver refuses)
  bool pdl = true;  // programmatic dependent launch of the uniform-mode step's kernels (WL_FLAG_NO_PDL)
  // own all-reduce of the solver's / CFL's scalars over peer memory (k_allreduce): every rank's mailbox mapped on every rank
  double* armb = nullptr;
  ArPeers ar_peers{};
  std::vector<void*> ar_opened;  // IPC mappings opened for it beyond the neighbours'
  bool ar_on = false;
  long long ar_seq = 0, bc_seq = 0;
  std::vector<char> ipc_all;                 // every rank's chunk handles (setup_p2p), for mappings opened later
  std::vector<std::vector<char*>> any_base;  // [rank][chunk] → base of that chunk as mapped here (null: not mapped yet)
  bool slot_ar[NSLOTS] = {false};  // the slot's latest value came from k_allreduce (its tag is in slot_tag: the host may poll it)
  float* uext = nullptr;  // second halo planes of the velocity beyond open z faces: [side][component][plane]
  int perz_global = 0;
  int slab_min_planes = 8;
  double slab_min_cells = 3.0e6;  // 512³ on 2/4/8 GPUs: levels 514³ and 258³ are z slabs, 130³ and below replicated
  // persistent coarse-level kernel: levels >= small_from run inside one cooperative launch per V-cycle (0 = disabled)
  int small_from = 0;
  int small_grid = 0;
  int* d_flags = nullptr;  // [0]: a velocity field holds a non-finite value (or |u| > 1e37); [2]: range word of the current u, [3]: its ticket (range_note)
  float* stage = nullptr;  // dense staging buffer of one component for host transfers
  size_t stage_cap = 0;
  // MeanFlow (src/Metrics.jl:205-257): device-resident running averages
  float *mfP = nullptr, *mfU = nullptr, *mfUU = nullptr;
  std::vector<float> mft;
  bool mf_uu = false;
  RedBuf force_red{nullptr, nullptr, nullptr};  // wl_body_forces: 12 sums per block
  size_t force_cap = 0;
  BodySet body;            // parametrised body registered with wl_set_body (np = 0: none)
  float body_eps = 1.f;
  bool remeasure = false;  // every step starts with measure!(sim, t=sum(Δt)) + update!(pois)
  // enumerated forcings in place of the closures g(i,x,t), uBC(i,x,t): g_i(t) = g0_i + g1_i·t, U_i(t) = uBC_i + U1_i·t + ½·U2_i·t²
  bool forcing = false;
  float fg0[3] = {0, 0, 0}, fg1[3] = {0, 0, 0}, fU1[3] = {0, 0, 0}, fU2[3] = {0, 0, 0};
  Force fc{0, {0.f, 0.f, 0.f}};  // what the flux kernels of the current stage add to r
  float ubc_now[3] = {0, 0, 0};  // uBC at the time the step's BC! calls are made at (t₁)
  // built-in udf: sgs! with νₜ = smagorinsky (src/util.jl:46-76): (Cs·Δ)², the νₜ scratch field (ghost cells stay 0)
  bool sgs_on = false;
  float sgs_c2 = 0.f;
  float* nut = nullptr;
  bool range_checked = false;  // the current u was range-checked by the kernel that wrote it (range_note, wl_common.cuh)
  SmallOp* d_ops = nullptr;
  SmallOp* h_ops = nullptr;  // pinned
  std::vector<SmallOp> ops;  // coarser levels with fewer planes per rank are replicated on every rank (exchange latency > redundant work)
  bool uni = false;  // uniform-coefficient kernels active (no body, fully periodic)
  bool fused_gs = true;
  bool attr_vs = false, attr_c4[3] = {false, false, false};  // dynamic shared-memory opt-in done on this handle's device
  unsigned char* nobody = nullptr;  // general mode: per k_bdim2 block, no body inside (k_nobody_flags); valid after wl_update until μ₀/μ₁/V change
  bool nobody_valid = false;
  // V, μ₀ or μ₁ were written through the ABI since the hierarchy was built: every entry point that solves or steps rebuilds it
  // first (update!(pois), src/MultiLevelPoisson.jl:79-86), so a host that forgets wl_update cannot step with stale D/iD/coarse L
  bool pois_dirty = false;
  bool tiny_on = true;     // uniform mode: levels ≤ 8192 cells on one block in shared memory (k_tiny_uni); WL_FLAG_NO_TINY: all in k_small_levels
  // general mode: levels ≥ tg_from run in k_tiny_gen (one block, dense copies in shared memory); its operation list and arguments
  // are built once (the levels' arrays do not move: nothing outside that kernel swaps r/r2 of those levels)
  int tg_from = 0;
  bool tg_ready = false, attr_tg = false;
  SmallOp* d_tgops = nullptr;
  int tg_nops = 0;
  TinyGenArgs tg_args;
  int tiny_from = 0;       // first level k_tiny_uni can take (0 = none): fully coarsened chain of 3-D periodic levels of ≤ 8192 cells
  bool attr_tiny = false;
  bool semi_on = true;     // general mode: semi-uniform march blocks (WL_SEMI=0: always read L)
  bool jacobi2 = true;     // uniform mode, level 1: f_jacobi_uni2 (WL_JACOBI2=0: f_jacobi<true>)
  bool divres_uni = true;  // uniform mode: f_divres_uni (WL_DIVRES_UNI=0: f_div_residual<true>)
  bool fuse_cfl = true;    // uniform mode: f_correct_cfl (WL_FUSE_CFL=0: f_correct + f_cfl)
  bool vsmooth = true;     // uniform mode: f_vsmooth fuses prolongation, GaussSeidelRB! and both increments (WL_VSMOOTH=0: separate launches)
  bool conv4 = true;       // uniform mode: fm_conv4 (WL_CONV4=0 falls back to fm_conv)
  int conv4_zchunk = 32;
  int vs_nz = 0;           // WL_VS_NZ: force the number of z chunks of f_vsmooth
  // Uniform mode on one GPU never reads the periodic ghost cells of u (every reader wraps its indices), so the BC! launches of
  // mom_step! are deferred until something outside the step loop looks at u (flush_ghosts).
  bool ghosts_dirty = false;
  double prof_cells = 0;  // ghost-padded cells of the level the next launches run on (profiling: algorithmic bytes per kernel)
  cudaStream_t st = nullptr;
  int64_t launches = 0;
  double tol;
  int itmx;
  std::vector<void*> allocs;
};

// Optional per-kernel CUDA-event timing on the launching stream (wl_set_profiling / wl_get_timings).
static void prof_begin(wl_handle* h, const char* name) {
  if (!h->prof) return;
  ProfRec r;
  r.name = name;
  r.cells = h->prof_cells;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, h->st);
  h->prof_recs.push_back(r);
}
static void prof_end(wl_handle* h) {
  if (!h->prof) return;
  cudaEventRecord(h->prof_recs.back().b, h->st);
}

struct ProfLevel {  // launches inside this scope are accounted to level `l`
  wl_handle* h;
  double saved;
  ProfLevel(wl_handle* h_, const Level& l) : h(h_), saved(h_->prof_cells) { h->prof_cells = (double)l.g.N[0] * l.g.N[1] * l.g.N[2]; }
  ~ProfLevel() { h->prof_cells = saved; }
};

static Grid make_grid(int D, const int* N, const int* per) {
  Grid g;
  g.D = D;
  for (int d = 0; d < 3; d++) {
    g.N[d] = d < D ? N[d] : 1;
    g.per[d] = d < D ? (per[d] != 0) : 0;
  }
  g.xo = 31;
  g.px = ((g.N[0] + g.xo + 3 + 31) / 32) * 32;
  g.s[0] = 1;
  g.s[1] = g.px;
  g.s[2] = (i64)g.px * g.N[1];
  g.sc = (i64)g.px * g.N[1] * g.N[2];
  g.zopen[0] = g.zopen[1] = g.zstale[0] = g.zstale[1] = 0;
  g.zoff = 0;
  return g;
}

static inline dim3 blk(int D) { return D == 3 ? dim3(32, 4, 4) : dim3(32, 8, 1); }
static inline dim3 grd(const Box& b, dim3 t) { return dim3(cdiv(b.n[0], t.x), cdiv(b.n[1], t.y), cdiv(b.n[2], t.z)); }
static inline size_t nblocks(const Box& b, dim3 t) {
  dim3 g = grd(b, t);
  return (size_t)g.x * g.y * g.z;
}

// Every launch of the step carries the programmatic-stream-serialization attribute, and every kernel of the library starts with
// pdl_wait() (wl_common.cuh): kernel n+1 is scheduled into the tail of kernel n and waits there for its completion.
template <typename... KArgs, typename... Args>
static cudaError_t pdl_launch(wl_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = h->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#define LAUNCH(h, kern, grid, block, ...)                \
  do {                                                   \
    prof_begin(h, #kern);                                \
    pdl_launch(h, kern, (grid), (block), 0, __VA_ARGS__); \
    prof_end(h);                                         \
    (h)->launches++;                                     \
  } while (0)
// dispatch on the spatial dimension
#define LAUNCH_D(h, kern, grid, block, ...)                       \
  do {                                                            \
    if ((h)->D == 3)                                              \
      LAUNCH(h, kern<3>, grid, block, __VA_ARGS__);               \
    else                                                          \
      LAUNCH(h, kern<2>, grid, block, __VA_ARGS__);               \
  } while (0)

#define LAUNCH_PDL(h, kern, grid, block, smem, ...)                 \
  do {                                                              \
    prof_begin(h, #kern);                                           \
    pdl_launch(h, kern, (grid), (block), (smem), __VA_ARGS__);      \
    prof_end(h);                                                    \
    (h)->launches++;                                                \
  } while (0)

// Field memory comes from a few big chunks (bump allocation, identical on every rank), so that a neighbouring rank can address
// any field of this rank through one CUDA-IPC mapping per chunk: peer pointer = peer chunk base + (local pointer − local chunk base).
static int dalloc(wl_handle* h, float** p, size_t nfloats) {
  const size_t bytes = (nfloats * sizeof(float) + 511) / 512 * 512;
  if (h->chunks.empty() || h->chunks.back().used + bytes > h->chunks.back().size) {
    Chunk c;
    c.size = std::max(bytes, (size_t)256 << 20);
    c.used = 0;
    void* q = pool_take(h->cfg.device, c.size);
    if (!q) {
      cudaError_t e = cudaMalloc(&q, c.size);
      if (e != cudaSuccess) {  // give the pooled memory of earlier handles back and try once more
        cudaGetLastError();
        pool_release(h->cfg.device);
        e = cudaMalloc(&q, c.size);
      }
      if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes): %s", c.size, cudaGetErrorString(e));
    }
    c.base = (char*)q;
    h->chunks.push_back(c);
  }
  Chunk& c = h->chunks.back();
  void* q = c.base + c.used;
  c.used += bytes;
  cudaError_t e = cudaMemsetAsync(q, 0, bytes, h->st);
  if (e != cudaSuccess) return fail("cudaMemset: %s", cudaGetErrorString(e));
  *p = (float*)q;
  return 0;
}

static const float* dtp(wl_handle* h) { return h->d_dthist + (h->dt_dev_len - 1); }

// ---- z-slab halo exchange and collectives (NCCL over NVLink) ---------------------------------
// Ghost planes of a field on a slab level: my top interior plane → the upper neighbour's lower ghost plane, my bottom interior
// plane → the lower neighbour's upper ghost plane.  Sends/receives to one peer are posted in matching order (P = 2 periodic).
// Peer pointer of a local field address on the lower (0) / upper (1) neighbour.
static float* peer_ptr(const wl_handle* h, int side, const float* local) {
  const char* q = (const char*)local;
  for (size_t k = 0; k < h->chunks.size(); k++) {
    const Chunk& c = h->chunks[k];
    if (q >= c.base && q < c.base + c.size) return (float*)(h->peer_base[side][k] + (q - c.base));
  }
  return nullptr;
}
// P2P exchange of up to 16 planes: planes[i] = {local source plane, side (0 lower / 1 upper), local address of the DESTINATION
// plane as it is laid out on the neighbour (same layout on every rank)}.
struct PlaneMove {
  const float* src;
  int side;
  const float* dst_local;
};
// `lane` 1: on the side stream (after everything enqueued on the main stream so far), see wl_handle::st2
// `war_safe`: at least one other exchange lies between the neighbours' last read of the destination planes and this push (true for
// every exchange of the uniform-mode step: each set of ghost planes is rewritten once per V-cycle / projection) → k_halo_push
// skips its "ready" round trip.  Only honoured in uniform mode: the general smoother refreshes ϵ's ghosts after every half-sweep.
static int p2p_push(wl_handle* h, const Grid& g, const PlaneMove* mv, int n, bool carries_u = false, int lane = 0, bool war_safe = false) {
  HaloSegs segs;
  memset(&segs, 0, sizeof segs);
  const Dist& d = h->dist;
  int m = 0;
  for (int i = 0; i < n; i++) {
    const int peer = mv[i].side ? d.up : d.down;
    if (peer < 0) continue;
    segs.src[m] = mv[i].src;
    segs.dst[m] = peer_ptr(h, mv[i].side, mv[i].dst_local);
    if (!segs.dst[m]) return fail("halo plane outside the IPC-mapped chunks");
    m++;
  }
  segs.nseg = m;
  const int cnt4 = (int)(g.s[2] / 4);
  int* const mb = h->mbox + (lane ? 16 : 0);
  int* plo = d.down >= 0 ? (int*)peer_ptr(h, 0, (const float*)mb) : nullptr;
  int* phi = d.up >= 0 ? (int*)peer_ptr(h, 1, (const float*)mb) : nullptr;
  const int seq = lane ? ++h->halo_seq2 : ++h->halo_seq;
  const long long total4 = (long long)m * cnt4;
  const int nb = (int)std::max<long long>(1, std::min<long long>(lane ? 64 : 128, (total4 + 2047) / 2048));
  const int* myf = carries_u ? h->d_flags : nullptr;
  int* flo = carries_u && d.down >= 0 ? (int*)peer_ptr(h, 0, (const float*)h->d_flags) : nullptr;
  int* fhi = carries_u && d.up >= 0 ? (int*)peer_ptr(h, 1, (const float*)h->d_flags) : nullptr;
  const int no_ready = (war_safe && h->uni && h->skip_ready) ? 1 : 0;
  if (lane) {
    CK(cudaEventRecord(h->ev_fork, h->st));
    CK(cudaStreamWaitEvent(h->st2, h->ev_fork, 0));
    k_halo_push<<<nb, 256, 0, h->st2>>>(segs, cnt4, seq, mb, plo, phi, 60000000000LL, myf, flo, fhi, no_ready);
    h->launches++;
    return 0;
  }
  prof_begin(h, "halo_exchange_p2p");
  pdl_launch(h, k_halo_push, dim3(nb), dim3(256), 0, segs, cnt4, seq, mb, plo, phi, 60000000000LL, myf, flo, fhi, no_ready);
  prof_end(h);
  h->launches++;
  return 0;
}

static int exch(wl_handle* h, const Level& l, float* a, int ncomp, bool war_safe = false) {
  if (!h->dist.on() || !l.slab) return 0;
  const Grid& g = l.g;
  const size_t cnt = (size_t)g.s[2];
  const Dist& d = h->dist;
  if (h->p2p && ncomp <= 8) {
    PlaneMove mv[16];
    int n = 0;
    for (int c = 0; c < ncomp; c++) {
      float* b = a + (size_t)c * g.sc;
      mv[n++] = {b + g.s[2] * (g.N[2] - 2), 1, b};                      // my top interior plane → upper neighbour's lower ghost
      mv[n++] = {b + g.s[2] * 1, 0, b + g.s[2] * (g.N[2] - 1)};         // my bottom interior plane → lower neighbour's upper ghost
    }
    return p2p_push(h, g, mv, n, false, 0, war_safe);
  }
  if (h->p2p) {  // more than 8 components (μ₁): in two batches
    TRY(exch(h, l, a, 8));
    return exch(h, l, a + (size_t)8 * g.sc, ncomp - 8);
  }
  prof_begin(h, "halo_exchange");
  NCK(g_nccl.GroupStart());
  for (int c = 0; c < ncomp; c++) {
    float* b = a + (size_t)c * g.sc;
    if (d.up >= 0) NCK(g_nccl.Send(b + g.s[2] * (g.N[2] - 2), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
    if (d.down >= 0) NCK(g_nccl.Send(b + g.s[2] * 1, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
    if (d.down >= 0) NCK(g_nccl.Recv(b, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
    if (d.up >= 0) NCK(g_nccl.Recv(b + g.s[2] * (g.N[2] - 1), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
  }
  NCK(g_nccl.GroupEnd());
  prof_end(h);
  return 0;
}
// Two scalar fields in one NCCL group (r and x after increment!)
static int exch2(wl_handle* h, const Level& l, float* a0, float* a1, bool war_safe = false) {
  if (!h->dist.on() || !l.slab) return 0;
  const Grid& g = l.g;
  const size_t cnt = (size_t)g.s[2];
  const Dist& d = h->dist;
  float* f[2] = {a0, a1};
  if (h->p2p) {
    PlaneMove mv[4];
    int n = 0;
    for (int c = 0; c < 2; c++) {
      float* b = f[c];
      mv[n++] = {b + g.s[2] * (g.N[2] - 2), 1, b};
      mv[n++] = {b + g.s[2] * 1, 0, b + g.s[2] * (g.N[2] - 1)};
    }
    return p2p_push(h, g, mv, n, false, 0, war_safe);
  }
  prof_begin(h, "halo_exchange");
  NCK(g_nccl.GroupStart());
  for (int c = 0; c < 2; c++) {
    float* b = f[c];
    if (d.up >= 0) NCK(g_nccl.Send(b + g.s[2] * (g.N[2] - 2), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
    if (d.down >= 0) NCK(g_nccl.Send(b + g.s[2] * 1, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
    if (d.down >= 0) NCK(g_nccl.Recv(b, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
    if (d.up >= 0) NCK(g_nccl.Recv(b + g.s[2] * (g.N[2] - 1), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
  }
  NCK(g_nccl.GroupEnd());
  prof_end(h);
  return 0;
}
// Velocity halo: two planes per side (QUICK reads I-2δ … I+δ, src/Flow.jl:8); the second plane lands in h->uext.
// `p`: the pressure's ghost planes ride along (after a projection)
static int exch_u(wl_handle* h, float* u, float* p = nullptr) {
  if (!h->dist.on()) return 0;
  const Grid& g = h->g;
  const size_t cnt = (size_t)g.s[2];
  const Dist& d = h->dist;
  if (p && !h->p2p) TRY(exch(h, h->levels[0], p, 1));
  if (h->p2p) {
    PlaneMove mv[14];
    int n = 0;
    if (p) {
      mv[n++] = {p + g.s[2] * (g.N[2] - 2), 1, p};
      mv[n++] = {p + g.s[2] * 1, 0, p + g.s[2] * (g.N[2] - 1)};
    }
    for (int c = 0; c < 3; c++) {
      float* b = u + (size_t)c * g.sc;
      float* elo = h->uext + (size_t)c * g.s[2];
      float* ehi = h->uext + (size_t)(3 + c) * g.s[2];
      mv[n++] = {b + g.s[2] * (g.N[2] - 2), 1, b};
      mv[n++] = {b + g.s[2] * (g.N[2] - 3), 1, elo};
      mv[n++] = {b + g.s[2] * 1, 0, b + g.s[2] * (g.N[2] - 1)};
      mv[n++] = {b + g.s[2] * 2, 0, ehi};
    }
    return p2p_push(h, g, mv, n, true, 0, true);
  }
  prof_begin(h, "halo_exchange_u");
  NCK(g_nccl.GroupStart());
  for (int c = 0; c < 3; c++) {
    float* b = u + (size_t)c * g.sc;
    float* elo = h->uext + (size_t)c * g.s[2];
    float* ehi = h->uext + (size_t)(3 + c) * g.s[2];
    if (d.up >= 0) {
      NCK(g_nccl.Send(b + g.s[2] * (g.N[2] - 2), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
      NCK(g_nccl.Send(b + g.s[2] * (g.N[2] - 3), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
    }
    if (d.down >= 0) {
      NCK(g_nccl.Send(b + g.s[2] * 1, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
      NCK(g_nccl.Send(b + g.s[2] * 2, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
    }
    if (d.down >= 0) {
      NCK(g_nccl.Recv(b, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
      NCK(g_nccl.Recv(elo, cnt, WL_NCCL_FLOAT, d.down, d.comm, h->st));
    }
    if (d.up >= 0) {
      NCK(g_nccl.Recv(b + g.s[2] * (g.N[2] - 1), cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
      NCK(g_nccl.Recv(ehi, cnt, WL_NCCL_FLOAT, d.up, d.comm, h->st));
    }
  }
  NCK(g_nccl.GroupEnd());
  prof_end(h);
  return 0;
}
// Uniform mode, between the flux kernel and the projection's velocity correction: the only ghost values of the intermediate velocity
// anyone reads are u_z one plane above the slab (div in f_divres_uni, flux_out in f_correct_cfl) — the corrected field is exchanged in
// full (two planes per side, three components, and p) after the correction.  One plane, one direction, instead of twelve.
static int exch_uz_up(wl_handle* h, float* u) {
  if (!h->dist.on()) return 0;
  if (!h->p2p) return exch_u(h, u);
  const Grid& g = h->g;
  float* b = u + (size_t)2 * g.sc;
  PlaneMove mv[1] = {{b + g.s[2] * 1, 0, b + g.s[2] * (g.N[2] - 1)}};  // my bottom interior plane → the lower neighbour's upper ghost plane
  return p2p_push(h, g, mv, 1, false, 0, true);
}
// In-place all-reduce of one reduction slot (double) across the ranks, on the compute stream.
static int allreduce_slot(wl_handle* h, int slot, int op, int count = 1) {  // `count` adjacent slots in one call
  if (!h->dist.on()) return 0;
  if (h->ar_on && count <= 2) {
    if (++h->tag_ctr == 0) h->tag_ctr = 1;
    const unsigned int tag = h->tag_ctr;
    h->slot_tag[slot] = tag;
    h->slot_ar[slot] = true;
    prof_begin(h, "allreduce_p2p");
    pdl_launch(h, k_allreduce, dim3(1), dim3(32), 0, h->ar_peers, h->dist.P, h->dist.rank, ++h->ar_seq, op == WL_NCCL_SUM ? RED_SUM : RED_MAX, count, h->red.out + slot,
                                     h->red.hout + slot, h->red.hseq + slot, tag, h->mbox + 40, 60000000000LL);
    prof_end(h);
    h->launches++;
    return 0;
  }
  prof_begin(h, "allreduce");
  NCK(g_nccl.AllReduce(h->red.out + slot, h->red.out + slot, count, WL_NCCL_DOUBLE, op, h->dist.comm, h->st));
  prof_end(h);
  return 0;
}
// Gather the interior planes of a replicated level's field from the ranks that each restricted their own slab into it.
// Base of chunk k of rank q as mapped into this process (opened on first use; the neighbours' mappings are reused).
static float* any_peer_ptr(wl_handle* h, int q, const float* local) {
  const Dist& d = h->dist;
  const char* lp = (const char*)local;
  for (size_t k = 0; k < h->chunks.size(); k++) {
    const Chunk& c = h->chunks[k];
    if (lp < c.base || lp >= c.base + c.size) continue;
    if (q == d.rank) return (float*)local;
    if (h->any_base.size() != (size_t)d.P) h->any_base.assign(d.P, std::vector<char*>());
    std::vector<char*>& tab = h->any_base[q];
    if (tab.size() < h->chunks.size()) tab.resize(h->chunks.size(), nullptr);
    if (!tab[k]) {
      if (q == d.down && k < h->peer_base[0].size() && h->peer_base[0][k])
        tab[k] = h->peer_base[0][k];
      else if (q == d.up && k < h->peer_base[1].size() && h->peer_base[1][k])
        tab[k] = h->peer_base[1][k];
      else {
        const size_t rec = sizeof(cudaIpcMemHandle_t);
        const size_t per = h->ipc_all.size() / d.P;
        if ((k + 1) * rec > per) {
          fail("chunk %zu was allocated after the IPC handles were exchanged", k);
          return nullptr;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, h->ipc_all.data() + per * q + k * rec, rec);
        void* m = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&m, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          fail("cudaIpcOpenMemHandle (rank %d chunk %zu): %s", q, k, cudaGetErrorString(e));
          return nullptr;
        }
        h->ar_opened.push_back(m);
        tab[k] = (char*)m;
      }
    }
    return (float*)(tab[k] + (lp - c.base));
  }
  fail("address outside the IPC-mapped chunks");
  return nullptr;
}
// `p2p_ok`: the call sits inside the V-cycle, where an all-reduce separates two gathers of the same array (the peer-memory version
// writes into the other ranks' copies without asking whether they are done reading the previous contents)
static int allgather_planes(wl_handle* h, const Level& lc, float* a, int planes_per_rank, bool p2p_ok = false) {
  if (!h->dist.on()) return 0;
  const Grid& g = lc.g;
  const size_t cnt = (size_t)g.s[2] * planes_per_rank;
  float* base = a + g.s[2];  // plane 1
  if (h->ar_on && p2p_ok && cnt % 4 == 0) {
    const Dist& d = h->dist;
    float* mine = base + cnt * d.rank;
    BcastDst dst;
    memset(&dst, 0, sizeof dst);
    for (int q = 0; q < d.P; q++) {
      if (q == d.rank) continue;
      float* pq = any_peer_ptr(h, q, mine);
      if (!pq) return 1;
      dst.p[q] = reinterpret_cast<float4*>(pq);
    }
    ArPeers bar = h->ar_peers;
    for (int q = 0; q < d.P; q++) bar.p[q] += 64;  // the barrier's region of the mailbox
    const long long n4 = (long long)(cnt / 4);
    const int nb = (int)std::max<long long>(1, std::min<long long>(96, (n4 + 1023) / 1024));
    prof_begin(h, "allgather_p2p");
    pdl_launch(h, k_bcast_planes, dim3(nb), dim3(256), 0, reinterpret_cast<const float4*>(mine), dst, n4, bar, d.P, d.rank, ++h->bc_seq, h->mbox + 44, h->mbox + 40, 60000000000LL);
    prof_end(h);
    h->launches++;
    return 0;
  }
  prof_begin(h, "allgather");
  NCK(g_nccl.AllGather(base + cnt * h->dist.rank, base, cnt, WL_NCCL_FLOAT, h->dist.comm, h->st));
  prof_end(h);
  return 0;
}

// ---- boundary conditions ------------------------------------------------------------------
static void launch_bc_vec(wl_handle* h, const Grid& g, float* a, const float* U, int saveexit, const float* keep_src) {
  // planes are at most max(N)² cells; one launch covers all 3·D planes (blockIdx.z)
  int m0 = std::max(g.N[0], g.N[1]), m1 = g.D == 3 ? std::max(g.N[1], g.N[2]) : 1;
  if (g.D == 3) m0 = std::max(m0, g.N[0]);
  dim3 b = g.D == 3 ? dim3(32, 8, 1) : dim3(64, 1, 1);
  dim3 gr(cdiv(m0, b.x), cdiv(m1, b.y), 3 * g.D);
  if (g.D == 3)
    LAUNCH(h, k_bc_vec<3>, gr, b, g, a, keep_src, U[0], U[1], U[2], saveexit);
  else
    LAUNCH(h, k_bc_vec<2>, gr, b, g, a, keep_src, U[0], U[1], U[2], saveexit);
}
static void launch_perbc(wl_handle* h, const Grid& g, float* a) {
  if (!(g.per[0] || g.per[1] || g.per[2])) return;
  int m0 = std::max(g.N[0], g.N[1]), m1 = g.D == 3 ? std::max(g.N[1], g.N[2]) : 1;
  dim3 b = g.D == 3 ? dim3(32, 8, 1) : dim3(64, 1, 1);
  dim3 gr(cdiv(m0, b.x), cdiv(m1, b.y), 2 * g.D);
  if (g.D == 3)
    LAUNCH(h, k_perbc<3>, gr, b, g, a);
  else
    LAUNCH(h, k_perbc<2>, gr, b, g, a);
}
static int launch_exitbc(wl_handle* h, float* u, const float* u0, float dt_scale) {
  const Grid& g = h->g;
  dim3 b = g.D == 3 ? dim3(16, 16, 1) : dim3(256, 1, 1);
  dim3 gr(cdiv(g.N[1] - 2, b.x), g.D == 3 ? cdiv(g.N[2] - 2, b.y) : 1, 1);
  // length(exitR): the whole y-z face, over all slabs
  const float len = g.D == 3 ? (float)((i64)(g.N[1] - 2) * (h->cfg.n[2])) : (float)(g.N[1] - 2);
  for (int stage = 0; stage < 3; stage++) {
    LAUNCH_D(h, k_exitbc, gr, b, g, u, u0, dtp(h), dt_scale, h->red, SLOT_EXIT0, stage, len);
    if (stage < 2) TRY(allreduce_slot(h, SLOT_EXIT0 + stage, WL_NCCL_SUM));
  }
  return 0;
}

// ---- Poisson hierarchy ---------------------------------------------------------------------
// (z slabs too: the z ghost planes come from the halo exchange, and BC! of the x and y ghosts — also those inside the z ghost planes — is local)
static inline bool lazy_bc(const wl_handle* h) { return h->uni && h->D == 3 && !h->cfg.exitBC; }
// BC!(u) of mom_step! (src/Flow.jl:194,209,230): deferred in uniform mode, see wl_handle::ghosts_dirty
static void step_bc(wl_handle* h, const float* keep) {
  if (lazy_bc(h)) {
    h->ghosts_dirty = true;
    return;
  }
  launch_bc_vec(h, h->g, h->u, h->forcing ? h->ubc_now : h->cfg.uBC, h->cfg.exitBC, keep);
}
static void flush_ghosts(wl_handle* h) {
  if (!h->ghosts_dirty) return;
  launch_bc_vec(h, h->g, h->u, h->cfg.uBC, h->cfg.exitBC, h->u);
  launch_bc_vec(h, h->g, h->u0, h->cfg.uBC, h->cfg.exitBC, h->u0);
  h->ghosts_dirty = false;
}
static inline bool divisible(int N) { return N % 2 == 0 && N > 4; }  // src/MultiLevelPoisson.jl:52

// Local grid of a z-slab level: nz/P interior planes, ghost planes fed by the neighbours (zopen), stale marks on the global
// periodic boundary, no local wrap in z.
static Grid slab_grid(const wl_handle* h, const Grid& gg) {
  const Dist& d = h->dist;
  const int nzl = (gg.N[2] - 2) / d.P;
  int N[3] = {gg.N[0], gg.N[1], nzl + 2};
  int per[3] = {gg.per[0], gg.per[1], 0};
  Grid g = make_grid(3, N, per);
  g.zopen[0] = d.down >= 0;
  g.zopen[1] = d.up >= 0;
  g.zstale[0] = h->perz_global && d.rank == 0;
  g.zstale[1] = h->perz_global && d.rank == d.P - 1;
  g.zoff = d.rank * nzl;
  return g;
}

static int build_levels(wl_handle* h) {
  // the hierarchy of the GLOBAL domain (src/MultiLevelPoisson.jl:68-72) …
  std::vector<Level> gl;
  {
    Level l0;
    int N[3], per[3];
    for (int d = 0; d < 3; d++) {
      N[d] = d < h->D ? h->cfg.n[d] + 2 : 1;
      per[d] = h->cfg.perdir[d];
    }
    l0.g = make_grid(h->D, N, per);
    gl.push_back(l0);
  }
  if (h->cfg.pois_kind == WL_POIS_MULTILEVEL) {
    const int maxlevels = 10;
    while ((int)gl.size() <= maxlevels) {
      const Grid& gf = gl.back().g;
      bool any = false;
      int N[3];
      Level lc;
      for (int d = 0; d < 3; d++) {
        const bool c = d < gf.D && divisible(gf.N[d]);
        lc.c[d] = c;
        N[d] = c ? 1 + gf.N[d] / 2 : gf.N[d];
        any |= c;
      }
      if (!any) break;
      lc.g = make_grid(gf.D, N, gf.per);
      lc.ownL = lc.ownz = true;
      lc.fullc = gf.D == 3 && lc.c[0] && lc.c[1] && lc.c[2];
      const Level& fl = gl.back();
      for (int d = 0; d < gf.D; d++) {  // restrictL of a uniform field: Σ over the transverse fine faces, /2 if the normal is coarsened
        float v = fl.Lc[d];
        for (int j = 0; j < gf.D; j++)
          if (j != d && lc.c[j]) v *= 2.f;
        if (lc.c[d]) v /= 2.f;
        lc.Lc[d] = v;
      }
      gl.push_back(lc);
    }
    if (gl.size() <= 2) return fail("MultiLevelPoisson requires size=a2^n, where n>2");
  }
  // … decomposed into z slabs while a rank keeps at least 4 (and an even number of) planes; coarser levels are replicated
  const int P = h->dist.P;
  for (size_t i = 0; i < gl.size(); i++) {
    Level l = gl[i];
    l.Ng2 = l.g.N[2];
    if (P > 1) {
      const int nz = l.g.N[2] - 2;
      const bool prev = i == 0 || h->levels[i - 1].slab;
      // a level is decomposed while its slabs keep enough planes AND it is big enough for 1/P of its kernels to outweigh the
      // exchanges after them (≈25 µs each); smaller levels are replicated on every rank
      const double gcells = (double)l.g.N[0] * l.g.N[1] * l.g.N[2];
      l.slab = prev && nz % P == 0 && nz / P >= (i == 0 ? 4 : h->slab_min_planes) && (nz / P) % 2 == 0 && (i == 0 || (l.c[2] && gcells > h->slab_min_cells));
      if (i == 0 && !l.slab) return fail("z-slab decomposition needs dims[3]=%d divisible by %d ranks with an even number (>=4) of planes each", nz, P);
      if (l.slab) {
        l.g = slab_grid(h, gl[i].g);
      } else if (i > 0 && h->levels[i - 1].slab) {
        if (!l.fullc) return fail("semi-coarsening across the slab/replicated level transition is not supported");
        l.zoffc = h->dist.rank * ((h->levels[i - 1].g.N[2] - 2) / 2);
      }
    }
    if (i == 0) {
      l.L = h->mu0;
      l.z = h->sigma;
      l.g = h->g;
      l.g.D = h->D;
    }
    h->levels.push_back(l);
  }
  for (size_t i = 0; i < h->levels.size(); i++) {
    Level& l = h->levels[i];
    const size_t n = l.cells();
    l.fast = l.g.D == 3 && (l.g.N[0] - 2) % 4 == 0 && l.g.N[2] > 3;
    if (l.slab && !l.fast) return fail("z-slab levels need the march kernels: interior x size must be a multiple of 4 (level %zu)", i);
    if (l.ownL) TRY(dalloc(h, &l.L, n * l.g.D));
    if (l.ownz) TRY(dalloc(h, &l.z, n));
    TRY(dalloc(h, &l.Dg, n));
    TRY(dalloc(h, &l.iD, n));
    TRY(dalloc(h, &l.x, n));
    TRY(dalloc(h, &l.eps, n));
    TRY(dalloc(h, &l.r, n));
    TRY(dalloc(h, &l.r2, n));
    if (l.fast && h->semi_on) {
      const dim3 fg = l.fgrid();
      float* q = nullptr;
      TRY(dalloc(h, &q, ((size_t)fg.x * fg.y * fg.z + 3) / 4 + 1));
      l.semi = (unsigned char*)q;
    }
    if (l.slab) {
      TRY(dalloc(h, &l.rext, (size_t)8 * l.g.s[2]));
      TRY(dalloc(h, &l.xext, (size_t)4 * l.g.s[2]));
    }
  }
  return 0;
}

static int update_levels(wl_handle* h) {  // update!(ml)  src/MultiLevelPoisson.jl:79-86
  h->pois_dirty = false;
  const float zero[3] = {0, 0, 0};
  for (size_t i = 0; i < h->levels.size(); i++) {
    Level& l = h->levels[i];
    dim3 b = blk(h->D);
    if (i > 0) {
      const Level& fl = h->levels[i - 1];
      Box in = l.inside();
      const bool transition = h->dist.on() && fl.slab && !l.slab;
      int planes = 0;
      if (transition) {  // every rank restricts its own slab into its part of the replicated level, then the parts are gathered
        planes = (fl.g.N[2] - 2) / 2;
        in.lo[2] = 1 + l.zoffc;
        in.n[2] = planes;
      }
      LAUNCH_D(h, k_restrictL, grd(in, b), b, l.g, fl.g, in, l.L, (const float*)fl.L, l.c[0], l.c[1], l.c[2], transition ? l.zoffc : 0);
      if (transition)
        for (int c = 0; c < h->D; c++) TRY(allgather_planes(h, l, l.L + (size_t)c * l.g.sc, planes));
      launch_bc_vec(h, l.g, l.L, zero, 0, l.L);
    }
    TRY(exch(h, l, l.L, h->D));
    LAUNCH_D(h, k_set_diag, grd(l.inside(), b), b, l.g, l.inside(), (const float*)l.L, l.Dg, l.iD);
    TRY(exch(h, l, l.iD, 1));
    if (l.semi) LAUNCH(h, k_semi_flags, l.fgrid(), 256, l.g, (const float*)l.L, l.Lc[0], l.Lc[1], l.Lc[2], l.zchunk(), l.semi);
  }
  if (h->semi_on) {  // body-free blocks of BDIM-2
    dim3 b = blk(h->D);
    Box in = h->levels[0].inside();
    const dim3 gr = grd(in, b);
    if (!h->nobody) {
      float* q = nullptr;
      TRY(dalloc(h, &q, ((size_t)gr.x * gr.y * gr.z + 3) / 4 + 1));
      h->nobody = (unsigned char*)q;
    }
    LAUNCH_D(h, k_nobody_flags, gr, b, h->g, in, (const float*)h->V, (const float*)h->mu0, (const float*)h->mu1, h->nobody);
    h->nobody_valid = true;
  }
  // uniform-coefficient specialisation: legal iff no body (μ₀≡1, μ₁≡0, V≡0) and every direction periodic
  h->uni = false;
  const Grid& g = h->g;
  // (the sub-grid udf reads the ghost cells of u and adds to the raw flux sum: general kernels)
  if (h->D == 3 && g.per[0] && g.per[1] && (g.per[2] || h->perz_global) && !(h->cfg.flags & WL_FLAG_GENERAL_COEFF) && !h->sgs_on) {
    LAUNCH(h, k_check_uniform, dim3(592, 1, 1), dim3(256, 1, 1), (const float*)h->mu0, (const float*)h->mu1, (const float*)h->V, g, h->red, SLOT_UNI);
    TRY(allreduce_slot(h, SLOT_UNI, WL_NCCL_MAX));
    CK(cudaMemcpyAsync(h->h_out + SLOT_UNI, h->red.out + SLOT_UNI, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->uni = (h->h_out[SLOT_UNI] == 0.0);
  }
  CK(cudaGetLastError());
  return 0;
}

static inline int ensure_hierarchy(wl_handle* h) { return h->pois_dirty ? update_levels(h) : 0; }

static void set_scalar(wl_handle* h, int idx, float v) {
  LAUNCH_PDL(h, k_set_scalar, dim3(1), dim3(1), 0, h->d_scal + idx, v);
}
// The reduction buffer for a launch that reduces into `slot` and whose result the host will read: tagged, so that the folding
// thread publishes the result to the mapped mirror (`active` = the launch really reduces).
static RedBuf red_for(wl_handle* h, int slot, bool active = true) {
  RedBuf r = h->red;
  if (active && r.hseq) {
    if (++h->tag_ctr == 0) h->tag_ctr = 1;
    r.tag = h->tag_ctr;
    h->slot_tag[slot] = r.tag;
  }
  return r;
}
// Result slot(s) → host.  One GPU: the host polls the word the folding thread writes after the values (≈ 1 µs after the kernel's last
// block instead of a copy + stream synchronisation); the stream query covers a launch that did not reduce after all.
// z slabs: the slot was all-reduced on the stream after the kernel — copy and synchronise.
static int read_slot(wl_handle* h, int slot, double* out, int n = 1) {
  const unsigned int want = h->slot_tag[slot];
  const bool reduced_here = !h->dist.on() || h->slot_ar[slot];  // (NCCL all-reduce: the mirror holds this rank's partial value)
  h->slot_tag[slot] = 0;
  h->slot_ar[slot] = false;
  if (h->fast_read && want && reduced_here) {
    volatile unsigned int* sq = h->red.hseq + slot;
    bool ok = true;
    for (unsigned int spins = 1; *sq != want; spins++) {
      if ((spins & 1023u) == 0 && cudaStreamQuery(h->st) != cudaErrorNotReady) {
        ok = (*sq == want);
        break;
      }
    }
    if (ok) {
      std::atomic_thread_fence(std::memory_order_acquire);
      for (int i = 0; i < n; i++) out[i] = reinterpret_cast<volatile double*>(h->red.hout)[slot + i];
      return 0;
    }
  }
  CK(cudaMemcpyAsync(h->h_out + slot, h->red.out + slot, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  for (int i = 0; i < n; i++) out[i] = h->h_out[slot + i];
  return 0;
}

// GaussSeidelRB!(p;it=4,ω)  src/Poisson.jl:141-148
static int gs_smooth(wl_handle* h, Level& l, const float* wp, int x_is_zero, int with_l2) {
  ProfLevel pl(h, l);
  dim3 b = blk(h->D);
  Box in = l.inside();
  if (l.fast && h->fused_gs && (l.g.N[2] - 2) % 2 == 0) {
    // ϵ⁰ + sweep 1, then sweeps 2-4 in place, then increment! (+L₂): five vectorised march launches
    dim3 fb(32, FTY);
    ProlongSrc ps{nullptr, l.g};
    // (z slabs: the ghost planes of ϵ are refreshed after every half-sweep; across the global periodic boundary the sweeps read
    //  the stale r·iD from the exchanged ghost plane of r instead)
    if (h->uni) {
      LAUNCH(h, f_gs_a<true>, l.fgrid(), fb, l.g, l.coef(true), (const float*)l.r, l.eps, l.zchunk());
      if (exch(h, l, l.eps, 1)) return 1;
      for (int k0 = 2; k0 <= 4; k0++) {
        LAUNCH(h, f_gs_half<true>, l.fgrid(), fb, l.g, l.coef(true), (const float*)l.r, l.eps, k0, l.zchunk());
        if (exch(h, l, l.eps, 1)) return 1;
      }
      LAUNCH(h, (f_increment<true, false>), l.fgrid(), fb, l.g, l.coef(true), (const float*)l.eps, ps, l.r, l.x, wp, x_is_zero, l.zchunk(), with_l2,
             red_for(h, SLOT_R2, with_l2), SLOT_R2);
    } else {
      LAUNCH(h, f_gs_a<false>, l.fgrid(), fb, l.g, l.coef(false), (const float*)l.r, l.eps, l.zchunk());
      if (exch(h, l, l.eps, 1)) return 1;
      for (int k0 = 2; k0 <= 4; k0++) {
        LAUNCH(h, f_gs_half<false>, l.fgrid(), fb, l.g, l.coef(false), (const float*)l.r, l.eps, k0, l.zchunk());
        if (exch(h, l, l.eps, 1)) return 1;
      }
      LAUNCH(h, (f_increment<false, false>), l.fgrid(), fb, l.g, l.coef(false), (const float*)l.eps, ps, l.r, l.x, wp, x_is_zero, l.zchunk(), with_l2,
             red_for(h, SLOT_R2, with_l2), SLOT_R2);
    }
    if (exch2(h, l, l.r, l.x)) return 1;
    if (with_l2 && allreduce_slot(h, SLOT_R2, WL_NCCL_SUM)) return 1;
    return 0;
  }
  if (l.slab) return fail("z-slab level without the fused Gauss-Seidel path");
  Lvl d = l.dev();
  LAUNCH_D(h, k_gs_init, grd(in, b), b, d, in);
  Box half = in;
  half.n[0] = (in.n[0] + 1) / 2;
  for (int k0 = 1; k0 <= 4; k0++) LAUNCH_D(h, k_gs_sweep, grd(half, b), b, d, in, k0);
  if (l.fast) {
    ProlongSrc ps{nullptr, l.g, 0, 0, 0};
    if (h->uni)
      LAUNCH(h, (f_increment<true, false>), l.fgrid(), dim3(32, FTY), l.g, l.coef(true), (const float*)l.eps, ps, l.r, l.x, wp, x_is_zero, l.zchunk(), with_l2,
             red_for(h, SLOT_R2, with_l2), SLOT_R2);
    else
      LAUNCH(h, (f_increment<false, false>), l.fgrid(), dim3(32, FTY), l.g, l.coef(false), (const float*)l.eps, ps, l.r, l.x, wp, x_is_zero, l.zchunk(),
             with_l2, red_for(h, SLOT_R2, with_l2), SLOT_R2);
  } else
    LAUNCH_D(h, k_increment, grd(in, b), b, d, in, wp, x_is_zero, with_l2, red_for(h, SLOT_R2, with_l2), SLOT_R2);
  return 0;
}
static int vs_push_r(wl_handle* h, size_t li, int lane);
// Jacobi!(p;ω=1)  src/Poisson.jl:111-114
// With `coarse` given (full coarsening, march kernels) restrict!(coarse.r, fine.r) is fused in; *fused tells whether it was.
// `skip_r_exch`: the caller follows up with vsmooth on this level, which pushes r's ghost plane together with its deeper halo
static int jacobi(wl_handle* h, Level& l, int x_is_zero, Level* coarse, bool* fused_out, bool skip_r_exch = false) {
  ProfLevel pl(h, l);
  dim3 b = blk(h->D);
  Box in = l.inside();
  bool fused = false;
  if (l.fast) {
    fused = coarse && coarse->fullc && l.zchunk() % 2 == 0 && (l.g.N[1] - 2) % 2 == 0 && (l.g.N[2] - 2) % 2 == 0;
    const Grid& gc = fused ? coarse->g : l.g;
    float* rc = fused ? coarse->r : nullptr;
    const int zoffc = (fused && l.slab && !coarse->slab) ? coarse->zoffc : 0;
    if (h->uni && h->divres_uni && fused && h->jacobi2)
      LAUNCH_PDL(h, f_jacobi_uni2, l.fgrid(), dim3(32, FTY / 2), 0, l.g, l.coef(true), (const float*)l.r, l.r2, l.x, l.zchunk(), gc, rc, zoffc, x_is_zero);
    else if (h->uni)
      LAUNCH(h, f_jacobi<true>, l.fgrid(), dim3(32, FTY), l.g, l.coef(true), (const float*)l.r, l.r2, l.x, x_is_zero, l.zchunk(), gc, rc, fused ? 1 : 0,
             zoffc);
    else
      LAUNCH(h, f_jacobi<false>, l.fgrid(), dim3(32, FTY), l.g, l.coef(false), (const float*)l.r, l.r2, l.x, x_is_zero, l.zchunk(), gc, rc, fused ? 1 : 0,
             zoffc);
  } else {
    if (l.slab) return fail("z-slab level without march kernels");
    LAUNCH_D(h, k_jacobi, grd(in, b), b, l.dev(), in, x_is_zero);
  }
  std::swap(l.r, l.r2);
  if (!skip_r_exch)
    TRY(exch(h, l, l.r, 1));
  else if (h->p2p && h->prefetch && h->st2) {
    const size_t li = (size_t)(&l - h->levels.data());
    if (li < 4) TRY(vs_push_r(h, li, 1));  // r is final: its halo for f_vsmooth travels while the coarse levels run
  }
  if (fused && h->dist.on()) {
    if (l.slab && !coarse->slab)
      TRY(allgather_planes(h, *coarse, coarse->r, (l.g.N[2] - 2) / 2, true));
    else
      TRY(exch(h, *coarse, coarse->r, 1, h->vsmooth));
  }
  if (fused_out) *fused_out = fused;
  return 0;
}
// pcg!(p;it=6)  src/Poisson.jl:166-186 — device-driven: all stages of the six iterations are enqueued, the early exits are taken
// on the device (PcgCtl), the host never waits
static int pcg(wl_handle* h, Level& l, int it = 6) {
  dim3 b = blk(h->D);
  Box in = l.inside();
  Lvl d = l.dev();
  PcgCtl* ctl = reinterpret_cast<PcgCtl*>(h->d_scal + 8);
  LAUNCH_D(h, k_pcg, grd(in, b), b, d, in, 0, ctl, h->red, SLOT_RHO);
  for (int i = 1; i <= it; i++) {
    LAUNCH_D(h, k_pcg, grd(in, b), b, d, in, 1, ctl, h->red, SLOT_SIG);
    LAUNCH_D(h, k_pcg, grd(in, b), b, d, in, 2, ctl, h->red, SLOT_RHO);
    if (i == it) break;
    LAUNCH_D(h, k_pcg, grd(in, b), b, d, in, 3, ctl, h->red, SLOT_RHO);
    LAUNCH_D(h, k_pcg, grd(in, b), b, d, in, 4, ctl, h->red, SLOT_RHO);
  }
  return 0;
}
static int l2_norm(wl_handle* h, Level& l, float* out, int want_max = 0) {
  dim3 b = blk(h->D);
  Box in = l.inside();
  LAUNCH_D(h, k_norms, grd(in, b), b, l.dev(), in, red_for(h, want_max ? SLOT_LINF : SLOT_R2), want_max ? SLOT_LINF : SLOT_R2, want_max);
  TRY(allreduce_slot(h, want_max ? SLOT_LINF : SLOT_R2, want_max ? WL_NCCL_MAX : WL_NCCL_SUM));
  double v;
  TRY(read_slot(h, want_max ? SLOT_LINF : SLOT_R2, &v));
  *out = (float)v;
  return 0;
}
// smooth!(p;ω)  src/MultiLevelPoisson.jl:106
static int smooth(wl_handle* h, Level& l, const float* wp, int x_is_zero, int with_l2) {
  if (h->cfg.smoother == WL_SMOOTH_GSRB) return gs_smooth(h, l, wp, x_is_zero, with_l2);
  if (h->dist.on()) return fail("the pcg smoother is not available with z-slab decomposition");
  if (x_is_zero) CK(cudaMemsetAsync(l.x, 0, l.cells() * sizeof(float), h->st));
  TRY(pcg(h, l));
  if (with_l2) {
    dim3 b = blk(h->D);
    Box in = l.inside();
    LAUNCH_D(h, k_norms, grd(in, b), b, l.dev(), in, red_for(h, SLOT_R2), SLOT_R2, 0);
  }
  return 0;
}
// f_vsmooth applies to level li: uniform mode on one GPU, a fully coarsened level below it that is not part of the persistent
// coarse-level kernel's range, sizes that fit the kernel's 8-cell groups and red/black planes, enough planes to fill the pipeline
static bool vs_fusable(const wl_handle* h, size_t li) {
  if (!h->vsmooth || !h->uni || h->D != 3 || h->cfg.smoother != WL_SMOOTH_GSRB) return false;
  if (li + 1 >= h->levels.size()) return false;
  if (h->small_from > 0 && (int)li >= h->small_from) return false;
  const Level& f = h->levels[li];
  const Level& c = h->levels[li + 1];
  const int n0 = f.g.N[0] - 2, n1 = f.g.N[1] - 2, n2 = f.g.N[2] - 2;
  if (f.slab) {  // z slab: the kernel's 5-plane halo comes from the neighbours by peer-to-peer pushes
    if (!h->p2p || n2 < 16 || n2 % 2 || (c.slab && c.g.N[2] - 2 < 4)) return false;
    return f.fast && c.fullc && n0 % 8 == 0 && n1 % 2 == 0 && n0 >= 64 && n1 >= 32;
  }
  return f.fast && c.fullc && n0 % 8 == 0 && n1 % 2 == 0 && n2 % 2 == 0 && n0 >= 64 && n1 >= 32 && n2 >= 64;
}
// The halo of r that f_vsmooth needs on a z slab: the ghost planes (Jacobi! left them to this push) and the four planes beyond them
// on each side, into the neighbours' r / rext.  lane 1: on the side stream, complete at ev_pre[li].
static int vs_push_r(wl_handle* h, size_t li, int lane) {
  Level& f = h->levels[li];
  const int n2 = f.g.N[2] - 2;
  PlaneMove mv[10];
  int m = 0;
  const i64 s2 = f.g.s[2];
  mv[m++] = {f.r + s2 * n2, 1, f.r};
  mv[m++] = {f.r + s2 * 1, 0, f.r + s2 * (n2 + 1)};
  for (int k = 0; k < 4; k++) {
    mv[m++] = {f.r + s2 * (n2 - 4 + k), 1, f.rext + s2 * k};
    mv[m++] = {f.r + s2 * (2 + k), 0, f.rext + s2 * (4 + k)};
  }
  TRY(p2p_push(h, f.g, mv, m, false, lane, true));
  if (lane) {
    CK(cudaEventRecord(h->ev_pre[li], h->st2));
    h->pre_done[li] = true;
  }
  return 0;
}
// prolongate! + increment! + GaussSeidelRB! + increment! (+ L₂) of level li in one launch (wl_vsmooth.cuh)
static int vsmooth(wl_handle* h, size_t li, const float* wp, int with_l2) {
  Level& f = h->levels[li];
  Level& c = h->levels[li + 1];
  ProfLevel pl(h, f);
  // coefficient form of the level (vs_pair): 1 = unit faces (finest level), 2 = powers of two ≥ 1, 0 = as written
  const Coef kk = f.coef(true);
  auto pow2ge1 = [](float v) { int e; return v >= 1.f && std::frexp(v, &e) == 0.5f; };
  const int lm = (kk.Lc[0] == 1.f && kk.Lc[1] == 1.f && kk.Lc[2] == 1.f) ? 1 : (pow2ge1(kk.Lc[0]) && pow2ge1(kk.Lc[1]) && pow2ge1(kk.Lc[2])) ? 2 : 0;
  typedef void (*VsKern)(const VsArgs, RedBuf, int);
  static const VsKern vs_tab[2][2][3] = {
      {{f_vsmooth<false, false, 0>, f_vsmooth<false, false, 1>, f_vsmooth<false, false, 2>},
       {f_vsmooth<false, true, 0>, f_vsmooth<false, true, 1>, f_vsmooth<false, true, 2>}},
      {{f_vsmooth<true, false, 0>, f_vsmooth<true, false, 1>, f_vsmooth<true, false, 2>},
       {f_vsmooth<true, true, 0>, f_vsmooth<true, true, 1>, f_vsmooth<true, true, 2>}}};
  if (!h->attr_vs) {  // per handle: the attribute belongs to the device the handle lives on
    for (int i = 0; i < 12; i++) cudaFuncSetAttribute((&vs_tab[0][0][0])[i], cudaFuncAttributeMaxDynamicSharedMemorySize, VS_SMEM);
    h->attr_vs = true;
  }
  const int n0 = f.g.N[0] - 2, n1 = f.g.N[1] - 2, n2 = f.g.N[2] - 2;
  // tile height: the level's rows split evenly over the smallest number of tiles (128 rows: 4 × 32 instead of 3 × 42 + 2 — the rows
  // beyond n1 would be marched for nothing)
  const int cy = cdiv(n1, cdiv(n1, VS_CY));
  const int tiles = cdiv(n0, VS_CX) * cdiv(n1, cy);
  // z chunks: every chunk pays 12 planes of pipeline fill; pick the count that minimises waves × (planes per chunk + 12)
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
  int best = 1;
  long bestcost = -1;
  for (int nz = 1; nz <= std::max(1, n2 / 8); nz++) {
    const long cost = (long)cdiv(tiles * nz, nsm * VS_BPS) * (cdiv(n2, nz) + 12);
    if (bestcost < 0 || cost < bestcost) {
      bestcost = cost;
      best = nz;
    }
  }
  if (h->vs_nz > 0) best = h->vs_nz;
  const int zc = cdiv(n2, best);
  const Coef k = f.coef(true);
  VsArgs a;
  a.g = f.g;
  a.gc = c.g;
  a.xc = c.x;
  a.r = f.r;
  a.r2 = f.r2;
  a.x = f.x;
  a.wp = wp;
  a.L0 = k.Lc[0];
  a.L1 = k.Lc[1];
  a.L2 = k.Lc[2];
  a.D = k.Dc;
  a.iD = k.iDc;
  a.zchunk = zc;
  a.cy = cy;
  a.th = cy + 2 * VS_HALO;
  a.slab = f.slab ? 1 : 0;
  a.cslab = c.slab ? 1 : 0;
  a.zoff = f.slab ? f.g.zoff : 0;
  a.n2g = f.Ng2 - 2;
  a.rext = f.rext;
  a.cxext = c.xext;
  if (f.slab) {
    // planes −4 … −1 and n2+2 … n2+5 of r (the ghost planes 0 and n2+1 are current since Jacobi!'s exchange), and, when the coarse
    // level is a slab too, its planes −2, −1 and nc+2, nc+3 of x: pushed straight into the neighbours' rext / xext
    if (li < 4 && h->pre_done[li]) {  // (the r planes went out on the side stream right after Jacobi!, vs_push_r)
      CK(cudaStreamWaitEvent(h->st, h->ev_pre[li], 0));
      h->pre_done[li] = false;
    } else
      TRY(vs_push_r(h, li, 0));
    PlaneMove mv[10];
    int m = 0;
    if (c.slab && !vs_fusable(h, li + 1)) {  // (a coarse level that ran vsmooth itself has pushed its xext planes already)
      const i64 c2 = c.g.s[2];
      const int nc = c.g.N[2] - 2;
      m = 0;
      for (int k = 0; k < 2; k++) {
        mv[m++] = {c.x + c2 * (nc - 2 + k), 1, c.xext + c2 * k};
        mv[m++] = {c.x + c2 * (2 + k), 0, c.xext + c2 * (2 + k)};
      }
      TRY(p2p_push(h, c.g, mv, m));
    }
  }
  dim3 gr(cdiv(n0, VS_CX), cdiv(n1, cy), cdiv(n2, zc));
  prof_begin(h, "f_vsmooth");
  pdl_launch(h, vs_tab[with_l2 ? 1 : 0][f.slab ? 1 : 0][lm], gr, dim3(VS_NT), VS_SMEM, a, red_for(h, SLOT_R2, with_l2), (int)SLOT_R2);
  prof_end(h);
  h->launches++;
  std::swap(f.r, f.r2);
  CK(cudaGetLastError());
  if (f.slab && li > 0) {
    // ghost planes of r and x, and for the finer level's vsmooth the two planes of x beyond them, in one push
    PlaneMove mv[8];
    int m = 0;
    const i64 s2 = f.g.s[2];
    mv[m++] = {f.r + s2 * n2, 1, f.r};
    mv[m++] = {f.r + s2 * 1, 0, f.r + s2 * (n2 + 1)};
    mv[m++] = {f.x + s2 * n2, 1, f.x};
    mv[m++] = {f.x + s2 * 1, 0, f.x + s2 * (n2 + 1)};
    for (int k = 0; k < 2; k++) {
      mv[m++] = {f.x + s2 * (n2 - 2 + k), 1, f.xext + s2 * k};
      mv[m++] = {f.x + s2 * (2 + k), 0, f.xext + s2 * (2 + k)};
    }
    TRY(p2p_push(h, f.g, mv, m, false, 0, true));
  } else
    TRY(exch2(h, f, f.r, f.x, true));
  if (with_l2) TRY(allreduce_slot(h, SLOT_R2, WL_NCCL_SUM));
  return 0;
}
// ---- flattening of the coarse end of the V-cycle for k_small_levels ----------------------------------------------------
static SmallOp op_base(const wl_handle* h, const Level& l, int type) {
  SmallOp o;
  memset(&o, 0, sizeof o);
  o.type = type;
  o.g = l.g;
  o.c = l.coef(h->uni);
  o.gc = l.g;
  o.lvl = l.dev();
  o.box = l.inside();
  o.zchunk = l.zchunk();
  if (type <= OP_F_PROLONG) {
    dim3 fg = l.fgrid();
    o.vg[0] = fg.x;
    o.vg[1] = fg.y;
    o.vg[2] = fg.z;
  } else {  // general bodies with a 32×8×1 block
    o.vg[0] = cdiv(o.box.n[0], 32);
    o.vg[1] = cdiv(o.box.n[1], FTY);
    o.vg[2] = o.box.n[2];
  }
  return o;
}
static bool gs_fusable(const wl_handle* h, const Level& l) { return l.fast && h->fused_gs && (l.g.N[2] - 2) % 2 == 0; }
static void emit_gs(wl_handle* h, Level& l, int x_is_zero) {
  if (gs_fusable(h, l)) {
    h->ops.push_back(op_base(h, l, OP_F_GSA));
    for (int k0 = 2; k0 <= 4; k0++) {
      SmallOp o = op_base(h, l, OP_F_GSHALF);
      o.k0 = k0;
      h->ops.push_back(o);
    }
  } else {
    h->ops.push_back(op_base(h, l, OP_K_GSINIT));
    for (int k0 = 1; k0 <= 4; k0++) {
      SmallOp o = op_base(h, l, OP_K_GSSWEEP);
      o.k0 = k0;
      o.vg[0] = cdiv((o.box.n[0] + 1) / 2, 32);
      h->ops.push_back(o);
    }
  }
  SmallOp o = op_base(h, l, l.fast ? OP_F_INC : OP_K_INC);
  o.x_is_zero = x_is_zero;
  h->ops.push_back(o);
}
// `tiny` > 0: levels ≥ tiny are run by k_tiny_uni: the recursion stops above them and *split marks the place in the list
static void emit_vcycle(wl_handle* h, size_t li, int tiny, int* split) {
  Level& fine = h->levels[li];
  Level& coarse = h->levels[li + 1];
  // Jacobi! (+ restrict!)
  const bool fused = fine.fast && coarse.fullc && fine.zchunk() % 2 == 0 && (fine.g.N[1] - 2) % 2 == 0 && (fine.g.N[2] - 2) % 2 == 0;
  {
    SmallOp o = op_base(h, fine, fine.fast ? OP_F_JACOBI : OP_K_JACOBI);
    o.x_is_zero = li > 0;
    o.do_restrict = fused;
    o.gc = fused ? coarse.g : fine.g;
    o.other = fused ? coarse.r : nullptr;
    h->ops.push_back(o);
    std::swap(fine.r, fine.r2);
  }
  if (!fused) {
    SmallOp o = op_base(h, fine, OP_K_RESTRICT);
    o.gc = coarse.g;
    o.box = coarse.inside();
    o.vg[0] = cdiv(o.box.n[0], 32);
    o.vg[1] = cdiv(o.box.n[1], FTY);
    o.vg[2] = o.box.n[2];
    o.other = coarse.r;
    for (int d = 0; d < 3; d++) o.cm[d] = coarse.c[d];
    h->ops.push_back(o);
  }
  const bool last = (li + 2 >= h->levels.size());
  if (tiny > 0 && (int)li + 1 == tiny) {
    *split = (int)h->ops.size();  // Vcycle!(l=tiny) and smooth!(levels[tiny]) happen here, in k_tiny_uni
  } else {
    if (!last) emit_vcycle(h, li + 1, tiny, split);
    emit_gs(h, coarse, last ? 1 : 0);
  }
  {
    const bool fp = fine.fast && coarse.fullc;
    SmallOp o = op_base(h, fine, fp ? OP_F_PROLONG : OP_K_PROLONG);
    o.gc = coarse.g;
    o.other = coarse.x;
    for (int d = 0; d < 3; d++) o.cm[d] = coarse.c[d];
    h->ops.push_back(o);
  }
}
static int launch_tiny(wl_handle* h, const float* wp) {
  TinyArgs a;
  memset(&a, 0, sizeof a);
  const size_t T = (size_t)h->tiny_from;
  a.nlev = (int)(h->levels.size() - T);
  size_t floats = 0;
  for (int q = 0; q < a.nlev; q++) {
    const Level& l = h->levels[T + q];
    const Coef k = l.coef(true);
    for (int d = 0; d < 3; d++) {
      a.n[q][d] = l.g.N[d] - 2;
      a.L[q][d] = k.Lc[d];
    }
    a.D[q] = k.Dc;
    a.iD[q] = k.iDc;
    floats += (size_t)3 * a.n[q][0] * a.n[q][1] * a.n[q][2];
  }
  Level& l0 = h->levels[T];
  a.g0 = l0.g;
  a.r0 = l0.r;
  a.x0 = l0.x;
  a.r0out = l0.r;
  a.wp = wp;
  if (!h->attr_tiny) {
    cudaFuncSetAttribute(k_tiny_uni, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    h->attr_tiny = true;
  }
  prof_begin(h, "k_tiny_uni");
  pdl_launch(h, k_tiny_uni, dim3(1), dim3(1024), floats * sizeof(float), a);
  prof_end(h);
  h->launches++;
  return 0;
}
static void emit_vcycle(wl_handle* h, size_t li, int tiny, int* split);
// k_tiny_gen's operation list: the list k_small_levels would run for "Vcycle!(ml; l=T); smooth!(levels[T])", emitted over shadow
// copies of the levels ≥ T whose Grid is dense and whose pointers are arena offsets (encoded (offset+1)·4, tinyg_fix).
static int build_tinyg(wl_handle* h) {
  const size_t T = (size_t)h->tg_from, nl = h->levels.size();
  TinyGenArgs& a = h->tg_args;
  memset(&a, 0, sizeof a);
  a.nlev = (int)(nl - T);
  std::vector<Level> real(h->levels.begin() + T, h->levels.end());
  auto enc = [](size_t off) { return reinterpret_cast<float*>((uintptr_t)(off + 1) * 4); };
  size_t off = 0;
  for (size_t q = T; q < nl; q++) {
    Level& l = h->levels[q];
    const size_t n = (size_t)l.g.N[0] * l.g.N[1] * l.g.N[2];
    TinyGenLvl& t = a.lv[q - T];
    t.g = l.g;
    t.L = l.L;
    t.Dg = l.Dg;
    t.iD = l.iD;
    t.off = (int)off;
    t.cells = (int)n;
    t.x = l.x;
    t.r = l.r;
    Grid d = l.g;  // dense: no pitch, no offset
    d.xo = 0;
    d.px = l.g.N[0];
    d.s[0] = 1;
    d.s[1] = l.g.N[0];
    d.s[2] = (i64)l.g.N[0] * l.g.N[1];
    d.sc = (i64)n;
    l.g = d;
    l.L = enc(off);
    l.Dg = enc(off + 3 * n);
    l.iD = enc(off + 4 * n);
    l.x = enc(off + 5 * n);
    l.eps = enc(off + 6 * n);
    l.r = enc(off + 7 * n);
    l.r2 = enc(off + 8 * n);
    l.z = nullptr;
    l.fast = false;
    l.semi = nullptr;
    off += 9 * n;
  }
  a.arena_floats = (int)off;
  a.r0 = real[0].r;
  a.r_in_off = a.lv[0].off + 7 * a.lv[0].cells;
  std::vector<SmallOp> saved;
  saved.swap(h->ops);
  int split = -1;
  const bool last = (T + 1 >= nl);
  if (!last) emit_vcycle(h, T, 0, &split);
  emit_gs(h, h->levels[T], last ? 1 : 0);
  // where every level's residual ended up (Jacobi! writes it out of place and the roles swap)
  for (size_t q = T; q < nl; q++) a.lv[q - T].r_out_off = (int)(((uintptr_t)h->levels[q].r >> 2) - 1);
  std::vector<SmallOp> ops;
  ops.swap(h->ops);
  h->ops.swap(saved);
  for (size_t q = T; q < nl; q++) h->levels[q] = real[q - T];
  for (const SmallOp& o : ops)
    if (o.type < OP_K_JACOBI) return fail("k_tiny_gen: unexpected march operation in the list");
  h->tg_nops = (int)ops.size();
  if (!h->d_tgops) {
    void* q = nullptr;
    CK(cudaMalloc(&q, 256 * sizeof(SmallOp)));
    h->d_tgops = (SmallOp*)q;
    h->allocs.push_back(q);
  }
  if (h->tg_nops > 256) return fail("too many tiny-level operations (%d)", h->tg_nops);
  CK(cudaMemcpyAsync(h->d_tgops, ops.data(), ops.size() * sizeof(SmallOp), cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));  // (ops is a local)
  h->tg_ready = true;
  return 0;
}
static int launch_tinyg(wl_handle* h, const float* wp) {
  if (!h->tg_ready) TRY(build_tinyg(h));
  const size_t bytes = (size_t)((h->tg_args.arena_floats + 3) & ~3) * sizeof(float) + (size_t)h->tg_nops * sizeof(SmallOp);
  if (bytes > (size_t)220 * 1024) return fail("k_tiny_gen: %zu bytes of shared memory", bytes);
  if (!h->attr_tg) {
    CK(cudaFuncSetAttribute(k_tiny_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    h->attr_tg = true;
  }
  prof_begin(h, "k_tiny_gen");
  pdl_launch(h, k_tiny_gen, dim3(1), dim3(16, 8, 8), bytes, (const SmallOp*)h->d_tgops, h->tg_nops, wp, h->tg_args);
  prof_end(h);
  h->launches++;
  return 0;
}
// Runs "Vcycle!(ml; l=small_from) ; smooth!(levels[small_from])" — everything a V-cycle does at and below level small_from:
// the levels ≥ small_from in the cooperative k_small_levels, except (uniform mode) the innermost levels ≥ tiny_from, which
// k_tiny_uni runs on one block in shared memory between the two halves of the list.
static int run_small_levels(wl_handle* h, const float* wp) {
  const size_t ls = (size_t)h->small_from;
  const int tiny = !h->tiny_on ? 0 : (h->uni ? h->tiny_from : h->tg_from);
  auto run_tiny = [&]() -> int { return h->uni ? launch_tiny(h, wp) : launch_tinyg(h, wp); };
  if (tiny > 0 && (size_t)tiny <= ls) return run_tiny();  // the whole coarse end is tiny
  h->ops.clear();
  int split = -1;
  const bool last = (ls + 1 >= h->levels.size());
  if (!last) emit_vcycle(h, ls, tiny, &split);
  emit_gs(h, h->levels[ls], last ? 1 : 0);
  const int nops = (int)h->ops.size();
  if (nops > 256) return fail("too many coarse-level operations (%d)", nops);
  memcpy(h->h_ops, h->ops.data(), nops * sizeof(SmallOp));
  CK(cudaMemcpyAsync(h->d_ops, h->h_ops, nops * sizeof(SmallOp), cudaMemcpyHostToDevice, h->st));
  auto coop = [&](int first, int n) -> int {
    if (n <= 0) return 0;
    const SmallOp* dops = h->d_ops + first;
    prof_begin(h, "k_small_levels");
    // cooperative + (if the driver takes the combination) programmatic dependent launch
    auto go = [&](int nattr) -> cudaError_t {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(h->small_grid);
      cfg.blockDim = dim3(32, FTY);
      cfg.stream = h->st;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeCooperative;
      at[0].val.cooperative = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = nattr;
      return h->uni ? cudaLaunchKernelEx(&cfg, k_small_levels<true>, dops, n, wp) : cudaLaunchKernelEx(&cfg, k_small_levels<false>, dops, n, wp);
    };
    cudaError_t e = go((h->pdl && h->coop_pdl) ? 2 : 1);
    if (e != cudaSuccess && h->pdl && h->coop_pdl) {  // not accepted together: cooperative only, from now on
      cudaGetLastError();
      h->coop_pdl = false;
      e = go(1);
    }
    prof_end(h);
    h->launches++;
    if (e != cudaSuccess) return fail("cooperative launch of k_small_levels: %s", cudaGetErrorString(e));
    return 0;
  };
  if (split < 0) return coop(0, nops);
  TRY(coop(0, split));
  TRY(run_tiny());
  return coop(split, nops - split);
}

// Vcycle!(ml;l,ω)  src/MultiLevelPoisson.jl:88-101
// `defer_up`: the caller runs level li's prolongation + increment! fused with its smoother (vsmooth) instead of here
static int vcycle(wl_handle* h, size_t li, const float* wp, bool defer_up = false) {
  Level& fine = h->levels[li];
  Level& coarse = h->levels[li + 1];
  dim3 b = blk(h->D);
  Box cin = coarse.inside();
  bool fused = false;
  TRY(jacobi(h, fine, li > 0, &coarse, &fused, defer_up && fine.slab));
  if (!fused) {
    if (fine.slab) return fail("z-slab levels need the fused restriction (even sizes, full coarsening)");
    LAUNCH_D(h, k_restrict, grd(cin, b), b, coarse.g, fine.g, cin, coarse.r, (const float*)fine.r, coarse.c[0], coarse.c[1], coarse.c[2]);
  }
  const bool last = (li + 2 >= h->levels.size());
  if (h->small_from > 0 && (int)li + 1 == h->small_from) {
    TRY(run_small_levels(h, wp));  // the whole coarse end in one cooperative launch
  } else {
    const bool fu = !last && vs_fusable(h, li + 1);
    if (!last) TRY(vcycle(h, li + 1, wp, fu));
    if (fu)
      TRY(vsmooth(h, li + 1, wp, 0));
    else
      TRY(smooth(h, coarse, wp, last ? 1 : 0, 0));
  }
  if (defer_up) return 0;
  ProfLevel pl(h, fine);
  Box fin = fine.inside();
  if (fine.fast && coarse.fullc) {
    ProlongSrc ps{coarse.x, coarse.g, 0, 0, 0};
    if (fine.slab && !coarse.slab) {  // slab level over a replicated level: map local fine planes to global coarse planes
      ps.zoff = fine.g.zoff;
      ps.N2g = fine.Ng2;
      ps.perg = h->perz_global;
    }
    if (h->uni)
      LAUNCH(h, (f_increment<true, true>), fine.fgrid(), dim3(32, FTY), fine.g, fine.coef(true), (const float*)nullptr, ps, fine.r, fine.x, wp, 0,
             fine.zchunk(), 0, h->red, SLOT_R2);
    else
      LAUNCH(h, (f_increment<false, true>), fine.fgrid(), dim3(32, FTY), fine.g, fine.coef(false), (const float*)nullptr, ps, fine.r, fine.x, wp, 0,
             fine.zchunk(), 0, h->red, SLOT_R2);
  } else {
    if (fine.slab) return fail("z-slab levels need the march prolongation");
    LAUNCH_D(h, k_prolong_inc, grd(fin, b), b, fine.dev(), coarse.g, (const float*)coarse.x, fin, wp, coarse.c[0], coarse.c[1], coarse.c[2]);
  }
  TRY(exch(h, fine, fine.r, 1));
  return 0;
}

static void log_row(wl_handle* h, int it, float rinf, float r2, float w) {
  if (!h->logging) return;
  h->log.push_back((float)it);
  h->log.push_back(rinf);
  h->log.push_back(r2);
  h->log.push_back(w);
}

// residual!(p) + L₂  (src/Poisson.jl:92-98,189); with_div fuses div + x.*=dt of mom_project! (src/Flow.jl:225)
static int residual(wl_handle* h, int with_div, float w, float* r2) {
  Level& l = h->levels[0];
  dim3 b = blk(h->D);
  Box in = l.inside();
  float count = 1;
  for (int d = 0; d < h->D; d++) count *= (float)((d == 2 && l.slab ? l.Ng2 : l.g.N[d]) - 2);
  if (h->dist.on() && !(l.fast && with_div)) return fail("standalone residual! is not available with z-slab decomposition");
  if (l.fast && with_div) {
    if (h->uni && h->divres_uni)
      LAUNCH_PDL(h, f_divres_uni, l.fgrid(), dim3(32, FTY), 0, l.g, l.coef(true), (const float*)h->u, (const float*)h->p, l.x, l.r, dtp(h), w, l.zchunk(),
                 red_for(h, SLOT_RSUM), (int)SLOT_RSUM);
    else if (h->uni)
      LAUNCH(h, f_div_residual<true>, l.fgrid(), dim3(32, FTY), l.g, l.coef(true), (const float*)h->u, (const float*)h->p, l.x, l.r, l.z, dtp(h), w,
             l.zchunk(), red_for(h, SLOT_RSUM), SLOT_RSUM);
    else
      LAUNCH(h, f_div_residual<false>, l.fgrid(), dim3(32, FTY), l.g, l.coef(false), (const float*)h->u, (const float*)h->p, l.x, l.r, l.z, dtp(h), w,
             l.zchunk(), red_for(h, SLOT_RSUM), SLOT_RSUM);
    TRY(allreduce_slot(h, SLOT_RSUM, WL_NCCL_SUM, 2));  // Σr and Σr² (adjacent slots)
    // residual! subtracts the mean only when |s| > 2eps (src/Poisson.jl:96); otherwise r is final and the Σr² of the same pass is L₂
    static_assert(SLOT_R2 == SLOT_RSUM + 1, "f_div_residual reduces Σr and Σr² into adjacent slots");
    double sums[2];
    TRY(read_slot(h, SLOT_RSUM, sums, 2));
    const float s = (float)sums[0] / count;
    if (std::fabs(s) > 2.f * 1.1920929e-7f) {
      LAUNCH(h, f_resid_fix, l.fgrid(), dim3(32, FTY), l.g, l.r, count, l.zchunk(), red_for(h, SLOT_R2), SLOT_RSUM, SLOT_R2);
      TRY(allreduce_slot(h, SLOT_R2, WL_NCCL_SUM));
      TRY(exch(h, l, l.r, 1, true));
    } else {
      TRY(exch(h, l, l.r, 1, true));
      *r2 = (float)sums[1];
      return 0;
    }
  } else {
    LAUNCH_D(h, k_div_residual, grd(in, b), b, l.dev(), in, (const float*)h->u, (const float*)h->p, dtp(h), w, with_div, h->red, SLOT_RSUM);
    LAUNCH_D(h, k_resid_fix, grd(in, b), b, l.dev(), in, count, red_for(h, SLOT_R2), SLOT_RSUM, SLOT_R2);
  }
  double v;
  TRY(read_slot(h, SLOT_R2, &v));
  *r2 = (float)v;
  return 0;
}

// solver!(ml) src/MultiLevelPoisson.jl:108-127  /  solver!(p::Poisson) src/Poisson.jl:204-214   (after residual!)
static int solve_after_residual(wl_handle* h, float r2, int* iters_out) {
  Level& p = h->levels[0];
  float rinf = NAN;
  if (h->logging) TRY(l2_norm(h, p, &rinf, 1));
  int np = 0;
  if (h->cfg.pois_kind == WL_POIS_MULTILEVEL) {
    float w = 1.f;
    log_row(h, np, rinf, r2, w);
    while (np < h->itmx) {
      Nvtx rv("Vcycle! + smooth!");
      set_scalar(h, 0, w);
      if (vs_fusable(h, 0)) {
        TRY(vcycle(h, 0, h->d_scal + 0, true));
        TRY(vsmooth(h, 0, h->d_scal + 0, 1));
      } else {
        TRY(vcycle(h, 0, h->d_scal + 0));
        TRY(smooth(h, p, h->d_scal + 0, 0, 1));
      }
      double v;
      TRY(read_slot(h, SLOT_R2, &v));
      const float rnew = (float)v;
      np++;
      if (h->logging) TRY(l2_norm(h, p, &rinf, 1));
      log_row(h, np, rinf, rnew, w);
      if (rnew >= r2)
        w = (float)std::max(0.2, 0.9 * (double)w);
      else if (rnew < r2)
        w = (float)std::min(1.0, 1.02 * (double)w);
      r2 = rnew;
      if ((double)r2 < h->tol) break;
    }
  } else {
    log_row(h, np, rinf, r2, 1.f);
    while (np < h->itmx) {
      TRY(pcg(h, p));
      TRY(l2_norm(h, p, &r2));
      np++;
      if (h->logging) TRY(l2_norm(h, p, &rinf, 1));
      log_row(h, np, rinf, r2, 1.f);
      if ((double)r2 < h->tol) break;
    }
  }
  h->iters.push_back((int16_t)np);
  if (iters_out) *iters_out = np;
  return 0;
}

// ---- momentum -------------------------------------------------------------------------------
template <int LAM>
static void conv_launch(wl_handle* h, const float* ua, int mode) {
  const Grid& g = h->g;
  dim3 b = blk(h->D);
  Box all = h->levels[0].all();
  prof_begin(h, "k_conv_bdim1");
  if (h->D == 3)
    pdl_launch(h, k_conv_bdim1<3, LAM>, grd(all, b), b, 0, g, all, ua, (const float*)h->u0, (const float*)h->V, h->f, h->sigma, dtp(h), h->cfg.nu, mode, h->fc);
  else
    pdl_launch(h, k_conv_bdim1<2, LAM>, grd(all, b), b, 0, g, all, ua, (const float*)h->u0, (const float*)h->V, h->f, h->sigma, dtp(h), h->cfg.nu, mode, h->fc);
  prof_end(h);
  h->launches++;
}
static void conv_bdim1(wl_handle* h, const float* ua, int mode) {
  if (h->cfg.lambda == WL_QUICK)
    conv_launch<0>(h, ua, mode);
  else if (h->cfg.lambda == WL_CDS)
    conv_launch<1>(h, ua, mode);
  else
    conv_launch<2>(h, ua, mode);
}

template <int LAM, bool FUSE>
static int fconv_launch(wl_handle* h, const float* ua, float* out, int corrector) {
  const Grid& g = h->g;
  const int XM = FUSE ? g.N[0] - 2 : g.N[0] - 1, YM = FUSE ? g.N[1] - 2 : g.N[1] - 1, ZM = FUSE ? g.N[2] - 2 : g.N[2] - 1;
  const int zchunk = std::min(32, ZM);
  dim3 gr(cdiv(XM, 32), cdiv(YM, CTY), cdiv(ZM, zchunk));
  const bool nowall = g.per[0] && g.per[1] && (g.per[2] || (g.zopen[0] && g.zopen[1]));
  if (FUSE && nowall && h->conv4 && (g.N[0] - 2) % 4 == 0) {  // uniform mode: four cells per thread
    bool& attr = h->attr_c4[LAM];
    if (!attr) {
      cudaFuncSetAttribute(fm_conv4<LAM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C4SMEM);
      cudaFuncSetAttribute(fm_conv4<LAM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C4SMEM);
      attr = true;
    }
    int zc = std::min(h->conv4_zchunk, g.N[2] - 2);
    // small grids: shorter z chunks (each pays one extra plane of fluxes) until every SM has two blocks
    while (zc > 8 && (long)cdiv(g.N[0] - 2, 128) * cdiv(g.N[1] - 2, C4TY) * cdiv(g.N[2] - 2, zc) < 2 * 148) zc /= 2;
    dim3 g4(cdiv(g.N[0] - 2, 128), cdiv(g.N[1] - 2, C4TY), cdiv(g.N[2] - 2, zc));
    // a field nobody range-checked while writing it (upload, BC kernels, the unfused correction) is checked now (range_note)
    if (!h->range_checked) {
      LAUNCH(h, k_range_check, 592, 256, reinterpret_cast<const float4*>(ua), (long long)(g.sc * 3 / 4), h->d_flags);
      if (h->dist.on()) LAUNCH(h, k_range_check, 64, 256, reinterpret_cast<const float4*>(h->uext), (long long)(g.s[2] * 6 / 4), h->d_flags);
    }
    h->range_checked = false;  // `out` is a new field
    // the fast-division instance, then the IEEE-division instance on the same grid: the device-side range word picks the one that
    // runs (the other returns at once), so no host decision and no second pass over marked blocks
    prof_begin(h, "fm_conv4");
    pdl_launch(h, fm_conv4<LAM, false>, g4, dim3(32, C4TY), C4SMEM, g, ua, (const float*)h->u0, out, dtp(h), h->cfg.nu, zc, corrector, h->red, (int)SLOT_PHIMAX, (const float*)h->uext, h->d_flags + 2, h->fc);
    prof_end(h);
    prof_begin(h, "fm_conv4_exact");
    pdl_launch(h, fm_conv4<LAM, true>, g4, dim3(32, C4TY), C4SMEM, g, ua, (const float*)h->u0, out, dtp(h), h->cfg.nu, zc, corrector, h->red, (int)SLOT_PHIMAX, (const float*)h->uext, h->d_flags + 2, h->fc);
    prof_end(h);
    h->launches += 2;
    return 0;
  }
  prof_begin(h, "fm_conv");
  if (nowall)
    pdl_launch(h, fm_conv<LAM, FUSE, true>, gr, dim3(32, CTY), sizeof(ConvTile) + 2 * 3 * CTY * 32 * sizeof(float), g, ua, (const float*)h->u0, (const float*)h->V, out, h->sigma, dtp(h),
               h->cfg.nu, zchunk, corrector, h->red, (int)SLOT_PHIMAX, (const float*)h->uext, h->d_flags, h->fc);
  else
    pdl_launch(h, fm_conv<LAM, FUSE, false>, gr, dim3(32, CTY), sizeof(ConvTile) + 2 * 3 * CTY * 32 * sizeof(float), g, ua, (const float*)h->u0, (const float*)h->V, out, h->sigma, dtp(h),
               h->cfg.nu, zchunk, corrector, h->red, (int)SLOT_PHIMAX, (const float*)h->uext, h->d_flags, h->fc);
  prof_end(h);
  h->launches++;
  return 0;
}
template <bool FUSE>
static int fconv(wl_handle* h, const float* ua, float* out, int corrector) {
  if (h->cfg.lambda == WL_QUICK) return fconv_launch<0, FUSE>(h, ua, out, corrector);
  if (h->cfg.lambda == WL_CDS) return fconv_launch<1, FUSE>(h, ua, out, corrector);
  return fconv_launch<2, FUSE>(h, ua, out, corrector);
}

// Momentum update of mom_predict!/mom_correct! up to (not including) BC!: conv_diff! → BDIM! → scale_u!
static int momentum(wl_handle* h, int corrector) {
  const Grid& g = h->g;
  dim3 b = blk(h->D);
  Level& l = h->levels[0];
  Box in = l.inside();
  if (h->D == 3 && h->uni) {
    // uniform mode: u_new straight from the flux kernel; the corrector must not update u in place (stencil reads), so it
    // writes into the f buffer (unused in this mode) and the two are swapped
    if (!corrector)
      TRY(fconv<true>(h, h->u0, h->u, 0));
    else {
      TRY(fconv<true>(h, h->u, h->f, 1));
      std::swap(h->u, h->f);
    }
    return 0;
  }
  if (h->sgs_on) {
    // conv_diff! → udf! = sgs! → accelerate! → BDIM-1 (src/Flow.jl:191-193,206-208): the raw flux sum, νₜ, then the gather kernel
    const float* ua = corrector ? h->u : h->u0;
    Box all = l.all();
    conv_bdim1(h, ua, 0);
    LAUNCH_D(h, k_sgs_nut, grd(in, b), b, g, in, ua, h->nut, h->sgs_c2);
    LAUNCH_D(h, k_sgs_apply, grd(all, b), b, g, all, ua, (const float*)h->nut, (const float*)h->u0, (const float*)h->V, h->f, h->sigma, dtp(h), h->fc);
  } else if (h->D == 3) {
    TRY(fconv<false>(h, corrector ? h->u : h->u0, h->f, corrector));
    dim3 pb(32, 8, 1);
    int m0 = std::max(g.N[0], g.N[1]), m1 = std::max(g.N[1], g.N[2]);
    LAUNCH(h, k_f_lowghost, dim3(cdiv(m0, 32), cdiv(m1, 8), 3), pb, g, (const float*)h->u0, (const float*)h->V, h->f, dtp(h), h->fc);
    TRY(exch(h, l, h->f, 3));  // BDIM-2 reads f one plane beyond the slab
  } else
    conv_bdim1(h, corrector ? h->u : h->u0, 1);
  LAUNCH_D(h, k_bdim2, grd(in, b), b, g, in, h->u, (const float*)h->f, (const float*)h->V, (const float*)h->mu0, (const float*)h->mu1, corrector,
           (const unsigned char*)(h->nobody_valid ? h->nobody : nullptr));
  return 0;
}

// CFL(a) → *dt_out on the device
static void cfl(wl_handle* h, float* dt_out) {
  Level& l = h->levels[0];
  const Grid& g = h->g;
  if (l.fast) {
    const int fin = h->dist.on() ? 0 : 1;  // z slabs: the maxima are all-reduced first, then k_cfl_final forms Δt
    const int sg = h->uni ? SLOT_PHIMAX : SLOT_CFL;
    if (h->uni) {
      LAUNCH(h, f_cfl<true>, l.fgrid(), dim3(32, FTY), g, (const float*)h->u, h->sigma, h->cfg.nu, dt_out, l.zchunk(), h->red, SLOT_CFLINT, sg, fin);
    } else {
      int m0 = std::max(g.N[0], g.N[1]), m1 = std::max(g.N[1], g.N[2]);
      LAUNCH(h, k_sigma_ghostmax, dim3(cdiv(m0, 32), cdiv(m1, 8), 6), dim3(32, 8, 1), g, (const float*)h->sigma, h->red, SLOT_CFL);
      LAUNCH(h, f_cfl<false>, l.fgrid(), dim3(32, FTY), g, (const float*)h->u, h->sigma, h->cfg.nu, dt_out, l.zchunk(), h->red, SLOT_CFLINT, sg, fin);
    }
    if (h->dist.on()) {
      static_assert(SLOT_CFLINT == SLOT_PHIMAX + 1, "the two CFL maxima of uniform mode are reduced together");
      if (sg == SLOT_PHIMAX)
        allreduce_slot(h, SLOT_PHIMAX, WL_NCCL_MAX, 2);
      else {
        allreduce_slot(h, SLOT_CFLINT, WL_NCCL_MAX);
        allreduce_slot(h, sg, WL_NCCL_MAX);
      }
      LAUNCH(h, k_cfl_final, 1, 1, h->red, SLOT_CFLINT, sg, h->cfg.nu, dt_out);
    }
    return;
  }
  dim3 b = blk(h->D);
  Box all = l.all();
  LAUNCH_D(h, k_cfl, grd(all, b), b, g, all, (const float*)h->u, h->sigma, h->cfg.nu, dt_out, h->red, SLOT_CFL);
}

// mom_project!(a,b,w,t)  src/Flow.jl:223-232
// `dt_cfl`: (corrector) the caller wants push!(Δt, CFL(a)) right after this projection: in uniform mode the correction kernel
// computes it on the way (f_correct_cfl) and *cfl_done is set
static int project(wl_handle* h, float w, float* dt_cfl = nullptr, bool* cfl_done = nullptr) {
  Nvtx rg(w == 1.f ? "mom_project!(w=1)" : "mom_project!(w=0.5)");
  float r2;
  TRY(residual(h, 1, w, &r2));
  TRY(solve_after_residual(h, r2, nullptr));
  Level& l = h->levels[0];
  dim3 b = blk(h->D);
  Box in = l.inside();
  if (l.fast && h->uni && dt_cfl && lazy_bc(h) && h->fuse_cfl) {
    const int fin = h->dist.on() ? 0 : 1;
    LAUNCH_PDL(h, f_correct_cfl<true>, l.fgrid(), dim3(32, FTY), 0, l.g, l.coef(true), (const float*)l.x, (const float*)h->u, h->f, h->p, dtp(h), w, l.zchunk(),
           h->cfg.nu, dt_cfl, h->red, SLOT_CFLINT, SLOT_PHIMAX, fin, h->d_flags);
    h->range_checked = true;
    std::swap(h->u, h->f);
    if (h->dist.on()) {
      allreduce_slot(h, SLOT_PHIMAX, WL_NCCL_MAX, 2);
      LAUNCH(h, k_cfl_final, 1, 1, h->red, SLOT_CFLINT, SLOT_PHIMAX, h->cfg.nu, dt_cfl);
    }
    if (cfl_done) *cfl_done = true;
  } else if (l.fast && h->uni && lazy_bc(h) && h->fuse_cfl) {  // out of place into f (free in uniform mode), then the two swap roles
    LAUNCH_PDL(h, f_correct_cfl<false>, l.fgrid(), dim3(32, FTY), 0, l.g, l.coef(true), (const float*)l.x, (const float*)h->u, h->f, h->p, dtp(h), w, l.zchunk(),
           h->cfg.nu, (float*)nullptr, h->red, SLOT_CFLINT, SLOT_PHIMAX, 0, h->d_flags);
    h->range_checked = true;
    std::swap(h->u, h->f);
  } else if (l.fast) {
    if (h->uni)
      LAUNCH(h, f_correct<true>, l.fgrid(), dim3(32, FTY), l.g, l.coef(true), (const float*)l.x, h->u, h->p, dtp(h), w, l.zchunk());
    else
      LAUNCH(h, f_correct<false>, l.fgrid(), dim3(32, FTY), l.g, l.coef(false), (const float*)l.x, h->u, h->p, dtp(h), w, l.zchunk());
  } else
    LAUNCH_D(h, k_correct, grd(in, b), b, l.dev(), in, h->u, h->p, dtp(h), w);
  step_bc(h, h->u);
  TRY(exch_u(h, h->u, h->p));
  return 0;
}

static int ensure_dt_capacity(wl_handle* h, size_t need) {
  if (need <= h->dt_cap) return 0;
  size_t cap = std::max<size_t>(need * 2, 1 << 16);
  float* q = nullptr;
  CK(cudaMalloc(&q, cap * sizeof(float)));
  if (h->d_dthist) {
    CK(cudaMemcpyAsync(q, h->d_dthist, h->dt_dev_len * sizeof(float), cudaMemcpyDeviceToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaFree(h->d_dthist));
  }
  h->d_dthist = q;
  h->dt_cap = cap;
  return 0;
}
static int check_flags(wl_handle* h) {
  int f = 0, e = 0;
  CK(cudaMemcpyAsync(&f, h->d_flags, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  int e2 = 0;
  if (h->mbox) CK(cudaMemcpyAsync(&e, h->mbox + 5, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  if (h->mbox) CK(cudaMemcpyAsync(&e2, h->mbox + 16 + 5, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  e |= e2;
  if (f) return fail("the flux kernel met a non-finite velocity (or |u| > 1e37): the velocity field has diverged");
  if (e) return fail("peer-to-peer halo exchange timed out waiting for a neighbouring rank (the ring is marked failed; results after the timeout are invalid)");
  return 0;
}
static int sync_dt(wl_handle* h) {  // mirror device Δt history to the host vector
  if (h->dt.size() < h->dt_dev_len) {
    const size_t old = h->dt.size();
    h->dt.resize(h->dt_dev_len);
    CK(cudaMemcpyAsync(h->dt.data() + old, h->d_dthist + old, (h->dt_dev_len - old) * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
  }
  return 0;
}

// mom_step!(a,b)  src/Flow.jl:156-167
// time(a) = sum(@view a.Δt[1:end-1]) (src/Flow.jl:174; all = false) and sum(a.Δt) (measure!(sim)'s default t; all = true).
// Julia's Float32 `sum` is pairwise in blocks of 1024 with a @simd inner loop: its rounding depends on the vector width (unpinned),
// but it stays within a few ulp of the exact sum for any length.  Here the sum is accumulated incrementally in double (O(1) per
// step; the running Float32 loop this replaces was O(n) per call and drifted by O(n·eps) from the reference after thousands of
// steps) and rounded to Float32 once, i.e. the Float32 nearest to the exact sum.
static int time_sum(wl_handle* h, bool all, double* t) {
  TRY(sync_dt(h));
  const size_t n = h->dt.empty() ? 0 : h->dt.size() - 1;
  if (h->tsum_n > n) h->tsum_n = 0, h->tsum = 0.0;
  for (; h->tsum_n < n; h->tsum_n++) h->tsum += (double)h->dt[h->tsum_n];
  *t = (double)(float)(all && !h->dt.empty() ? h->tsum + (double)h->dt.back() : h->tsum);
  return 0;
}
static int mom_step(wl_handle* h) {
  Nvtx rg("mom_step!");
  TRY(ensure_hierarchy(h));
  TRY(ensure_dt_capacity(h, h->dt_dev_len + 1));
  float t0 = 0.f, t1 = 0.f;
  if (h->forcing) {  // t₁ = sum(a.Δt); t₀ = t₁ − a.Δt[end]  (src/Flow.jl:157); BC! at t₁ (src/Flow.jl:194,209,230)
    double s;
    TRY(time_sum(h, true, &s));
    t1 = (float)s;
    t0 = t1 - h->dt.back();
    for (int i = 0; i < 3; i++) h->ubc_now[i] = h->cfg.uBC[i] + h->fU1[i] * t1 + (h->fU2[i] * t1) * t1 / 2.f;
  }
  auto stage_force = [&](float t) {  // accelerate!(a.f, t, a.g, a.uBC): g(i,x,t) + dU(i,x,t)/dt
    h->fc.on = h->forcing ? 1 : 0;
    for (int i = 0; i < 3; i++) h->fc.a[i] = (h->fg0[i] + h->fg1[i] * t) + (h->fU1[i] + h->fU2[i] * t);
  };
  dim3 b = blk(h->D);
  Level& l = h->levels[0];
  Box in = l.inside(), all = l.all();
  std::swap(h->u, h->u0);  // u⁰ .= u ; the new u is rebuilt from scratch below (scale_u!(a,0))
  // predictor  src/Flow.jl:190-196
  stage_force(t0);
  {
    Nvtx r1("mom_predict! (conv_diff!, BDIM!)");
    TRY(momentum(h, 0));
  }
  step_bc(h, h->u0);
  if (h->cfg.exitBC) TRY(launch_exitbc(h, h->u, h->u0, 1.f));
  TRY(lazy_bc(h) ? exch_uz_up(h, h->u) : exch_u(h, h->u));
  TRY(project(h, 1.f));
  // corrector  src/Flow.jl:205-210
  stage_force(t1);
  {
    Nvtx r2("mom_correct! (conv_diff!, BDIM!)");
    TRY(momentum(h, 1));
  }
  step_bc(h, h->u);
  TRY(lazy_bc(h) ? exch_uz_up(h, h->u) : exch_u(h, h->u));
  bool cfl_done = false;
  TRY(project(h, 0.5f, h->d_dthist + h->dt_dev_len, &cfl_done));
  // push!(a.Δt, CFL(a))
  if (!cfl_done) cfl(h, h->d_dthist + h->dt_dev_len);
  h->dt_dev_len++;
  CK(cudaGetLastError());
  return 0;
}

// ---- field table --------------------------------------------------------------------------
static int field_ptr(wl_handle* h, int field, float** p, int* ncomp) {
  const int D = h->D;
  switch (field) {
    case WL_U: *p = h->u; *ncomp = D; return 0;
    case WL_U0: *p = h->u0; *ncomp = D; return 0;
    case WL_F: *p = h->f; *ncomp = D; return 0;
    case WL_P: *p = h->p; *ncomp = 1; return 0;
    case WL_SIGMA: *p = h->sigma; *ncomp = 1; return 0;
    case WL_V: *p = h->V; *ncomp = D; return 0;
    case WL_MU0: *p = h->mu0; *ncomp = D; return 0;
    case WL_MU1: *p = h->mu1; *ncomp = D * D; return 0;
  }
  return fail("unknown field id %d", field);
}
// Host transfers go through a dense device staging buffer of one component: one contiguous PCIe copy (a 2-D copy with 2 KB rows
// runs at 3–6 GB/s, a contiguous one at the link rate), then a device-side repack between the reference layout and the pitched one.
static int stage_ensure(wl_handle* h, size_t nfloats) {
  if (h->stage_cap >= nfloats) return 0;
  if (h->stage) CK(cudaFree(h->stage));
  h->stage = nullptr;
  h->stage_cap = 0;
  CK(cudaMalloc((void**)&h->stage, nfloats * sizeof(float)));
  h->stage_cap = nfloats;
  return 0;
}
static int copy_in(wl_handle* h, const Grid& g, float* dst, const float* src, int ncomp, int src_is_device) {
  const size_t rows = (size_t)g.N[1] * g.N[2];
  const size_t dense = (size_t)g.N[0] * rows;
  if (src_is_device) {
    CK(cudaMemcpy2DAsync(dst + g.xo, (size_t)g.px * 4, src, (size_t)g.N[0] * 4, (size_t)g.N[0] * 4, rows * ncomp, cudaMemcpyDeviceToDevice, h->st));
    return 0;
  }
  TRY(stage_ensure(h, dense));
  const size_t total = dense * 4, npc = (total + PIN_BYTES - 1) / PIN_BYTES;
  const bool ring = total >= ((size_t)64 << 20) && !host_is_pinned(src) && pin_ring_init();
  bool used[PIN_NB] = {false};
  for (int c = 0; c < ncomp; c++) {
    const char* hs = (const char*)(src + (size_t)c * dense);
    if (ring) {  // host threads fill pinned pieces (PIN_NB − 1 ahead), the copy engine drains them in order
      std::future<void> fut[PIN_NB];
      auto fill = [&](size_t i) {
        const int k = (int)(i % PIN_NB);
        if (used[k]) cudaEventSynchronize(g_pin_ev[k]);  // the DMA out of this buffer's previous piece is complete
        const size_t off = i * PIN_BYTES, len = std::min(PIN_BYTES, total - off);
        char* pk = g_pin[k];
        fut[k] = std::async(std::launch::async, [=]() { memcpy(pk, hs + off, len); });
      };
      for (size_t i = 0; i < std::min<size_t>(npc, PIN_NB); i++) fill(i);
      for (size_t i = 0; i < npc; i++) {
        const int k = (int)(i % PIN_NB);
        fut[k].get();
        const size_t off = i * PIN_BYTES, len = std::min(PIN_BYTES, total - off);
        CK(cudaMemcpyAsync((char*)h->stage + off, g_pin[k], len, cudaMemcpyHostToDevice, h->st));
        CK(cudaEventRecord(g_pin_ev[k], h->st));
        used[k] = true;
        if (i + PIN_NB < npc) fill(i + PIN_NB);
      }
    } else
      CK(cudaMemcpyAsync(h->stage, hs, total, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpy2DAsync(dst + (size_t)c * g.sc + g.xo, (size_t)g.px * 4, h->stage, (size_t)g.N[0] * 4, (size_t)g.N[0] * 4, rows, cudaMemcpyDeviceToDevice, h->st));
  }
  CK(cudaStreamSynchronize(h->st));
  return 0;
}
// A freshly allocated host buffer costs a page fault per 4 KB when the driver's copy thread first writes it (0.4 s for the 1.6 GB
// of u at 512³, against 0.09 s for the copy itself): fault the pages in from several threads first.  The buffer is about to be
// overwritten completely, so writing one byte per page is harmless.
static void prefault(void* dst, size_t bytes) {
  if (bytes < ((size_t)64 << 20)) return;
  const size_t page = 4096;
  char* lo = (char*)(((uintptr_t)dst + page - 1) / page * page);
  char* hi = (char*)(((uintptr_t)dst + bytes) / page * page);
  if (hi <= lo) return;
  madvise(lo, (size_t)(hi - lo), MADV_HUGEPAGE);
  const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const size_t npages = (size_t)(hi - lo) / page;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) {
    th.emplace_back([=]() {
      const size_t a = npages * t / nt, b = npages * (t + 1) / nt;
      for (size_t q = a; q < b; q++) ((volatile char*)lo)[q * page] = 0;
    });
  }
  for (auto& x : th) x.join();
}
static int copy_out(wl_handle* h, const Grid& g, float* dst, const float* src, int ncomp, int dst_is_device) {
  const size_t rows = (size_t)g.N[1] * g.N[2];
  const size_t dense = (size_t)g.N[0] * rows;
  if (dst_is_device) {
    CK(cudaMemcpy2DAsync(dst, (size_t)g.N[0] * 4, src + g.xo, (size_t)g.px * 4, (size_t)g.N[0] * 4, rows * ncomp, cudaMemcpyDeviceToDevice, h->st));
    return 0;
  }
  TRY(stage_ensure(h, dense));
  const size_t total = dense * 4, npc = (total + PIN_BYTES - 1) / PIN_BYTES;
  const bool ring = total >= ((size_t)64 << 20) && !host_is_pinned(dst) && pin_ring_init();
  if (!ring)
    prefault(dst, total * ncomp);
  else {  // (fresh pages are first touched by the copying threads below, PIN_NB wide; ask for huge pages first)
    const size_t page = 4096;
    char* lo = (char*)(((uintptr_t)dst + page - 1) / page * page);
    char* hi = (char*)(((uintptr_t)dst + total * ncomp) / page * page);
    if (hi > lo) madvise(lo, (size_t)(hi - lo), MADV_HUGEPAGE);
  }
  std::future<void> fut[PIN_NB];
  int rc = 0;
  for (int c = 0; c < ncomp && !rc; c++) {
    cudaError_t e = cudaMemcpy2DAsync(h->stage, (size_t)g.N[0] * 4, src + (size_t)c * g.sc + g.xo, (size_t)g.px * 4, (size_t)g.N[0] * 4, rows, cudaMemcpyDeviceToDevice, h->st);
    char* hd = (char*)(dst + (size_t)c * dense);
    if (e == cudaSuccess && ring) {
      // the copy engine fills pinned pieces in order; a host thread per piece waits for its event and copies it into the caller's
      // array (first touch of fresh pages included: the page faults run PIN_NB wide instead of on the driver's one copy thread)
      for (size_t i = 0; i < npc && e == cudaSuccess; i++) {
        const int k = (int)(i % PIN_NB);
        if (fut[k].valid()) fut[k].get();
        const size_t off = i * PIN_BYTES, len = std::min(PIN_BYTES, total - off);
        e = cudaMemcpyAsync(g_pin[k], (const char*)h->stage + off, len, cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaEventRecord(g_pin_ev[k], h->st);
        if (e != cudaSuccess) break;
        char* pk = g_pin[k];
        cudaEvent_t ev = g_pin_ev[k];
        fut[k] = std::async(std::launch::async, [=]() {
          cudaEventSynchronize(ev);
          memcpy(hd + off, pk, len);
        });
      }
    } else if (e == cudaSuccess)
      e = cudaMemcpyAsync(hd, h->stage, total, cudaMemcpyDeviceToHost, h->st);
    if (e != cudaSuccess) rc = fail("download: %s", cudaGetErrorString(e));
  }
  for (int k = 0; k < PIN_NB; k++)
    if (fut[k].valid()) fut[k].get();
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->st));
  return 0;
}

// ============================================================================================
extern "C" {

const char* wl_last_error(void) { return g_err.c_str(); }

int wl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

static int create_impl(const wl_config* cfg, int rank, int nranks, const void* nccl_id, wl_handle** out);

// Map the neighbours' field chunks (CUDA IPC over NVLink).  The handles travel through an NCCL all-gather.
static int setup_p2p(wl_handle* h) {
  const Dist& d = h->dist;
  const int nch = (int)h->chunks.size();
  // all ranks must have made the same allocations
  const size_t rec = sizeof(cudaIpcMemHandle_t);
  std::vector<char> mine((size_t)nch * rec);
  for (int k = 0; k < nch; k++) {
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, h->chunks[k].base));
    memcpy(mine.data() + (size_t)k * rec, &hd, rec);
  }
  char* dbuf = nullptr;
  const size_t per = (size_t)nch * rec;
  CK(cudaMalloc((void**)&dbuf, per * d.P));
  CK(cudaMemcpyAsync(dbuf + per * d.rank, mine.data(), per, cudaMemcpyHostToDevice, h->st));
  NCK(g_nccl.AllGather(dbuf + per * d.rank, dbuf, per, 0 /*ncclInt8*/, d.comm, h->st));
  std::vector<char> all(per * d.P);
  CK(cudaMemcpyAsync(all.data(), dbuf, per * d.P, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaFree(dbuf));
  const int nb[2] = {d.down, d.up};
  for (int side = 0; side < 2; side++) {
    h->peer_base[side].assign(nch, nullptr);
    if (nb[side] < 0) continue;
    if (side == 1 && d.up == d.down) {  // two ranks, periodic: one neighbour on both sides
      h->peer_base[1] = h->peer_base[0];
      continue;
    }
    for (int k = 0; k < nch; k++) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, all.data() + per * nb[side] + (size_t)k * rec, rec);
      void* q = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail("cudaIpcOpenMemHandle (rank %d chunk %d): %s", nb[side], k, cudaGetErrorString(e));
      h->peer_base[side][k] = (char*)q;
    }
  }
  h->p2p = true;
  h->ipc_all = all;
  // every rank's all-reduce mailbox (the chunk it lives in: neighbours' mappings are reused, the others are opened here)
  if (d.P <= 8 && h->armb && !(h->cfg.flags & WL_FLAG_NCCL_ALLREDUCE)) {
    for (int q = 0; q < d.P; q++) {
      h->ar_peers.p[q] = (double*)any_peer_ptr(h, q, (const float*)h->armb);
      if (!h->ar_peers.p[q]) return 1;
    }
    h->ar_on = true;
  }
  return 0;
}

int wl_create(const wl_config* cfg, wl_handle** out) { return create_impl(cfg, 0, 1, nullptr, out); }

int wl_dist_unique_id(void* id128) {
  if (!id128) return fail("null argument");
  const char* why = "";
  if (!g_nccl.load(&why)) return fail("NCCL unavailable: %s", why);
  ncclUniqueId id;
  NCK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

int wl_create_dist(const wl_config* cfg, int rank, int nranks, const void* nccl_id128, wl_handle** out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("bad rank %d of %d", rank, nranks);
  if (nranks > 1 && !nccl_id128) return fail("null NCCL id");
  return create_impl(cfg, rank, nranks, nccl_id128, out);
}

static int create_impl(const wl_config* cfg, int rank, int nranks, const void* nccl_id, wl_handle** out) {
  if (!cfg || !out) return fail("null argument");
  if (cfg->D != 2 && cfg->D != 3) return fail("D must be 2 or 3 (got %d)", cfg->D);
  for (int d = 0; d < cfg->D; d++)
    if (cfg->n[d] < 2) return fail("dims[%d]=%d too small", d, cfg->n[d]);
  if (cfg->lambda < 0 || cfg->lambda > 2) return fail("unknown convective scheme %d", cfg->lambda);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: libwl_b200 has no CPU fallback (cudaGetDeviceCount found %d devices)", ndev);
  if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (%d devices)", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  wl_handle* h = new wl_handle();
  h->cfg = *cfg;
  h->D = cfg->D;
  h->tol = cfg->tol > 0 ? (double)cfg->tol : 1e-4;
  h->fused_gs = !(cfg->flags & WL_FLAG_UNFUSED_GS);
  // A/B switches between kernel variants that compute the same bits (debugging aids, part of the ABI: wl_config.flags)
  h->vsmooth = !(cfg->flags & WL_FLAG_NO_VSMOOTH);
  h->conv4 = !(cfg->flags & WL_FLAG_NO_CONV4);
  h->semi_on = !(cfg->flags & WL_FLAG_NO_SEMI);
  h->tiny_on = !(cfg->flags & WL_FLAG_NO_TINY);
  h->pdl = !(cfg->flags & WL_FLAG_NO_PDL);
  h->fuse_cfl = h->divres_uni = h->jacobi2 = !(cfg->flags & WL_FLAG_NO_FUSED_UNI);

  h->itmx = cfg->itmx > 0 ? cfg->itmx : (cfg->pois_kind == WL_POIS_MULTILEVEL ? 32 : 1000);
  int N[3];
  for (int d = 0; d < 3; d++) N[d] = d < cfg->D ? cfg->n[d] + 2 : 1;
  h->g = make_grid(cfg->D, N, cfg->perdir);
  h->perz_global = cfg->D == 3 && cfg->perdir[2] != 0;
  int rc = 0;
  do {
    if (nranks > 1) {
      bool ok = cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
      for (int q = 0; q < 4 && ok; q++) ok = cudaEventCreateWithFlags(&h->ev_pre[q], cudaEventDisableTiming) == cudaSuccess;
      if (!ok) {
        rc = fail("side stream / events: %s", cudaGetErrorString(cudaGetLastError()));
        break;
      }
      h->prefetch = !(cfg->flags & WL_FLAG_NO_PREFETCH);
      h->skip_ready = !(cfg->flags & WL_FLAG_NO_PREFETCH);
    }
    if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) {
      rc = fail("cudaStreamCreate failed");
      break;
    }
    if (nranks > 1) {  // z-slab decomposition: one process per GPU, NCCL communicator over all ranks
      if (cfg->D != 3) { rc = fail("z-slab decomposition needs a 3-D domain"); break; }
      if (cfg->pois_kind != WL_POIS_MULTILEVEL) { rc = fail("z-slab decomposition supports the MultiLevelPoisson solver only"); break; }
      const char* why = "";
      if (!g_nccl.load(&why)) { rc = fail("NCCL unavailable: %s", why); break; }
      // One communicator per (device, rank, size) and process: ncclCommInitRank costs 0.3–3 s (it dominated the set-up of every
      // further Simulation of a multi-GPU host program), so later handles reuse the first one's and ignore their own id.  Every
      // rank creates its handles in the same order, so all ranks take the same decision.
      const auto key = std::make_tuple(cfg->device, rank, nranks);
      auto it = g_comm_cache.find(key);
      if (it != g_comm_cache.end()) {
        h->dist.comm = it->second;
      } else {
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof id);
        int e = g_nccl.CommInitRank(&h->dist.comm, nranks, id, rank);
        if (e != 0) { rc = fail("ncclCommInitRank: %s", g_nccl.GetErrorString(e)); break; }
        g_comm_cache[key] = h->dist.comm;
      }
      h->dist.rank = rank;
      h->dist.P = nranks;
      h->dist.up = rank + 1 < nranks ? rank + 1 : (h->perz_global ? 0 : -1);
      h->dist.down = rank > 0 ? rank - 1 : (h->perz_global ? nranks - 1 : -1);
      Level tmp;
      tmp.g = h->g;
      if ((cfg->n[2] % nranks) != 0 || cfg->n[2] / nranks < 4 || (cfg->n[2] / nranks) % 2) {
        rc = fail("z-slab decomposition needs dims[3]=%d divisible by %d ranks with an even number (>=4) of planes each", cfg->n[2], nranks);
        break;
      }
      h->g = slab_grid(h, h->g);
      if ((rc = dalloc(h, &h->uext, (size_t)6 * h->g.s[2]))) break;
      {
        float* q = nullptr;
        if ((rc = dalloc(h, &q, 64))) break;
        h->mbox = (int*)q;
        if ((rc = dalloc(h, &q, 2 * 2 * 8 * 4 * 2))) break;  // k_allreduce mailbox [2][8][4] doubles, then the same for k_bcast_planes' barrier
        h->armb = (double*)q;
      }
    }
    const size_t n = (size_t)h->g.sc;
    const int D = h->D;
    if ((rc = dalloc(h, &h->u, n * D)) || (rc = dalloc(h, &h->u0, n * D)) || (rc = dalloc(h, &h->f, n * D)) || (rc = dalloc(h, &h->p, n)) ||
        (rc = dalloc(h, &h->sigma, n)) || (rc = dalloc(h, &h->V, n * D)) || (rc = dalloc(h, &h->mu0, n * D)) || (rc = dalloc(h, &h->mu1, n * D * D)))
      break;
    if (!h->uext && (rc = dalloc(h, &h->uext, 32))) break;
    {
      float* q = nullptr;
      if ((rc = dalloc(h, &q, 8))) break;
      h->d_flags = (int*)q;
    }
    if ((rc = build_levels(h))) break;
    // fields of this size move through the pinned staging ring: allocate it now (≈ 90 ms once per process), not inside the first download
    if ((size_t)h->g.N[0] * h->g.N[1] * h->g.N[2] * sizeof(float) >= ((size_t)64 << 20)) pin_ring_init();
    if (h->D == 3 && h->cfg.pois_kind == WL_POIS_MULTILEVEL && h->cfg.smoother == WL_SMOOTH_GSRB && !(cfg->flags & WL_FLAG_NO_PERSISTENT)) {
      // levels of at most ~0.6 M cells (and, with z slabs, only replicated ones) go to the persistent coarse-level kernel
      for (size_t i = 1; i < h->levels.size(); i++) {
        const Level& l = h->levels[i];
        if (!l.slab && (i64)l.g.N[0] * l.g.N[1] * l.g.N[2] <= 600000) {
          h->small_from = (int)i;
          break;
        }
      }
      if (h->small_from > 0) {
        const int nl = (int)h->levels.size();
        int t = nl;  // grow the tiny range upwards from the coarsest level while the chain stays fully coarsened, periodic and whole
        auto ok = [&](int q) {
          const Level& l = h->levels[q];
          return !l.slab && l.g.per[0] && l.g.per[1] && l.g.per[2] && (i64)l.g.N[0] * l.g.N[1] * l.g.N[2] <= 8192;
        };
        while (t - 1 >= h->small_from && t - 1 >= 1 && nl - (t - 1) <= TINY_MAXLEV && ok(t - 1) && (t == nl || h->levels[t].fullc)) t--;
        h->tiny_from = t < nl ? t : 0;
        // general mode: the innermost levels whose nine arrays (ghost cells included) fit the shared-memory arena together
        int tg = nl;
        size_t fl = 0;
        while (tg - 1 >= h->small_from && tg - 1 >= 1 && nl - (tg - 1) <= TINYG_MAXLEV && !h->levels[tg - 1].slab) {
          const Level& l = h->levels[tg - 1];
          const size_t n = (size_t)l.g.N[0] * l.g.N[1] * l.g.N[2];
          if ((fl + 9 * n) * sizeof(float) > (size_t)176 * 1024) break;  // (the operation list, ≤ 64 × sizeof(SmallOp), follows the arena)
          fl += 9 * n;
          tg--;
        }
        h->tg_from = tg < nl ? tg : 0;
      }
      if (h->small_from > 0) {
        int dev = 0, nsm = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_small_levels<true>, 32 * FTY, 0);
        int per_sm2 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_small_levels<false>, 32 * FTY, 0);
        per_sm = std::min(per_sm, per_sm2);
        if (!coop || per_sm < 1) {
          h->small_from = 0;
        } else {
          h->small_grid = nsm * std::min(per_sm, 2);
          void* q = nullptr;
          if (cudaMalloc(&q, 256 * sizeof(SmallOp)) != cudaSuccess) { rc = fail("cudaMalloc ops"); break; }
          h->d_ops = (SmallOp*)q;
          h->allocs.push_back(q);
          if (cudaMallocHost((void**)&h->h_ops, 256 * sizeof(SmallOp)) != cudaSuccess) { rc = fail("cudaMallocHost ops"); break; }
        }
      }
    }
    // reduction buffers
    Box all = h->levels[0].all();
    const size_t nb = std::max(nblocks(all, blk(D)), (size_t)1024) + 4096;
    void* q;
    if (cudaMalloc(&q, nb * 2 * sizeof(double)) != cudaSuccess) { rc = fail("cudaMalloc partials"); break; }
    h->red.partials = (double*)q;
    h->allocs.push_back(q);
    if (cudaMalloc(&q, 8192) != cudaSuccess) { rc = fail("cudaMalloc ticket"); break; }
    cudaMemsetAsync(q, 0, 8192, h->st);
    h->red.ticket = (unsigned int*)q;
    h->allocs.push_back(q);
    if (cudaMalloc(&q, NSLOTS * sizeof(double)) != cudaSuccess) { rc = fail("cudaMalloc out"); break; }
    cudaMemsetAsync(q, 0, NSLOTS * sizeof(double), h->st);
    h->red.out = (double*)q;
    h->allocs.push_back(q);
    if (cudaMallocHost((void**)&h->h_out, NSLOTS * sizeof(double)) != cudaSuccess) { rc = fail("cudaMallocHost"); break; }
    {  // mapped mirror: NSLOTS doubles, then NSLOTS sequence words (unified addressing: the host pointer is valid on the device)
      void* m = nullptr;
      if (cudaHostAlloc(&m, NSLOTS * (sizeof(double) + sizeof(unsigned int)), cudaHostAllocMapped) != cudaSuccess) { rc = fail("cudaHostAlloc (mapped)"); break; }
      memset(m, 0, NSLOTS * (sizeof(double) + sizeof(unsigned int)));
      h->red.hout = (double*)m;
      h->red.hseq = (unsigned int*)((char*)m + NSLOTS * sizeof(double));
      h->red.tag = 0;
      h->fast_read = !(cfg->flags & WL_FLAG_NO_FAST_READ);
    }
    if ((rc = dalloc(h, &h->d_scal, 16))) break;
    if ((rc = ensure_dt_capacity(h, 1024))) break;
    if (h->dist.on() && !(cfg->flags & WL_FLAG_NCCL_HALO)) {
      if (cudaStreamSynchronize(h->st) != cudaSuccess) { rc = fail("init sync"); break; }
      if ((rc = setup_p2p(h))) break;
    }
    // Flow ctor: Δt=[dt0]; u = uBC everywhere; BC!; exitBC!(u,u,0); u⁰=copy(u); μ₀=1; BC!(μ₀,0)
    h->dt.assign(1, cfg->dt0);
    if (cudaMemcpyAsync(h->d_dthist, h->dt.data(), sizeof(float), cudaMemcpyHostToDevice, h->st) != cudaSuccess) { rc = fail("memcpy dt"); break; }
    h->dt_dev_len = 1;
    set_scalar(h, 0, 1.f);
    set_scalar(h, 1, 1.f);
    for (int i = 0; i < D; i++) {
      LAUNCH(h, k_fill, 1024, 256, h->u + (size_t)i * n, n, cfg->uBC[i]);
      LAUNCH(h, k_fill, 1024, 256, h->mu0 + (size_t)i * n, n, 1.f);
    }
    const float zero[3] = {0, 0, 0};
    launch_bc_vec(h, h->g, h->mu0, zero, 0, h->mu0);
    if (cudaStreamSynchronize(h->st) != cudaSuccess) { rc = fail("init sync: %s", cudaGetErrorString(cudaGetLastError())); break; }
    if ((rc = wl_apply_bc(h))) break;
    if ((rc = update_levels(h))) break;
    if (cudaStreamSynchronize(h->st) != cudaSuccess) { rc = fail("init sync: %s", cudaGetErrorString(cudaGetLastError())); break; }
  } while (0);
  if (rc) {
    std::string keep = g_err;
    wl_destroy(h);
    g_err = keep;
    return rc;
  }
  *out = h;
  return 0;
}

int wl_destroy(wl_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->p2p) {
    for (int side = 0; side < 2; side++) {
      if (side == 1 && h->dist.up == h->dist.down) break;
      for (char* q : h->peer_base[side])
        if (q) cudaIpcCloseMemHandle(q);
    }
    for (void* q : h->ar_opened) cudaIpcCloseMemHandle(q);
  }
  // (the communicator belongs to the process-wide cache and outlives the handle)
  for (void* q : h->allocs) cudaFree(q);
  for (const Chunk& c : h->chunks) pool_give(h->cfg.device, c.size, c.base);  // (the stream was synchronised above: nothing uses them)
  if (h->stage) cudaFree(h->stage);
  if (h->force_red.partials) cudaFree(h->force_red.partials);
  if (h->force_red.out) cudaFree(h->force_red.out);
  if (h->d_dthist) cudaFree(h->d_dthist);
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->red.hout) cudaFreeHost(h->red.hout);
  if (h->h_ops) cudaFreeHost(h->h_ops);
  if (h->st2) {
    cudaStreamSynchronize(h->st2);
    cudaStreamDestroy(h->st2);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int q = 0; q < 4; q++)
    if (h->ev_pre[q]) cudaEventDestroy(h->ev_pre[q]);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

int wl_release_pool(void) {
  pool_release(-1);
  return 0;
}

int wl_upload(wl_handle* h, int field, const float* src, int src_is_device) {
  if (!h || !src) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  float* p;
  int nc;
  TRY(field_ptr(h, field, &p, &nc));
  if (field == WL_V || field == WL_MU0 || field == WL_MU1) h->nobody_valid = false, h->pois_dirty = true;  // rebuilt by wl_update or lazily
  if (field == WL_U || field == WL_U0) h->range_checked = false;
  return copy_in(h, h->g, p, src, nc, src_is_device);  // z slabs: the caller's slab carries its own ghost planes
}

int wl_upload_component(wl_handle* h, int field, int comp, const float* src, int src_is_device) {
  if (!h || !src) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  float* p;
  int nc;
  TRY(field_ptr(h, field, &p, &nc));
  if (comp < 0 || comp >= nc) return fail("component %d out of range (field has %d)", comp, nc);
  if (field == WL_V || field == WL_MU0 || field == WL_MU1) h->nobody_valid = false, h->pois_dirty = true;
  if (field == WL_U || field == WL_U0) h->range_checked = false;
  return copy_in(h, h->g, p + (size_t)comp * h->g.sc, src, 1, src_is_device);
}

int wl_download(wl_handle* h, int field, float* dst, int dst_is_device) {
  if (!h || !dst) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  float* p;
  int nc;
  TRY(field_ptr(h, field, &p, &nc));
  if (field == WL_P) launch_perbc(h, h->g, h->p);  // perBC!(p.x) at the end of solver! (ghosts are materialised lazily)
  TRY(copy_out(h, h->g, dst, p, nc, dst_is_device));
  return dst_is_device ? 0 : check_flags(h);
}

int wl_apply_bc(wl_handle* h) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  h->range_checked = false;
  launch_bc_vec(h, h->g, h->u, h->cfg.uBC, h->cfg.exitBC, h->u);
  TRY(launch_exitbc(h, h->u, h->u, 0.f));
  TRY(exch_u(h, h->u));
  CK(cudaMemcpyAsync(h->u0, h->u, (size_t)h->g.sc * h->D * sizeof(float), cudaMemcpyDeviceToDevice, h->st));
  CK(cudaGetLastError());
  return 0;
}

int wl_measure_bc(wl_handle* h) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  const float zero[3] = {0, 0, 0};
  launch_bc_vec(h, h->g, h->mu0, zero, 0, h->mu0);
  launch_bc_vec(h, h->g, h->V, zero, h->cfg.exitBC, h->V);
  TRY(exch(h, h->levels[0], h->mu0, h->D));
  TRY(exch(h, h->levels[0], h->V, h->D));
  TRY(exch(h, h->levels[0], h->mu1, h->D * h->D));
  CK(cudaGetLastError());
  return 0;
}

int wl_update(wl_handle* h) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  return update_levels(h);
}

// measure!(flow, body; t, ϵ) for the registered body (src/Body.jl:28-51)
static int measure_body(wl_handle* h, float t) {
  if (h->body.np <= 0) return fail("wl_measure: no body registered (wl_set_body)");
  flush_ghosts(h);
  dim3 b = h->D == 3 ? dim3(32, 4, 2) : dim3(32, 8, 1);
  Box in = h->levels[0].inside();
  LAUNCH_D(h, k_measure, grd(in, b), b, h->g, in, h->body, h->body_eps, t, h->sigma, h->V, h->mu0, h->mu1);
  h->nobody_valid = false;
  h->pois_dirty = true;
  return wl_measure_bc(h);  // BC!(μ₀,0,false,perdir); BC!(V,0,exitBC,perdir) (+ z-slab halo planes)
}
// sim_step!(sim; remeasure) (src/WaterLily.jl:136-139)
static int sim_step(wl_handle* h) {
  if (h->remeasure && h->body.np > 0) {
    double t;
    TRY(time_sum(h, true, &t));  // measure!(sim, t = sum(sim.flow.Δt)) (src/WaterLily.jl:146-149)
    TRY(measure_body(h, (float)t));
    TRY(update_levels(h));
  }
  return mom_step(h);
}

int wl_set_body(wl_handle* h, const wl_body_prim* prims, int nprims, float eps) {
  if (!h) return fail("null handle");
  if (nprims < 0 || nprims > 8) return fail("wl_set_body: %d primitives (0 … 8 supported)", nprims);
  if (nprims > 0 && !prims) return fail("null argument");
  static_assert(sizeof(wl_body_prim) == sizeof(BodyPrim), "wl_body_prim and BodyPrim must have the same layout");
  memset(&h->body, 0, sizeof h->body);
  for (int q = 0; q < nprims; q++) {
    if (prims[q].kind != WL_BODY_SPHERE && prims[q].kind != WL_BODY_TORUS) return fail("wl_set_body: unknown primitive kind %d", prims[q].kind);
    if (prims[q].kind == WL_BODY_TORUS && h->D != 3) return fail("wl_set_body: the torus is a 3-D body");
    if (prims[q].op < 0 || prims[q].op > 2) return fail("wl_set_body: unknown set operation %d", prims[q].op);
    memcpy(&h->body.p[q], &prims[q], sizeof(BodyPrim));
  }
  h->body.np = nprims;
  h->body_eps = eps;
  return 0;
}

int wl_measure(wl_handle* h, float t) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  return measure_body(h, t);
}

int wl_body_forces(wl_handle* h, const float* x0, double* out12) {
  if (!h || !out12) return fail("null argument");
  if (h->body.np <= 0) return fail("wl_body_forces: no body registered (wl_set_body)");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  dim3 b = h->D == 3 ? dim3(32, 4, 2) : dim3(32, 8, 1);
  Box in = h->levels[0].inside();
  const dim3 gr = grd(in, b);
  const size_t need = (size_t)12 * ((size_t)gr.x * gr.y * gr.z + gr.z + 1);
  if (!h->force_red.partials || h->force_cap < need) {  // own partial sums: 12 values per block
    if (h->force_red.partials) CK(cudaFree(h->force_red.partials));
    if (!h->force_red.out) CK(cudaMalloc((void**)&h->force_red.out, 16 * sizeof(double)));
    CK(cudaMalloc((void**)&h->force_red.partials, need * sizeof(double)));
    h->force_cap = need;
    h->force_red.ticket = h->red.ticket;
  }
  double t;
  TRY(time_sum(h, false, &t));  // time(flow)
  const float z[3] = {0.f, 0.f, 0.f};
  const float* c = x0 ? x0 : z;
  LAUNCH_D(h, k_forces, gr, b, h->g, in, h->body, (float)t, h->cfg.nu, c[0], h->D > 1 ? c[1] : 0.f, h->D > 2 ? c[2] : 0.f, (const float*)h->u,
           (const float*)h->p, h->force_red, 0);
  if (h->dist.on()) {
    prof_begin(h, "allreduce");
    NCK(g_nccl.AllReduce(h->force_red.out, h->force_red.out, 12, WL_NCCL_DOUBLE, WL_NCCL_SUM, h->dist.comm, h->st));
    prof_end(h);
  }
  CK(cudaMemcpyAsync(out12, h->force_red.out, 12 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return 0;
}

int wl_meanflow_init(wl_handle* h, int uu_stats) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  if (h->mfP && (uu_stats != 0) != h->mf_uu && uu_stats) {
    if (!h->mfUU) TRY(dalloc(h, &h->mfUU, (size_t)h->g.sc * h->D * h->D));
  }
  if (!h->mfP) {
    TRY(dalloc(h, &h->mfP, (size_t)h->g.sc));
    TRY(dalloc(h, &h->mfU, (size_t)h->g.sc * h->D));
    if (uu_stats) TRY(dalloc(h, &h->mfUU, (size_t)h->g.sc * h->D * h->D));
  }
  h->mf_uu = uu_stats != 0;
  double t;
  TRY(time_sum(h, false, &t));
  return wl_meanflow_reset(h, (float)t);
}

int wl_meanflow_reset(wl_handle* h, float t_init) {
  if (!h) return fail("null handle");
  if (!h->mfP) return fail("no MeanFlow (wl_meanflow_init)");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemsetAsync(h->mfP, 0, (size_t)h->g.sc * sizeof(float), h->st));
  CK(cudaMemsetAsync(h->mfU, 0, (size_t)h->g.sc * h->D * sizeof(float), h->st));
  if (h->mfUU) CK(cudaMemsetAsync(h->mfUU, 0, (size_t)h->g.sc * h->D * h->D * sizeof(float), h->st));
  h->mft.assign(1, t_init);
  return 0;
}

int wl_meanflow_update(wl_handle* h) {
  if (!h) return fail("null handle");
  if (!h->mfP) return fail("no MeanFlow (wl_meanflow_init)");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  launch_perbc(h, h->g, h->p);  // the reference's solver leaves p with its periodic ghosts filled
  double tf;
  TRY(time_sum(h, false, &tf));
  const float dt = (float)tf - h->mft.back();
  const float tm = h->mft.back() - h->mft.front();  // time(meanflow)
  float eps = dt / (dt + tm + 1.1920929e-7f);
  if (h->mft.size() == 1) eps = 1.f;
  LAUNCH(h, k_meanflow, 1184, 256, (long long)h->g.sc, h->D, h->mf_uu ? 1 : 0, eps, (const float*)h->p, (const float*)h->u, h->mfP, h->mfU, h->mfUU);
  h->mft.push_back(h->mft.back() + dt);
  CK(cudaGetLastError());
  return 0;
}

int wl_meanflow_copy_to_flow(wl_handle* h) {
  if (!h) return fail("null handle");
  if (!h->mfP) return fail("no MeanFlow (wl_meanflow_init)");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  h->range_checked = false;
  CK(cudaMemcpyAsync(h->u, h->mfU, (size_t)h->g.sc * h->D * sizeof(float), cudaMemcpyDeviceToDevice, h->st));
  CK(cudaMemcpyAsync(h->p, h->mfP, (size_t)h->g.sc * sizeof(float), cudaMemcpyDeviceToDevice, h->st));
  return 0;
}

static int meanflow_ptr(wl_handle* h, int which, float** p, int* nc) {
  if (!h->mfP) return fail("no MeanFlow (wl_meanflow_init)");
  if (which == 0) { *p = h->mfP; *nc = 1; return 0; }
  if (which == 1) { *p = h->mfU; *nc = h->D; return 0; }
  if (which == 2) {
    if (!h->mfUU) return fail("this MeanFlow was created without uu_stats");
    *p = h->mfUU; *nc = h->D * h->D; return 0;
  }
  return fail("unknown MeanFlow array %d", which);
}
int wl_meanflow_download(wl_handle* h, int which, float* dst, int dst_is_device) {
  if (!h || !dst) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  float* p;
  int nc;
  TRY(meanflow_ptr(h, which, &p, &nc));
  return copy_out(h, h->g, dst, p, nc, dst_is_device);
}
int wl_meanflow_upload(wl_handle* h, int which, const float* src, int src_is_device) {
  if (!h || !src) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  float* p;
  int nc;
  TRY(meanflow_ptr(h, which, &p, &nc));
  return copy_in(h, h->g, p, src, nc, src_is_device);
}
int wl_meanflow_get_times(wl_handle* h, float* buf, int* len) {
  if (!h || !len) return fail("null argument");
  if (buf) std::copy(h->mft.begin(), h->mft.begin() + std::min<int>(*len, (int)h->mft.size()), buf);
  *len = (int)h->mft.size();
  return 0;
}
int wl_meanflow_set_times(wl_handle* h, const float* buf, int len) {
  if (!h || !buf || len < 1) return fail("bad argument");
  h->mft.assign(buf, buf + len);
  return 0;
}

int wl_set_forcing(wl_handle* h, const float* g0, const float* g1, const float* U1, const float* U2) {
  if (!h) return fail("null handle");
  const float* src[4] = {g0, g1, U1, U2};
  float* dst[4] = {h->fg0, h->fg1, h->fU1, h->fU2};
  bool any = false;
  for (int q = 0; q < 4; q++)
    for (int i = 0; i < 3; i++) {
      dst[q][i] = (src[q] && i < h->D) ? src[q][i] : 0.f;
      any = any || dst[q][i] != 0.f;
    }
  h->forcing = any;
  for (int i = 0; i < 3; i++) h->ubc_now[i] = h->cfg.uBC[i];
  return 0;
}

int wl_set_sgs(wl_handle* h, float Cs, float Delta) {
  if (!h) return fail("null handle");
  const float c = Cs * Delta;
  const bool on = c != 0.f;
  if (on && h->dist.on()) return fail("the built-in sgs! udf is not available with z-slab decomposition");
  if (on && !h->nut) TRY(dalloc(h, &h->nut, (size_t)h->g.sc));
  if (on != h->sgs_on) {
    flush_ghosts(h);       // (uniform mode defers BC!(u); the general kernels read the ghost cells)
    h->pois_dirty = true;  // the uniform-coefficient kernels are re-decided at the next step (update_levels)
  }
  h->sgs_on = on;
  h->sgs_c2 = c * c;
  return 0;
}

int wl_set_remeasure(wl_handle* h, int enabled) {
  if (!h) return fail("null handle");
  h->remeasure = enabled != 0;
  return 0;
}

int wl_time_next(wl_handle* h, double* t) {
  if (!h || !t) return fail("null argument");
  return time_sum(h, true, t);
}

int wl_mom_step(wl_handle* h) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  TRY(sim_step(h));
  return check_flags(h);
}

int wl_sim_step_n(wl_handle* h, int nsteps) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  for (int i = 0; i < nsteps; i++) {
    TRY(sim_step(h));
    if ((i & 63) == 63) TRY(check_flags(h));  // a diverged field or a dead peer stops the loop within 64 steps
  }
  return check_flags(h);
}

int wl_time(wl_handle* h, double* t) {
  if (!h || !t) return fail("null argument");
  return time_sum(h, false, t);
}

int wl_sim_step_until(wl_handle* h, double t_end, double U, double L, int64_t max_steps, int64_t* steps_taken) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  int64_t k = 0;
  for (;;) {
    double t;
    TRY(wl_time(h, &t));
    if (!(t * U / L < t_end) || k >= max_steps) break;
    TRY(sim_step(h));
    TRY(check_flags(h));
    k++;
  }
  if (steps_taken) *steps_taken = k;
  return 0;
}

int wl_project(wl_handle* h, float w) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  TRY(ensure_hierarchy(h));
  return project(h, w);
}

int wl_conv_diff(wl_handle* h, int from_u0) {
  if (!h) return fail("null handle");
  if (h->dist.on()) return fail("standalone conv_diff! is not available with z-slab decomposition");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  conv_bdim1(h, from_u0 ? h->u0 : h->u, 0);
  CK(cudaGetLastError());
  return 0;
}

int wl_cfl(wl_handle* h, float* dt_out) {
  if (!h || !dt_out) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  flush_ghosts(h);
  cfl(h, h->d_scal + 4);
  CK(cudaMemcpyAsync(dt_out, h->d_scal + 4, sizeof(float), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return 0;
}

int wl_pois_mult(wl_handle* h) {
  if (!h) return fail("null handle");
  if (h->dist.on()) return fail("standalone mult! is not available with z-slab decomposition");
  CK(cudaSetDevice(h->cfg.device));
  TRY(ensure_hierarchy(h));
  Level& l = h->levels[0];
  CK(cudaMemsetAsync(l.z, 0, l.cells() * sizeof(float), h->st));
  dim3 b = blk(h->D);
  Box in = l.inside();
  LAUNCH_D(h, k_mult, grd(in, b), b, l.dev(), in, (const float*)h->p, l.z);
  CK(cudaGetLastError());
  return 0;
}

static int load_x_from_p(wl_handle* h) {
  Level& l = h->levels[0];
  CK(cudaMemcpyAsync(l.x, h->p, l.cells() * sizeof(float), cudaMemcpyDeviceToDevice, h->st));
  return 0;
}
static int store_x_to_p(wl_handle* h) {
  Level& l = h->levels[0];
  CK(cudaMemcpyAsync(h->p, l.x, l.cells() * sizeof(float), cudaMemcpyDeviceToDevice, h->st));
  return 0;
}

int wl_pois_residual(wl_handle* h, float* r2_out) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  TRY(ensure_hierarchy(h));
  TRY(load_x_from_p(h));
  float r2;
  TRY(residual(h, 0, 1.f, &r2));
  if (r2_out) *r2_out = r2;
  return 0;
}

int wl_pois_solve(wl_handle* h, int* iters_out) {
  if (!h) return fail("null handle");
  if (h->dist.on()) return fail("standalone solver! is not available with z-slab decomposition");
  CK(cudaSetDevice(h->cfg.device));
  TRY(ensure_hierarchy(h));
  TRY(load_x_from_p(h));
  float r2;
  TRY(residual(h, 0, 1.f, &r2));
  TRY(solve_after_residual(h, r2, iters_out));
  TRY(store_x_to_p(h));
  launch_perbc(h, h->g, h->p);
  CK(cudaGetLastError());
  return 0;
}

int wl_pois_smooth(wl_handle* h, int level, int kind, float omega) {
  if (!h) return fail("null handle");
  if (level < 0 || level >= (int)h->levels.size()) return fail("level %d out of range", level);
  CK(cudaSetDevice(h->cfg.device));
  TRY(ensure_hierarchy(h));
  Level& l = h->levels[level];
  set_scalar(h, 0, omega);
  if (kind == 0)
    TRY(gs_smooth(h, l, h->d_scal + 0, 0, 0));
  else if (kind == 1)
    TRY(jacobi(h, l, 0, nullptr, nullptr));  // reference Jacobi!(p) default ω=1
  else
    TRY(pcg(h, l));
  CK(cudaGetLastError());
  return 0;
}

int wl_pois_vcycle(wl_handle* h, float omega) {
  if (!h) return fail("null handle");
  if (h->levels.size() < 2) return fail("single-level Poisson has no V-cycle");
  CK(cudaSetDevice(h->cfg.device));
  TRY(ensure_hierarchy(h));
  set_scalar(h, 0, omega);
  TRY(vcycle(h, 0, h->d_scal + 0));
  CK(cudaGetLastError());
  return 0;
}

int wl_num_levels(wl_handle* h, int* nlevels) {
  if (!h || !nlevels) return fail("null argument");
  *nlevels = (int)h->levels.size();
  return 0;
}

int wl_level_dims(wl_handle* h, int level, int32_t* N) {
  if (!h || !N) return fail("null argument");
  if (level < 0 || level >= (int)h->levels.size()) return fail("level %d out of range", level);
  for (int d = 0; d < 3; d++) N[d] = h->levels[level].g.N[d];
  return 0;
}

static int level_ptr(wl_handle* h, int level, int which, float** p, int* nc) {
  if (level < 0 || level >= (int)h->levels.size()) return fail("level %d out of range", level);
  Level& l = h->levels[level];
  *nc = 1;
  switch (which) {
    case WL_LVL_L: *p = l.L; *nc = h->D; return 0;
    case WL_LVL_D: *p = l.Dg; return 0;
    case WL_LVL_ID: *p = l.iD; return 0;
    case WL_LVL_X: *p = l.x; return 0;
    case WL_LVL_EPS: *p = l.eps; return 0;
    case WL_LVL_R: *p = l.r; return 0;
    case WL_LVL_Z: *p = l.z; return 0;
  }
  return fail("unknown level array %d", which);
}
int wl_download_level(wl_handle* h, int level, int which, float* dst) {
  if (!h || !dst) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  float* p;
  int nc;
  TRY(level_ptr(h, level, which, &p, &nc));
  if (which == WL_LVL_L || which == WL_LVL_D || which == WL_LVL_ID) TRY(ensure_hierarchy(h));
  return copy_out(h, h->levels[level].g, dst, p, nc, 0);
}
int wl_upload_level(wl_handle* h, int level, int which, const float* src) {
  if (!h || !src) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  float* p;
  int nc;
  TRY(level_ptr(h, level, which, &p, &nc));
  if (which == WL_LVL_L && level == 0) h->nobody_valid = false, h->pois_dirty = true;
  return copy_in(h, h->levels[level].g, p, src, nc, 0);
}

int wl_get_dt(wl_handle* h, float* buf, int* len) {
  if (!h || !len) return fail("null argument");
  TRY(sync_dt(h));
  if (buf) {
    const int n = std::min<int>(*len, (int)h->dt.size());
    std::copy(h->dt.begin(), h->dt.begin() + n, buf);
  }
  *len = (int)h->dt.size();
  return 0;
}

int wl_set_dt(wl_handle* h, const float* buf, int len) {
  if (!h || !buf || len < 1) return fail("bad argument");
  CK(cudaSetDevice(h->cfg.device));
  TRY(ensure_dt_capacity(h, (size_t)len + 1));
  h->dt.assign(buf, buf + len);
  h->tsum_n = 0, h->tsum = 0.0;
  CK(cudaMemcpyAsync(h->d_dthist, h->dt.data(), len * sizeof(float), cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  h->dt_dev_len = len;
  return 0;
}

int wl_get_iters(wl_handle* h, int16_t* buf, int* len) {
  if (!h || !len) return fail("null argument");
  if (buf) {
    const int n = std::min<int>(*len, (int)h->iters.size());
    std::copy(h->iters.begin(), h->iters.begin() + n, buf);
  }
  *len = (int)h->iters.size();
  return 0;
}

int wl_get_solver_log(wl_handle* h, float* buf, int* rows) {
  if (!h || !rows) return fail("null argument");
  const int have = (int)h->log.size() / 4;
  if (buf) {
    const int n = std::min(*rows, have);
    std::copy(h->log.begin(), h->log.begin() + 4 * n, buf);
  }
  *rows = have;
  return 0;
}

int wl_set_logging(wl_handle* h, int enabled) {
  if (!h) return fail("null handle");
  h->logging = enabled != 0;
  return 0;
}

int wl_sync(wl_handle* h) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaGetLastError());
  return check_flags(h);
}

int wl_launch_count(wl_handle* h, int64_t* count) {
  if (!h || !count) return fail("null argument");
  *count = h->launches;
  return 0;
}

int wl_selftest_div6(uint64_t* nbad) {
  if (!nbad) return fail("null argument");
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, sizeof(unsigned long long)));
  CK(cudaMemset(d, 0, sizeof(unsigned long long)));
  k_selftest_div6<<<148 * 16, 256>>>(d);
  unsigned long long v = 0;
  CK(cudaMemcpy(&v, d, sizeof v, cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
  *nbad = v;
  return 0;
}

int wl_set_profiling(wl_handle* h, int enabled) {
  if (!h) return fail("null handle");
  h->prof_cells = (double)h->g.N[0] * h->g.N[1] * h->g.N[2];
  h->prof = enabled != 0;
  return 0;
}

// Text table "kernel launches total_ms\n" of everything recorded since the last call; clears the records.
int wl_get_timings(wl_handle* h, char* buf, int* len) {
  if (!h || !len) return fail("null argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->st));
  std::vector<std::string> names;
  std::vector<double> ms, cells;
  std::vector<long> cnt;
  for (auto& r : h->prof_recs) {
    float t = 0;
    cudaEventElapsedTime(&t, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    std::string nm(r.name);
    if (!nm.empty() && nm[0] == '(') nm = nm.substr(1);
    const size_t lt = nm.find('<');
    if (lt != std::string::npos) nm = nm.substr(0, lt);
    size_t k = 0;
    for (; k < names.size(); k++)
      if (names[k] == nm) break;
    if (k == names.size()) {
      names.push_back(nm);
      ms.push_back(0);
      cells.push_back(0);
      cnt.push_back(0);
    }
    ms[k] += t;
    cells[k] += r.cells;
    cnt[k]++;
  }
  h->prof_recs.clear();
  std::string& out = h->prof_text;
  char line[256];
  for (size_t k = 0; k < names.size(); k++) {
    snprintf(line, sizeof line, "%s %ld %.6f %.0f\n", names[k].c_str(), cnt[k], ms[k], cells[k]);
    out += line;
  }
  const int need = (int)out.size() + 1;
  if (buf) {
    const int n = std::min<int>(*len - 1, (int)out.size());
    if (n >= 0) {
      memcpy(buf, out.data(), n);
      buf[n] = 0;
    }
    out.clear();
  }
  *len = need;
  return 0;
}

int wl_set_tuning(wl_handle* h, const char* key, int value) {
  if (!h || !key) return fail("null argument");
  if (!strcmp(key, "vs_nz")) h->vs_nz = value;
  else if (!strcmp(key, "conv4_zchunk")) h->conv4_zchunk = std::max(1, value);
  else return fail("unknown tuning key '%s'", key);
  return 0;
}

int wl_is_const_coeff(wl_handle* h, int* flag) {
  if (!h || !flag) return fail("null argument");
  *flag = h->uni ? 1 : 0;
  return 0;
}

int wl_stream(wl_handle* h, void** stream) {
  if (!h || !stream) return fail("null argument");
  *stream = (void*)h->st;
  return 0;
}

}  // extern "C"
