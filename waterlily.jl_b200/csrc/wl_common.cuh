// wl_common.cuh — grid descriptor, launch geometry and the deterministic single-pass
// grid reduction shared by every kernel of libwl_b200.
//
// Internal field layout (library-owned): 0-based cell (i,j,k) of a
// ghost-padded N0×N1×N2 scalar lives at  xo + i + px*(j + N1*k).  xo = 31 puts the first
// INTERIOR cell (i=1) of every row on a 128-byte line (allocations are 256-byte aligned and
// the x-pitch px is a multiple of 32 floats), so warps that own 4 interior cells per lane
// issue perfectly aligned 16-byte vector loads; px leaves room for one vector past the
// upper ghost.  Vector component c adds c*sc.  2-D fields have N2 = 1.  All linear offsets
// are 64-bit (1026³ > 2³¹).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef long long i64;

struct Grid {
  int D;
  int N[3];    // cells incl. ghosts (N[2]==1 in 2-D)
  int px;      // x pitch in floats
  int xo;      // storage offset of cell i=0 inside a row
  i64 s[3];    // strides in floats: {1, px, px*N1}
  i64 sc;      // component stride = px*N1*N2
  int per[3];  // periodic flags (LOCAL wrap: in z only when the rank owns the whole periodic extent)
  // z-slab decomposition (multi-GPU): zopen[side] = the ghost plane of that z face holds the neighbouring rank's interior plane
  // (refreshed by halo exchange) instead of a wall / local periodic image; zstale[side] = that face is the GLOBAL periodic
  // boundary, across which GaussSeidelRB! sweeps must see the stale r·iD; zoff = global z index of local plane 0.
  int zopen[2];
  int zstale[2];
  int zoff;
};

// accelerate!(r,t,g,U) (src/Flow.jl:64-73) for forcings that are uniform in space: r[I,i] += a[i] on every cell of r, with
// a[i] = g_i(t) + dU_i/dt(t) evaluated by the host for the stage's time.  on = 0: no forcing (r is left untouched, bit for bit).
struct Force {
  int on;
  float a[3];
};

struct Box {
  int lo[3];
  int n[3];
};

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// `vb` is the block index a kernel body works for: blockIdx for an ordinary launch, a virtual block index when the persistent
// coarse-level kernel (k_small_levels) loops the same body over the blocks of a small grid.
__device__ __forceinline__ int3 real_block() { return make_int3(blockIdx.x, blockIdx.y, blockIdx.z); }

template <int D>
__device__ __forceinline__ bool thread_cell(const Box& b, int I[3], const int3 vb) {
  I[0] = b.lo[0] + vb.x * blockDim.x + threadIdx.x;
  I[1] = b.lo[1] + vb.y * blockDim.y + threadIdx.y;
  I[2] = (D == 3) ? b.lo[2] + vb.z * blockDim.z + threadIdx.z : 0;
  bool ok = I[0] < b.lo[0] + b.n[0] && I[1] < b.lo[1] + b.n[1];
  if (D == 3) ok = ok && I[2] < b.lo[2] + b.n[2];
  return ok;
}
template <int D>
__device__ __forceinline__ bool thread_cell(const Box& b, int I[3]) {
  return thread_cell<D>(b, I, real_block());
}

__device__ __forceinline__ i64 cell_off(const Grid& g, const int I[3]) { return (i64)(g.xo + I[0]) + g.s[1] * I[1] + g.s[2] * I[2]; }

// Offsets to the lower / upper neighbour of an INTERIOR cell in dimension d for the scalar
// Poisson fields.  Periodic dimensions wrap to the opposite interior cell, which is what
// the reference's perBC! ghost fill (src/core.jl:239-243) makes the ghost hold.
template <int D>
__device__ __forceinline__ void nbr_offsets(const Grid& g, const int I[3], i64 lo[3], i64 hi[3]) {
#pragma unroll
  for (int d = 0; d < D; d++) {
    lo[d] = (g.per[d] && I[d] == 1) ? (i64)(g.N[d] - 3) * g.s[d] : -g.s[d];
    hi[d] = (g.per[d] && I[d] == g.N[d] - 2) ? -(i64)(g.N[d] - 3) * g.s[d] : g.s[d];
  }
}

// ---- packed Float32 pairs (sm_100 FADD2 / FMUL2 / FFMA2) ---------------------------------------------------------------
// Blackwell issues one instruction for two independent IEEE round-to-nearest Float32 operations on a 64-bit register pair
// (PTX add/sub/mul/fma.rn.f32x2).  Each half is the ordinary scalar operation — same bits as FADD/FMUL/FFMA — so kernels that are
// bound by instruction issue (the flux kernel: one thread owns four x cells = two pairs) halve their FP32 instruction count
// without giving up the reference's association order.
// CONTRACTION HAZARD: ptxas 12.9 fuses a packed multiplication that feeds a packed addition / subtraction into FFMA2 even for
// .rn operands and under --fmad=false (it honours the flag for scalar code only) — one rounding instead of two, different bits.
// Rule for every kernel of this library: a packed add2 / sub2 NEVER takes a packed product as an operand; such sums are formed
// with add2x / sub2x (two scalar FADD on the halves, which ptxas leaves alone).  fma2 is used only where the scalar code already
// used __fmaf_rn (the proven two-FMA x/6).
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(const float2 a, const float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
// exact (uncontracted) sum / difference when an operand is a packed product
__device__ __forceinline__ float2 add2x(const float2 a, const float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2x(const float2 a, const float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat2(const float2 a, const float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

// ---- range of a velocity field (range_note) -------------------------------------------------------------------------------
// The uniform-mode flux kernel (fm_conv4) divides by 6 with two FMAs, exact when every velocity it reads is 0 or 2^-77 ≤ |v| ≤ 1e37
// (wl_conv4.cuh).  Instead of testing every flux input, the kernel that WRITES a velocity field tests what it writes (it is
// HBM-bound and has the issue slots to spare) and raises flags[2] when a value falls outside; fm_conv4 then runs its IEEE-division
// instance on that field.  |v| > 1e37, Inf and NaN — a diverged run — also raise the error word flags[0], which the host reports.
// Unsigned integer min / max of the magnitude bits: 0 maps to 0xffffffff in the min (b − 1), so zeros never count as small.
#define WL_RANGE_LO 0x19000000u  // bits of 2^-77
#define WL_RANGE_HI 0x7cf0bdc2u  // bits of 1e37
struct RangeAcc {
  unsigned mn = 0xffffffffu, mx = 0u;
  __device__ __forceinline__ void add(float v) {
    const unsigned b = __float_as_uint(v) & 0x7fffffffu;
    mn = min(mn, b - 1u);
    mx = max(mx, b);
  }
  __device__ __forceinline__ void add4(const float4& v) {
    add(v.x);
    add(v.y);
    add(v.z);
    add(v.w);
  }
  __device__ __forceinline__ void publish(int* flags) const {
    const bool small = mn < WL_RANGE_LO - 1u, big = mx > WL_RANGE_HI;
    if (small || big) flags[2] = 1;  // (benign race: every writer stores the same value)
    if (big) flags[0] = 1;
  }
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may be scheduled
// while its predecessor in the stream drains; it must not touch the predecessor's results before this wait (a no-op for an
// ordinary launch).  The uniform-mode step launches its kernels this way (pdl_launch, wl_b200.cu): the launch latency and the
// block ramp-up of kernel n+1 overlap the tail of kernel n.
// The trigger right behind the wait lets the NEXT kernel's blocks be scheduled as soon as every block of this one has started, i.e.
// into the free slots of this kernel's last wave, where they sit in their own wait until this grid has completed and flushed.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- deterministic single-pass grid reduction ------------------------------------------
// Every block reduces its values in double with a fixed shuffle tree, writes one partial
// per value, and the last block to arrive (atomic ticket) folds all partials in a fixed
// order and stores the results to out[slot0 .. slot0+NV).  Bit-reproducible for a fixed
// launch geometry; no host round trip.
struct RedBuf {
  double* partials;      // capacity >= NV * (number of blocks + gridDim.z)
  unsigned int* ticket;  // [0] = top ticket, [1+z] = per-layer tickets; all zero between launches
  double* out;           // result slots
  // host-visible mirror (mapped pinned memory; null = none): when `tag` is non-zero the folding thread also stores the results to
  // hout[slot…] and then, after a system-wide fence, `tag` to hseq[slot0] — the host polls that word instead of paying a copy and a
  // stream synchronisation for eight bytes (read_slot, wl_b200.cu)
  double* hout;
  unsigned int* hseq;
  unsigned int tag;
};

enum { RED_SUM = 0, RED_MAX = 1 };

template <int OP>
__device__ __forceinline__ double red_op(double a, double b) {
  return OP == RED_SUM ? a + b : (a > b ? a : b);
}
template <int OP>
__device__ __forceinline__ double red_identity() {
  return OP == RED_SUM ? 0.0 : -1.0e300;
}

template <int OP>
__device__ __forceinline__ double warp_reduce(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_op<OP>(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// Returns true in every thread of the block that performed the final fold (the last block),
// after out[] has been written; `fin` then holds the folded values in thread 0.
// Two levels keep the serial tail short on big grids: the last block of every blockIdx.z
// layer folds that layer's partials, the last layer folds the layer results.
template <int OP, int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], const RedBuf& R, int slot0, double (&fin)[NV]) {
  __shared__ double sm[NV][32];
  __shared__ int last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
  const unsigned int nlayer = gridDim.x * gridDim.y;  // blocks per layer
  const unsigned int nz = gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * blockIdx.y;
  double* lay = R.partials;                               // [NV][nz][nlayer]
  double* top = R.partials + (size_t)NV * nz * nlayer;    // [NV][nz]
  unsigned int* tick_layer = R.ticket + 1 + blockIdx.z;   // one ticket per layer
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double w = warp_reduce<OP>(v[q]);
    if (lane == 0) sm[q][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; q++) {
      double w = lane < nwarps ? sm[q][lane] : red_identity<OP>();
      w = warp_reduce<OP>(w);
      if (lane == 0) lay[((size_t)q * nz + blockIdx.z) * nlayer + bid] = w;
    }
    if (lane == 0) {
      __threadfence();
      last = (atomicAdd(tick_layer, 1u) == nlayer - 1) ? 1 : 0;
    }
  }
  __syncthreads();
  if (!last) return false;
  // ---- fold this layer ----
  __threadfence();
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double w = red_identity<OP>();
    const double* src = lay + ((size_t)q * nz + blockIdx.z) * nlayer;
    for (unsigned int b = tid; b < nlayer; b += nthreads) w = red_op<OP>(w, __ldcg(src + b));
    w = warp_reduce<OP>(w);
    __syncthreads();
    if (lane == 0) sm[q][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; q++) {
      double w = lane < nwarps ? sm[q][lane] : red_identity<OP>();
      w = warp_reduce<OP>(w);
      if (lane == 0) top[(size_t)q * nz + blockIdx.z] = w;
    }
    if (lane == 0) {
      *tick_layer = 0u;
      __threadfence();
      last = (atomicAdd(R.ticket, 1u) == nz - 1) ? 2 : 0;
    }
  }
  __syncthreads();
  if (last != 2) return false;
  // ---- fold the layers ----
  __threadfence();
#pragma unroll
  for (int q = 0; q < NV; q++) {
    double w = red_identity<OP>();
    for (unsigned int b = tid; b < nz; b += nthreads) w = red_op<OP>(w, __ldcg(top + (size_t)q * nz + b));
    w = warp_reduce<OP>(w);
    __syncthreads();
    if (lane == 0) sm[q][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; q++) {
      double w = lane < nwarps ? sm[q][lane] : red_identity<OP>();
      w = warp_reduce<OP>(w);
      if (lane == 0) {
        R.out[slot0 + q] = w;
        fin[q] = w;
        if (R.tag) R.hout[slot0 + q] = w;
      }
    }
    if (lane == 0) {
      *R.ticket = 0u;
      if (R.tag) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(R.hseq + slot0) = R.tag;
      }
    }
  }
  return true;
}
