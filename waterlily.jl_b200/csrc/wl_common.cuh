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

// ---- deterministic single-pass grid reduction ------------------------------------------
// Every block reduces its values in double with a fixed shuffle tree, writes one partial
// per value, and the last block to arrive (atomic ticket) folds all partials in a fixed
// order and stores the results to out[slot0 .. slot0+NV).  Bit-reproducible for a fixed
// launch geometry; no host round trip.
struct RedBuf {
  double* partials;      // capacity >= NV * (number of blocks + gridDim.z)
  unsigned int* ticket;  // [0] = top ticket, [1+z] = per-layer tickets; all zero between launches
  double* out;           // result slots
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
      }
    }
    if (lane == 0) *R.ticket = 0u;
  }
  return true;
}
