// wl_fast.cuh — the bandwidth-tuned 3-D kernels ("march" kernels) of the pressure path.
//
// Geometry: blockDim = (32, FTY).  A lane owns 4 consecutive interior x cells (one aligned 16-byte
// vector: the layout puts interior cell i=1 on a 128-byte line), a warp owns a 128-cell row segment,
// a block owns FTY rows and marches over a chunk of z planes keeping the z-1 / z / z+1 values of the
// stencil field in registers.  Per plane a thread issues one vector load for the new plane and two
// for the y neighbours (L1/L2 hits: sibling rows of the same block loaded them one step earlier);
// x neighbours come from warp shuffles, with one scalar load at each end of the row segment.
// Every DRAM byte of a field is therefore fetched about once per kernel.
//
// UNI = uniform-coefficient specialisation: no body and all directions periodic, so on every level
// L ≡ Lc[d] on all faces, D ≡ -2ΣLc, iD ≡ 1/D and the coefficient arrays are never read
// (SURVEY.md §8d "constant-coefficient specialisation"; legality is checked on the device by
// k_check_uniform when the hierarchy is (re)built).  !UNI reads L, D, iD like the reference.
//
// All arithmetic is written in the reference's association order; results are bit-identical to the
// general kernels in wl_kernels.cuh and to the oracle.  Requires (N0-2) % 4 == 0.
#pragma once
#include "wl_kernels.cuh"

#define FTY 8
#define MARCH_MINB 6  // ≤40 registers ⇒ 6 blocks (48 warps) per SM: measured +10% on every march kernel over the compiler's default
#ifndef MARCH_MINB_GEN
#define MARCH_MINB_GEN 4  // general-coefficient variants: 64 registers, no spills (sphere 512×256×256: 6 → 13.0, 5 → 12.0, 4 → 11.6 ms/step)
#endif
#ifndef JACOBI_MINB
#define JACOBI_MINB MARCH_MINB
#endif
#ifndef DIVRES_MINB
#define DIVRES_MINB 4  // f_div_residual carries two double accumulators and three velocity stencils: 40 registers spill
#endif
#define FULLMASK 0xffffffffu

struct Coef {
  const float* L;
  const float* Dg;
  const float* iD;
  // general mode: per march block, 1 if every face coefficient the block reads is the level's fluid value Lc (0 on wall faces, the
  // pattern BC!(L,0) leaves) — no body nearby.  Such blocks take L from registers and read only D and iD ("semi-uniform"): 12 B
  // per cell less in every pressure kernel, same bits.  nullptr: disabled.
  const unsigned char* semi;
  float Lc[3];  // UNI: face coefficient per direction
  float Dc, iDc;
};

// 4-byte asynchronous global→shared copy (LDGSTS) and its completion fence
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float get4(const float4& a, int q) { return q == 0 ? a.x : (q == 1 ? a.y : (q == 2 ? a.z : a.w)); }

// Per-thread marching frame
struct Frame {
  int lane, x0, y, z0, z1;  // cells x0..x0+3, row y, planes z0..z1-1
  bool on;                  // this lane owns interior cells
  bool lastgrp;             // the group ends at the last interior cell (its right neighbour is not in the next lane)
  int xl, xr;               // x index of the left neighbour of cell x0 and of the right neighbour of cell x0+3 (periodic wrap applied)
  int ym, yp;               // neighbouring rows (periodic wrap applied)
  i64 row;                  // offset of (i=0, y, k=0) incl. xo
  i64 rowm, rowp;           // same for rows ym, yp
  bool semi;                // this block is semi-uniform (Coef::semi)
  bool wx0, wx1, wy0, wy1;  // the group's first / last cell, the row: next to a wall in -x, +x, -y, +y (face coefficient 0)
};

__device__ __forceinline__ Frame make_frame(const Grid& g, int zchunk, const int3 vb) {
  Frame f;
  f.lane = threadIdx.x;
  f.x0 = 1 + 4 * (32 * vb.x + threadIdx.x);
  f.y = 1 + FTY * vb.y + threadIdx.y;
  f.z0 = 1 + zchunk * vb.z;
  f.z1 = min(f.z0 + zchunk, g.N[2] - 1);
  f.on = f.x0 <= g.N[0] - 2 && f.y <= g.N[1] - 2;
  f.lastgrp = f.x0 + 4 > g.N[0] - 2;
  f.xl = (g.per[0] && f.x0 == 1) ? g.N[0] - 2 : f.x0 - 1;
  f.xr = (g.per[0] && f.x0 + 3 == g.N[0] - 2) ? 1 : f.x0 + 4;
  const int yy = min(f.y, g.N[1] - 2);
  f.ym = (g.per[1] && yy == 1) ? g.N[1] - 2 : yy - 1;
  f.yp = (g.per[1] && yy == g.N[1] - 2) ? 1 : yy + 1;
  f.row = (i64)g.xo + g.s[1] * yy;
  f.rowm = (i64)g.xo + g.s[1] * f.ym;
  f.rowp = (i64)g.xo + g.s[1] * f.yp;
  f.semi = false;
  f.wx0 = !g.per[0] && f.x0 == 1;
  f.wx1 = !g.per[0] && f.x0 + 3 == g.N[0] - 2;
  f.wy0 = !g.per[1] && yy == 1;
  f.wy1 = !g.per[1] && yy == g.N[1] - 2;
  return f;
}
// the same for a kernel that reads coefficients: picks up the block's semi-uniform flag
__device__ __forceinline__ Frame make_frame(const Grid& g, int zchunk, const int3 vb, const Coef& c) {
  Frame f = make_frame(g, zchunk, vb);
  if (c.semi) f.semi = c.semi[vb.x + cdiv(g.N[0] - 2, 128) * (vb.y + cdiv(g.N[1] - 2, FTY) * vb.z)] != 0;
  return f;
}
// face coefficients of a semi-uniform block: lower / upper z faces of plane z, and the splats used by the operators
__device__ __forceinline__ bool wall_zlo(const Grid& g, int z) { return !g.per[2] && !g.zopen[0] && z == 1; }
__device__ __forceinline__ bool wall_zhi(const Grid& g, int z) { return !g.per[2] && !g.zopen[1] && z == g.N[2] - 2; }
__device__ __forceinline__ float4 splat4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ Frame make_frame(const Grid& g, int zchunk) { return make_frame(g, zchunk, real_block()); }
__device__ __forceinline__ int zwrap_lo(const Grid& g, int z) { return (g.per[2] && z == 1) ? g.N[2] - 2 : z - 1; }
__device__ __forceinline__ int zwrap_hi(const Grid& g, int z) { return (g.per[2] && z == g.N[2] - 2) ? 1 : z + 1; }

// x neighbours of a (transformed) vector: left of .x and right of .w.  `edge_l`/`edge_r` are the already transformed
// scalars to use at the ends of the warp's row segment.  All lanes must call this (warp shuffles).
__device__ __forceinline__ void x_nbrs(const Frame& f, const float4& c, float edge_l, float edge_r, float& left, float& right) {
  left = __shfl_up_sync(FULLMASK, c.w, 1);
  right = __shfl_down_sync(FULLMASK, c.x, 1);
  if (f.lane == 0) left = edge_l;
  if (f.lane == 31 || f.lastgrp) right = edge_r;
}

// A x at the 4 cells of a lane for the uniform operator: s = x·D + (xl·L0 + xr·L0) + (ym·L1 + yp·L1) + (zm·L2 + zp·L2)
__device__ __forceinline__ float4 mult_uni(const Coef& c, const float4& xc, float left, float right, const float4& ym, const float4& yp, const float4& zm,
                                           const float4& zp) {
  float4 s;
  const float L0 = c.Lc[0], L1 = c.Lc[1], L2 = c.Lc[2], D = c.Dc;
  s.x = xc.x * D;
  s.x += left * L0 + xc.y * L0;
  s.x += ym.x * L1 + yp.x * L1;
  s.x += zm.x * L2 + zp.x * L2;
  s.y = xc.y * D;
  s.y += xc.x * L0 + xc.z * L0;
  s.y += ym.y * L1 + yp.y * L1;
  s.y += zm.y * L2 + zp.y * L2;
  s.z = xc.z * D;
  s.z += xc.y * L0 + xc.w * L0;
  s.z += ym.z * L1 + yp.z * L1;
  s.z += zm.z * L2 + zp.z * L2;
  s.w = xc.w * D;
  s.w += xc.z * L0 + right * L0;
  s.w += ym.w * L1 + yp.w * L1;
  s.w += zm.w * L2 + zp.w * L2;
  return s;
}

// General operator: coefficients from memory.  Llo0 = L[I,0] (4 values), Lhi0r = L[I+δ0,0] of the LAST cell (the others are
// Llo0 shifted), Llo1/Lhi1 = L[I,1], L[I+δ1,1]; Llo2/Lhi2 likewise; Dg = D[I].
__device__ __forceinline__ float4 mult_gen(const float4& xc, float left, float right, const float4& ym, const float4& yp, const float4& zm,
                                           const float4& zp, const float4& Dg, const float4& Llo0, float Lhi0r, const float4& Llo1, const float4& Lhi1,
                                           const float4& Llo2, const float4& Lhi2) {
  float4 s;
  s.x = xc.x * Dg.x;
  s.x += left * Llo0.x + xc.y * Llo0.y;
  s.x += ym.x * Llo1.x + yp.x * Lhi1.x;
  s.x += zm.x * Llo2.x + zp.x * Lhi2.x;
  s.y = xc.y * Dg.y;
  s.y += xc.x * Llo0.y + xc.z * Llo0.z;
  s.y += ym.y * Llo1.y + yp.y * Lhi1.y;
  s.y += zm.y * Llo2.y + zp.y * Lhi2.y;
  s.z = xc.z * Dg.z;
  s.z += xc.y * Llo0.z + xc.w * Llo0.w;
  s.z += ym.z * Llo1.z + yp.z * Lhi1.z;
  s.z += zm.z * Llo2.z + zp.z * Lhi2.z;
  s.w = xc.w * Dg.w;
  s.w += xc.z * Llo0.w + right * Lhi0r;
  s.w += ym.w * Llo1.w + yp.w * Lhi1.w;
  s.w += zm.w * Llo2.w + zp.w * Lhi2.w;
  return s;
}

// Loads the coefficient vectors of the lane's 4 cells on plane offset `pz` (= s2*z) and applies the operator.
template <bool UNI>
__device__ __forceinline__ float4 apply_A(const Grid& g, const Coef& c, const Frame& f, i64 pz, const float4& xc, float left, float right, const float4& ym,
                                          const float4& yp, const float4& zm, const float4& zp) {
  if (UNI) return mult_uni(c, xc, left, right, ym, yp, zm, zp);
  const i64 o = f.row + pz + f.x0;
  if (f.semi) {
    const bool wz0 = !g.per[2] && !g.zopen[0] && pz == g.s[2], wz1 = !g.per[2] && !g.zopen[1] && pz == g.s[2] * (g.N[2] - 2);
    float4 Llo0 = splat4(c.Lc[0]);
    if (f.wx0) Llo0.x = 0.f;
    return mult_gen(xc, left, right, ym, yp, zm, zp, ld4(c.Dg + o), Llo0, f.wx1 ? 0.f : c.Lc[0], splat4(f.wy0 ? 0.f : c.Lc[1]), splat4(f.wy1 ? 0.f : c.Lc[1]),
                    splat4(wz0 ? 0.f : c.Lc[2]), splat4(wz1 ? 0.f : c.Lc[2]));
  }
  const float4 Dg = ld4(c.Dg + o);
  const float4 Llo0 = ld4(c.L + o);
  const float Lhi0r = c.L[o + 4];
  const float4 Llo1 = ld4(c.L + g.sc + o), Lhi1 = ld4(c.L + g.sc + o + g.s[1]);
  const float4 Llo2 = ld4(c.L + 2 * g.sc + o), Lhi2 = ld4(c.L + 2 * g.sc + o + g.s[2]);
  return mult_gen(xc, left, right, ym, yp, zm, zp, Dg, Llo0, Lhi0r, Llo1, Lhi1, Llo2, Lhi2);
}

// ------------------------------------------------------------------------------------------------
// Stencil-field reader with a pointwise transform:  kind 0: q          (ϵ, x)
//                                                   kind 1: q · s      (x = p·dt,  s scalar)
//                                                   kind 2: q · w[o]   (ϵ = r·iD with iD from memory)
// In UNI mode r·iD uses kind 1 with s = iDc.
// ------------------------------------------------------------------------------------------------
struct SField {
  const float* q;
  const float* w;
  float s;
  int kind;
  __device__ __forceinline__ float4 v4(i64 o) const {
    float4 a = ld4(q + o);
    if (kind == 1) a = scale4(a, s);
    if (kind == 2) a = mul4(a, ld4(w + o));
    return a;
  }
  __device__ __forceinline__ float v1(i64 o) const {
    float a = q[o];
    if (kind == 1) a = a * s;
    if (kind == 2) a = a * w[o];
    return a;
  }
};

// The common marching loop.  For every plane z of the chunk it hands the functor the centre vector and its six
// neighbours of the stencil field:  body(z, pz, o, c, left, right, ym, yp, zm, zp).
template <class Body>
__device__ __forceinline__ void march7(const Grid& g, const Frame& f, const SField& F, Body body) {
  float4 zm = f4zero(), c = f4zero(), zp = f4zero();
  if (f.on) {
    zm = F.v4(f.row + g.s[2] * zwrap_lo(g, f.z0) + f.x0);
    c = F.v4(f.row + g.s[2] * f.z0 + f.x0);
  }
  for (int z = f.z0; z < f.z1; z++) {
    const i64 pz = g.s[2] * z;
    float4 ym = f4zero(), yp = f4zero();
    float el = 0.f, er = 0.f;
    if (f.on) {
      zp = F.v4(f.row + g.s[2] * zwrap_hi(g, z) + f.x0);
      ym = F.v4(f.rowm + pz + f.x0);
      yp = F.v4(f.rowp + pz + f.x0);
      if (f.lane == 0) el = F.v1(f.row + pz + f.xl);
      if (f.lane == 31 || f.lastgrp) er = F.v1(f.row + pz + f.xr);
    }
    float left, right;
    x_nbrs(f, c, el, er, left, right);
    body(z, pz, f.row + pz + f.x0, c, left, right, ym, yp, zm, zp);
    zm = c;
    c = zp;
  }
}

// ------------------------------------------------------------------------------------------------
// Jacobi!(p) with ω=1 fused with increment! and restrict! (src/Poisson.jl:111-114,100-104; src/MultiLevelPoisson.jl:49,92-94):
//   ϵ = r·iD;  r' = r − Aϵ → r2;  x (+)= ϵ;   coarse.r[I] = Σ r' over up(I)  (x fastest, then y, then z)
// The y pair of a coarse cell sits in two adjacent warps: r' is exchanged through shared memory once per plane.
// ------------------------------------------------------------------------------------------------
template <bool UNI>
__device__ __forceinline__ void b_f_jacobi(const Grid& g, const Coef& c, const float* r, float* r2, float* x,
                                                     int x_is_zero, int zchunk, Grid gc, float* rc, int do_restrict, int zoffc, const int3 vb) {
  __shared__ float4 ex[FTY][32];
  const Frame f = make_frame(g, zchunk, vb, c);
  SField F;
  F.q = r;
  F.w = c.iD;
  F.s = c.iDc;
  F.kind = UNI ? 1 : 2;
  float2 acc = make_float2(0.f, 0.f);  // running sums of the two coarse cells this lane contributes to
  march7(g, f, F, [&](int z, i64 pz, i64 o, const float4& e, float left, float right, const float4& ym, const float4& yp, const float4& zm,
                      const float4& zp) {
    float4 rn = f4zero();
    if (f.on) {
      const float4 Ae = apply_A<UNI>(g, c, f, pz, e, left, right, ym, yp, zm, zp);
      const float4 ro = ld4(r + o);
      rn = make_float4(ro.x - 1.f * Ae.x, ro.y - 1.f * Ae.y, ro.z - 1.f * Ae.z, ro.w - 1.f * Ae.w);
      st4(r2 + o, rn);
      if (x_is_zero)
        st4(x + o, e);
      else {
        const float4 xo = ld4(x + o);
        st4(x + o, make_float4(xo.x + 1.f * e.x, xo.y + 1.f * e.y, xo.z + 1.f * e.z, xo.w + 1.f * e.w));
      }
    }
    if (do_restrict) {
      // restrict(I,b,c) sums b[J] for J ∈ up(I): x fastest, then y, then z  (src/MultiLevelPoisson.jl:13-19)
      ex[threadIdx.y][f.lane] = rn;
      __syncthreads();
      const bool ylow = (threadIdx.y & 1) == 0;  // tile origin y=1 is the first row of a pair
      if (ylow) {
        const float4 up = ex[threadIdx.y + 1][f.lane];
        const bool zlow = ((z - 1) & 1) == 0;
        if (zlow) {
          acc.x = 0.f;
          acc.y = 0.f;
        }
        acc.x += rn.x;
        acc.x += rn.y;
        acc.x += up.x;
        acc.x += up.y;
        acc.y += rn.z;
        acc.y += rn.w;
        acc.y += up.z;
        acc.y += up.w;
        if (!zlow && f.on) {
          // coarse cell indices: (x0+1)/2, (x0+3)/2 ; (y+1)/2 ; (z+1)/2
          const i64 oc = (i64)gc.xo + (f.x0 + 1) / 2 + gc.s[1] * ((f.y + 1) / 2) + gc.s[2] * ((z + 1) / 2 + zoffc);
          rc[oc] = acc.x;
          rc[oc + 1] = acc.y;
        }
      }
      __syncthreads();
    }
  });
}
template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, (UNI ? JACOBI_MINB : MARCH_MINB_GEN)) f_jacobi(Grid g, Coef c, const float* __restrict__ r, float* __restrict__ r2, float* __restrict__ x,
                                                     int x_is_zero, int zchunk, Grid gc, float* __restrict__ rc, int do_restrict, int zoffc) {
  pdl_wait();
  b_f_jacobi<UNI>(g, c, r, r2, x, x_is_zero, zchunk, gc, rc, do_restrict, zoffc, real_block());
}

// f_jacobi<true> (uniform mode, restriction fused) with every load of a plane issued before the first use and the planes
// of r carried in registers — the structure of f_divres_uni / f_correct_cfl — and TWO rows per thread (blockDim = (32, FTY/2)): the y pair
// of a coarse cell sits in one thread, so the restriction needs neither shared memory nor block barriers, and the rows serve as each
// other's y neighbours.  Same operations in the same order as b_f_jacobi.  0.37 ms at 512³ (94 % of the copy peak) against 0.47 ms.
__global__ void __launch_bounds__(32 * FTY / 2, 6) f_jacobi_uni2(const __grid_constant__ Grid g, const __grid_constant__ Coef c, const float* __restrict__ r,
                                                                float* __restrict__ r2, float* __restrict__ x, int zchunk, const __grid_constant__ Grid gc,
                                                                float* __restrict__ rc, int zoffc, int x_is_zero) {
  pdl_wait();
  const int lane = threadIdx.x;
  const int x0 = 1 + 4 * (32 * blockIdx.x + lane);
  const int ya = 1 + FTY * blockIdx.y + 2 * threadIdx.y;  // rows ya, ya+1 (the interior height is even)
  const int z0 = 1 + zchunk * blockIdx.z, z1 = min(z0 + zchunk, g.N[2] - 1);
  const bool on = x0 <= g.N[0] - 2 && ya <= g.N[1] - 2;
  const bool lastgrp = x0 + 4 > g.N[0] - 2;
  const int xl = (g.per[0] && x0 == 1) ? g.N[0] - 2 : x0 - 1;
  const int xr = (g.per[0] && x0 + 3 == g.N[0] - 2) ? 1 : x0 + 4;
  const int yy = min(ya, g.N[1] - 3);
  const int ym_ = (g.per[1] && yy == 1) ? g.N[1] - 2 : yy - 1;
  const int yp_ = (g.per[1] && yy + 1 == g.N[1] - 2) ? 1 : yy + 2;
  const i64 rowa = (i64)g.xo + g.s[1] * yy, rowb = rowa + g.s[1], rowm = (i64)g.xo + g.s[1] * ym_, rowp = (i64)g.xo + g.s[1] * yp_;
  const float iD = c.iDc;
  float4 zma = f4zero(), zmb = f4zero(), ra = f4zero(), rb = f4zero();
  if (on) {
    const i64 pm = g.s[2] * zwrap_lo(g, z0), p0 = g.s[2] * z0;
    zma = scale4(ld4(r + rowa + pm + x0), iD);
    zmb = scale4(ld4(r + rowb + pm + x0), iD);
    ra = ld4(r + rowa + p0 + x0);
    rb = ld4(r + rowb + p0 + x0);
  }
  float2 acc = make_float2(0.f, 0.f);
  for (int z = z0; z < z1; z++) {
    const i64 pz = g.s[2] * z, pp = g.s[2] * zwrap_hi(g, z);
    float4 rpa = f4zero(), rpb = f4zero(), ym = f4zero(), yp = f4zero(), xa = f4zero(), xb = f4zero();
    float ela = 0.f, era = 0.f, elb = 0.f, erb = 0.f;
    if (on) {
      rpa = ld4(r + rowa + pp + x0);
      rpb = ld4(r + rowb + pp + x0);
      ym = ld4(r + rowm + pz + x0);
      yp = ld4(r + rowp + pz + x0);
      if (!x_is_zero) {
        xa = ld4(x + rowa + pz + x0);
        xb = ld4(x + rowb + pz + x0);
      }
      if (lane == 0) {
        ela = r[rowa + pz + xl];
        elb = r[rowb + pz + xl];
      }
      if (lane == 31 || lastgrp) {
        era = r[rowa + pz + xr];
        erb = r[rowb + pz + xr];
      }
    }
    const float4 ea = scale4(ra, iD), eb = scale4(rb, iD), zpa = scale4(rpa, iD), zpb = scale4(rpb, iD);
    ym = scale4(ym, iD);
    yp = scale4(yp, iD);
    float lfa = __shfl_up_sync(FULLMASK, ea.w, 1), rta = __shfl_down_sync(FULLMASK, ea.x, 1);
    float lfb = __shfl_up_sync(FULLMASK, eb.w, 1), rtb = __shfl_down_sync(FULLMASK, eb.x, 1);
    if (lane == 0) {
      lfa = ela * iD;
      lfb = elb * iD;
    }
    if (lane == 31 || lastgrp) {
      rta = era * iD;
      rtb = erb * iD;
    }
    if (on) {
      const float4 Aa = mult_uni(c, ea, lfa, rta, ym, eb, zma, zpa);
      const float4 Ab = mult_uni(c, eb, lfb, rtb, ea, yp, zmb, zpb);
      const float4 na = make_float4(ra.x - 1.f * Aa.x, ra.y - 1.f * Aa.y, ra.z - 1.f * Aa.z, ra.w - 1.f * Aa.w);
      const float4 nb = make_float4(rb.x - 1.f * Ab.x, rb.y - 1.f * Ab.y, rb.z - 1.f * Ab.z, rb.w - 1.f * Ab.w);
      st4(r2 + rowa + pz + x0, na);
      st4(r2 + rowb + pz + x0, nb);
      if (x_is_zero) {  // fill!(x,0) before it: x = ϵ
        st4(x + rowa + pz + x0, ea);
        st4(x + rowb + pz + x0, eb);
      } else {
        st4(x + rowa + pz + x0, make_float4(xa.x + 1.f * ea.x, xa.y + 1.f * ea.y, xa.z + 1.f * ea.z, xa.w + 1.f * ea.w));
        st4(x + rowb + pz + x0, make_float4(xb.x + 1.f * eb.x, xb.y + 1.f * eb.y, xb.z + 1.f * eb.z, xb.w + 1.f * eb.w));
      }
      // restrict!(coarse.r, fine.r): x fastest, then y, then z  (src/MultiLevelPoisson.jl:13-19)
      const bool zlow = ((z - 1) & 1) == 0;
      if (zlow) {
        acc.x = 0.f;
        acc.y = 0.f;
      }
      acc.x += na.x;
      acc.x += na.y;
      acc.x += nb.x;
      acc.x += nb.y;
      acc.y += na.z;
      acc.y += na.w;
      acc.y += nb.z;
      acc.y += nb.w;
      if (!zlow) {
        const i64 oc = (i64)gc.xo + (x0 + 1) / 2 + gc.s[1] * ((ya + 1) / 2) + gc.s[2] * ((z + 1) / 2 + zoffc);
        rc[oc] = acc.x;
        rc[oc + 1] = acc.y;
      }
    }
    zma = ea;
    zmb = eb;
    ra = rpa;
    rb = rpb;
  }
}

// ------------------------------------------------------------------------------------------------
// increment!(p;ω) (src/Poisson.jl:100-104) with the ϵ source either the level's own ϵ array (after GaussSeidelRB!) or
// the prolongation of the coarse solution, ϵ[I] = coarse.x[down(I)] (src/MultiLevelPoisson.jl:50,99-100), never stored:
//   r −= ω·Aϵ ;  x += ω·ϵ ;  optionally Σr² → out[slot]  (L₂, src/Poisson.jl:189)
// ------------------------------------------------------------------------------------------------
struct ProlongSrc {  // reads ϵ = xc[down(·)] for a fine row segment
  const float* xc;
  Grid gc;
  // z-slab fine level over a replicated (whole-domain) coarse level: global index of fine plane 0, global fine N2 and whether z is
  // globally periodic (the replicated level wraps indices itself); all zero when fine and coarse share the decomposition
  int zoff, N2g, perg;
};

template <bool UNI, bool PROLONG>
__device__ __forceinline__ void b_f_increment(const Grid& g, const Coef& c, const float* eps, ProlongSrc ps, float* r,
                                                        float* x, const float* wp, int x_is_zero, int zchunk, int with_l2,
                                                        RedBuf R, int slot, const int3 vb) {
  const Frame f = make_frame(g, zchunk, vb, c);
  const float w = *wp;
  double l2 = 0.0;
  if (!PROLONG) {
    SField F;
    F.q = eps;
    F.w = nullptr;
    F.s = 1.f;
    F.kind = 0;
    march7(g, f, F, [&](int z, i64 pz, i64 o, const float4& e, float left, float right, const float4& ym, const float4& yp, const float4& zm,
                        const float4& zp) {
      if (f.on) {
        const float4 Ae = apply_A<UNI>(g, c, f, pz, e, left, right, ym, yp, zm, zp);
        const float4 ro = ld4(r + o);
        const float4 rn = make_float4(ro.x - w * Ae.x, ro.y - w * Ae.y, ro.z - w * Ae.z, ro.w - w * Ae.w);
        st4(r + o, rn);
        if (x_is_zero)
          st4(x + o, make_float4(w * e.x, w * e.y, w * e.z, w * e.w));
        else {
          const float4 xo = ld4(x + o);
          st4(x + o, make_float4(xo.x + w * e.x, xo.y + w * e.y, xo.z + w * e.z, xo.w + w * e.w));
        }
        l2 += (double)rn.x * rn.x + (double)rn.y * rn.y + (double)rn.z * rn.z + (double)rn.w * rn.w;
      }
    });
  } else {
    // Prolongation source: a fine group x0..x0+3 (x0 odd) maps to coarse cells cx, cx+1 with cx=(x0+1)/2:
    //   fine x0-1 → cx-1 ; x0,x0+1 → cx ; x0+2,x0+3 → cx+1 ; x0+4 → cx+2   (with the fine periodic wrap applied first)
    const Grid& gc = ps.gc;
    auto crow = [&](int yf, int zf) -> i64 {
      int zg = zf + ps.zoff;
      if (ps.perg) {
        if (zg == 0) zg = ps.N2g - 2;
        else if (zg == ps.N2g - 1) zg = 1;
      }
      return (i64)gc.xo + gc.s[1] * ((yf + 1) / 2) + gc.s[2] * ((zg + 1) / 2);
    };
    auto ld2 = [&](i64 rowc) -> float4 {  // ϵ at the lane's 4 fine cells on a coarse row
      const int cx = (f.x0 + 1) / 2;
      const float a = ps.xc[rowc + cx], b = ps.xc[rowc + cx + 1];
      return make_float4(a, a, b, b);
    };
    for (int z = f.z0; z < f.z1; z++) {
      if (!f.on) continue;
      const i64 pz = g.s[2] * z;
      const i64 o = f.row + pz + f.x0;
      const i64 rc0 = crow(f.y, z);
      const float4 e = ld2(rc0);
      const float left = ps.xc[rc0 + (f.xl + 1) / 2];
      const float right = ps.xc[rc0 + (f.xr + 1) / 2];
      const float4 ym = ld2(crow(f.ym, z)), yp = ld2(crow(f.yp, z));
      const float4 zm = ld2(crow(f.y, zwrap_lo(g, z))), zp = ld2(crow(f.y, zwrap_hi(g, z)));
      const float4 Ae = apply_A<UNI>(g, c, f, pz, e, left, right, ym, yp, zm, zp);
      const float4 ro = ld4(r + o);
      const float4 rn = make_float4(ro.x - w * Ae.x, ro.y - w * Ae.y, ro.z - w * Ae.z, ro.w - w * Ae.w);
      st4(r + o, rn);
      const float4 xo = ld4(x + o);
      st4(x + o, make_float4(xo.x + w * e.x, xo.y + w * e.y, xo.z + w * e.z, xo.w + w * e.w));
    }
  }
  if (with_l2) {
    double v[1] = {l2}, fin[1];
    grid_reduce<RED_SUM, 1>(v, R, slot, fin);
  }
}
template <bool UNI, bool PROLONG>
__global__ void __launch_bounds__(32 * FTY, (UNI ? MARCH_MINB : MARCH_MINB_GEN)) f_increment(Grid g, Coef c, const float* __restrict__ eps, ProlongSrc ps, float* __restrict__ r,
                                                        float* __restrict__ x, const float* __restrict__ wp, int x_is_zero, int zchunk, int with_l2,
                                                        RedBuf R, int slot) {
  pdl_wait();
  b_f_increment<UNI, PROLONG>(g, c, eps, ps, r, x, wp, x_is_zero, zchunk, with_l2, R, slot, real_block());
}

// ------------------------------------------------------------------------------------------------
// div + x.*=dt + residual! part 1 (src/Flow.jl:225, src/Poisson.jl:93-95):
//   z = Σ_d (u_d[I+δ_d] − u_d[I]);  x = p·dt;  r = iD==0 ? 0 : z − A x;   Σr → out[slot], Σr² → out[slot+1]  (slot+1 must be SLOT_R2)
// σ (=z) is not stored in UNI mode (it is pure scratch there; CFL rewrites the interior every step).
// ------------------------------------------------------------------------------------------------
template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, DIVRES_MINB) f_div_residual(Grid g, Coef c, const float* __restrict__ u, const float* __restrict__ p, float* __restrict__ x,
                                                           float* __restrict__ r, float* __restrict__ zarr, const float* __restrict__ dtp, float wdt,
                                                           int zchunk, RedBuf R, int slot) {
  pdl_wait();
  const Frame f = make_frame(g, zchunk, real_block(), c);
  const float dt = wdt * (*dtp);
  SField F;
  F.q = p;
  F.w = nullptr;
  F.s = dt;
  F.kind = 1;
  double sum = 0.0, l2 = 0.0;  // Σr and, for the case that residual! leaves r alone (|mean| ≤ 2eps), Σr² = L₂
  march7(g, f, F, [&](int z, i64 pz, i64 o, const float4& xs, float left, float right, const float4& ym, const float4& yp, const float4& zm,
                      const float4& zp) {
    // u_x needs the value one cell to the right of the group
    float4 ux = f4zero();
    float uxr_edge = 0.f;
    if (f.on) {
      ux = ld4(u + o);
      if (f.lane == 31 || f.lastgrp) uxr_edge = u[f.row + pz + f.xr];
    }
    float uxr = __shfl_down_sync(FULLMASK, ux.x, 1);
    if (f.lane == 31 || f.lastgrp) uxr = uxr_edge;
    if (f.on) {
      // upper neighbours through the periodic wrap (= the ghost value after BC!), so u's periodic ghosts need not be current
      const float4 uy = ld4(u + g.sc + o), uyp = ld4(u + g.sc + f.rowp + pz + f.x0);
      const float4 uz = ld4(u + 2 * g.sc + o), uzp = ld4(u + 2 * g.sc + f.row + g.s[2] * zwrap_hi(g, z) + f.x0);
      float4 dv;
      dv.x = 0.f + (ux.y - ux.x);
      dv.x += uyp.x - uy.x;
      dv.x += uzp.x - uz.x;
      dv.y = 0.f + (ux.z - ux.y);
      dv.y += uyp.y - uy.y;
      dv.y += uzp.y - uz.y;
      dv.z = 0.f + (ux.w - ux.z);
      dv.z += uyp.z - uy.z;
      dv.z += uzp.z - uz.z;
      dv.w = 0.f + (uxr - ux.w);
      dv.w += uyp.w - uy.w;
      dv.w += uzp.w - uz.w;
      const float4 Ax = apply_A<UNI>(g, c, f, pz, xs, left, right, ym, yp, zm, zp);
      float4 rr = make_float4(dv.x - Ax.x, dv.y - Ax.y, dv.z - Ax.z, dv.w - Ax.w);
      if (!UNI) {
        const float4 iD = ld4(c.iD + o);
        if (iD.x == 0.f) rr.x = 0.f;
        if (iD.y == 0.f) rr.y = 0.f;
        if (iD.z == 0.f) rr.z = 0.f;
        if (iD.w == 0.f) rr.w = 0.f;
        st4(zarr + o, dv);
      }
      st4(x + o, xs);
      st4(r + o, rr);
      sum += (double)rr.x + (double)rr.y + (double)rr.z + (double)rr.w;
      l2 += (double)rr.x * rr.x + (double)rr.y * rr.y + (double)rr.z * rr.z + (double)rr.w * rr.w;
    }
  });
  double v[2] = {sum, l2}, fin[2];
  grid_reduce<RED_SUM, 2>(v, R, slot, fin);
}

// f_div_residual<true> with every load of a plane issued before the first use and the planes of p and u_z carried in registers
// (the structure that took f_correct_cfl to the copy peak): same operations in the same order.
__global__ void __launch_bounds__(32 * FTY, 4) f_divres_uni(const __grid_constant__ Grid g, const __grid_constant__ Coef c, const float* __restrict__ u,
                                                           const float* __restrict__ p, float* __restrict__ x, float* __restrict__ r,
                                                           const float* __restrict__ dtp, float wdt, int zchunk, RedBuf R, int slot) {
  pdl_wait();
  const Frame f = make_frame(g, zchunk);
  const float dt = wdt * (*dtp);
  double sum = 0.0, l2 = 0.0;
  float4 zm = f4zero(), xc = f4zero(), uz = f4zero();
  if (f.on) {
    zm = scale4(ld4(p + f.row + g.s[2] * zwrap_lo(g, f.z0) + f.x0), dt);
    xc = scale4(ld4(p + f.row + g.s[2] * f.z0 + f.x0), dt);
    uz = ld4(u + 2 * g.sc + f.row + g.s[2] * f.z0 + f.x0);
  }
  for (int z = f.z0; z < f.z1; z++) {
    const i64 pz = g.s[2] * z;
    const i64 o = f.row + pz + f.x0;
    const i64 ozp = f.row + g.s[2] * zwrap_hi(g, z) + f.x0;
    float4 zp = f4zero(), ym = f4zero(), yp = f4zero(), ux = f4zero(), uy = f4zero(), uyp = f4zero(), uzp = f4zero();
    float el = 0.f, er = 0.f, uxr = 0.f;
    if (f.on) {
      zp = ld4(p + ozp);
      ym = ld4(p + f.rowm + pz + f.x0);
      yp = ld4(p + f.rowp + pz + f.x0);
      ux = ld4(u + o);
      uy = ld4(u + g.sc + o);
      uyp = ld4(u + g.sc + f.rowp + pz + f.x0);
      uzp = ld4(u + 2 * g.sc + ozp);
      if (f.lane == 0) el = p[f.row + pz + f.xl];
      if (f.lane == 31 || f.lastgrp) {
        er = p[f.row + pz + f.xr];
        uxr = u[f.row + pz + f.xr];
      }
    }
    zp = scale4(zp, dt);
    ym = scale4(ym, dt);
    yp = scale4(yp, dt);
    el = el * dt;
    er = er * dt;
    float left, right;
    x_nbrs(f, xc, el, er, left, right);
    float unext = __shfl_down_sync(FULLMASK, ux.x, 1);
    if (f.lane == 31 || f.lastgrp) unext = uxr;
    if (f.on) {
      float4 dv;
      dv.x = 0.f + (ux.y - ux.x);
      dv.x += uyp.x - uy.x;
      dv.x += uzp.x - uz.x;
      dv.y = 0.f + (ux.z - ux.y);
      dv.y += uyp.y - uy.y;
      dv.y += uzp.y - uz.y;
      dv.z = 0.f + (ux.w - ux.z);
      dv.z += uyp.z - uy.z;
      dv.z += uzp.z - uz.z;
      dv.w = 0.f + (unext - ux.w);
      dv.w += uyp.w - uy.w;
      dv.w += uzp.w - uz.w;
      const float4 Ax = mult_uni(c, xc, left, right, ym, yp, zm, zp);
      const float4 rr = make_float4(dv.x - Ax.x, dv.y - Ax.y, dv.z - Ax.z, dv.w - Ax.w);
      st4(x + o, xc);
      st4(r + o, rr);
      sum += (double)rr.x + (double)rr.y + (double)rr.z + (double)rr.w;
      l2 += (double)rr.x * rr.x + (double)rr.y * rr.y + (double)rr.z * rr.z + (double)rr.w * rr.w;
    }
    zm = xc;
    xc = zp;
    uz = uzp;
  }
  double v[2] = {sum, l2}, fin[2];
  grid_reduce<RED_SUM, 2>(v, R, slot, fin);
}

// residual! part 2 + L₂ (src/Poisson.jl:95-97,189): s = Σr/|inside|; |s|>2eps ⇒ r −= s; Σr² → out[slot_out]
__global__ void __launch_bounds__(32 * FTY, MARCH_MINB) f_resid_fix(Grid g, float* __restrict__ r, float count, int zchunk, RedBuf R, int slot_in, int slot_out) {
  pdl_wait();
  const Frame f = make_frame(g, zchunk);
  const float s = (float)R.out[slot_in] / count;
  const bool fix = fabsf(s) > 2.f * 1.1920929e-7f;
  double l2 = 0.0;
  if (f.on) {
    for (int z = f.z0; z < f.z1; z++) {
      const i64 o = f.row + g.s[2] * z + f.x0;
      float4 v = ld4(r + o);
      if (fix) {
        v = make_float4(v.x - s, v.y - s, v.z - s, v.w - s);
        st4(r + o, v);
      }
      l2 += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  }
  double v[1] = {l2}, fin[1];
  grid_reduce<RED_SUM, 1>(v, R, slot_out, fin);
}

// ------------------------------------------------------------------------------------------------
// Velocity correction + pressure unscale of mom_project! (src/Flow.jl:227-230), optionally fused with CFL's flux_out
// on the corrected field is NOT done here: flux_out needs BC-filled ghosts, it stays in k_cfl.
//   u_d[I] −= L[I,d]·(x[I] − x[I−δ_d]);  p = x/dt
// ------------------------------------------------------------------------------------------------
template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, (UNI ? MARCH_MINB : MARCH_MINB_GEN)) f_correct(Grid g, Coef c, const float* __restrict__ x, float* __restrict__ u, float* __restrict__ p,
                                                      const float* __restrict__ dtp, float wdt, int zchunk) {
  pdl_wait();
  const Frame f = make_frame(g, zchunk, real_block(), c);
  const float dt = wdt * (*dtp);
  float4 zm = f4zero();
  if (f.on) zm = ld4(x + f.row + g.s[2] * zwrap_lo(g, f.z0) + f.x0);
  for (int z = f.z0; z < f.z1; z++) {
    const i64 pz = g.s[2] * z;
    const i64 o = f.row + pz + f.x0;
    float4 xc = f4zero(), ym = f4zero();
    float el = 0.f;
    if (f.on) {
      xc = ld4(x + o);
      ym = ld4(x + f.rowm + pz + f.x0);
      if (f.lane == 0) el = x[f.row + pz + f.xl];
    }
    float left = __shfl_up_sync(FULLMASK, xc.w, 1);
    if (f.lane == 0) left = el;
    if (f.on) {
      float4 L0, L1, L2;
      if (UNI) {
        L0 = make_float4(c.Lc[0], c.Lc[0], c.Lc[0], c.Lc[0]);
        L1 = make_float4(c.Lc[1], c.Lc[1], c.Lc[1], c.Lc[1]);
        L2 = make_float4(c.Lc[2], c.Lc[2], c.Lc[2], c.Lc[2]);
      } else if (f.semi) {
        L0 = splat4(c.Lc[0]);
        if (f.wx0) L0.x = 0.f;
        L1 = splat4(f.wy0 ? 0.f : c.Lc[1]);
        L2 = splat4(wall_zlo(g, z) ? 0.f : c.Lc[2]);
      } else {
        L0 = ld4(c.L + o);
        L1 = ld4(c.L + g.sc + o);
        L2 = ld4(c.L + 2 * g.sc + o);
      }
      float4 a = ld4(u + o);
      a.x -= L0.x * (xc.x - left);
      a.y -= L0.y * (xc.y - xc.x);
      a.z -= L0.z * (xc.z - xc.y);
      a.w -= L0.w * (xc.w - xc.z);
      st4(u + o, a);
      float4 b = ld4(u + g.sc + o);
      b.x -= L1.x * (xc.x - ym.x);
      b.y -= L1.y * (xc.y - ym.y);
      b.z -= L1.z * (xc.z - ym.z);
      b.w -= L1.w * (xc.w - ym.w);
      st4(u + g.sc + o, b);
      float4 d = ld4(u + 2 * g.sc + o);
      d.x -= L2.x * (xc.x - zm.x);
      d.y -= L2.y * (xc.y - zm.y);
      d.z -= L2.z * (xc.z - zm.z);
      d.w -= L2.w * (xc.w - zm.w);
      st4(u + 2 * g.sc + o, d);
      st4(p + o, make_float4(xc.x / dt, xc.y / dt, xc.z / dt, xc.w / dt));
    }
    zm = xc;
  }
}

// ------------------------------------------------------------------------------------------------
// f_correct_cfl (uniform mode): the corrector's velocity correction + pressure unscale (f_correct) fused with CFL's flux_out on the
// corrected field (f_cfl; src/Flow.jl:234-244): flux_out(I) needs the corrected u_d at I+δ_d, which this thread forms itself from the
// raw neighbours (u is corrected OUT OF PLACE, u_in → u_out, so raw values stay readable while other blocks work).  Same
// operations in the same order as the two kernels ⇒ same bits; one pass over u instead of two.
// ------------------------------------------------------------------------------------------------
template <bool CFLQ>
__global__ void __launch_bounds__(32 * FTY, 4) f_correct_cfl(const __grid_constant__ Grid g, const __grid_constant__ Coef c, const float* __restrict__ x,
                                                            const float* __restrict__ ui, float* __restrict__ uo, float* __restrict__ p,
                                                            const float* __restrict__ dtp, float wdt, int zchunk, float nu, float* __restrict__ dt_out,
                                                            RedBuf R, int slot, int slot_ghost, int finalize, int* __restrict__ flags) {
  pdl_wait();
  const Frame f = make_frame(g, zchunk);
  const float dt = wdt * (*dtp);
  const float L0 = c.Lc[0], L1 = c.Lc[1], L2 = c.Lc[2];
  double m = 0.0;
  RangeAcc ra;  // range of the velocities this thread writes (the next reader is the flux kernel, see range_note)
  float4 zm = f4zero(), xc = f4zero(), uz = f4zero();
  if (f.on) {
    zm = ld4(x + f.row + g.s[2] * zwrap_lo(g, f.z0) + f.x0);
    xc = ld4(x + f.row + g.s[2] * f.z0 + f.x0);
    uz = ld4(ui + 2 * g.sc + f.row + g.s[2] * f.z0 + f.x0);
  }
  for (int z = f.z0; z < f.z1; z++) {
    const i64 pz = g.s[2] * z;
    const i64 o = f.row + pz + f.x0;
    const i64 ozp = f.row + g.s[2] * zwrap_hi(g, z) + f.x0;
    float4 ym = f4zero(), yp = f4zero(), xzp = f4zero(), ux = f4zero(), uy = f4zero(), uyp = f4zero(), uzp = f4zero();
    float el = 0.f, xr = 0.f, uxr = 0.f;
    if (f.on) {
      ym = ld4(x + f.rowm + pz + f.x0);
      if (CFLQ) yp = ld4(x + f.rowp + pz + f.x0);
      xzp = ld4(x + ozp);
      ux = ld4(ui + o);
      uy = ld4(ui + g.sc + o);
      if (CFLQ) uyp = ld4(ui + g.sc + f.rowp + pz + f.x0);
      uzp = ld4(ui + 2 * g.sc + ozp);
      if (f.lane == 0) el = x[f.row + pz + f.xl];
      if (CFLQ && (f.lane == 31 || f.lastgrp)) {
        xr = x[f.row + pz + f.xr];
        uxr = ui[f.row + pz + f.xr];
      }
    }
    float left = __shfl_up_sync(FULLMASK, xc.w, 1);
    if (f.lane == 0) left = el;
    // corrected velocity on the own cells (f_correct)
    float4 a = ux, b = uy, d = uz;
    a.x -= L0 * (xc.x - left);
    a.y -= L0 * (xc.y - xc.x);
    a.z -= L0 * (xc.z - xc.y);
    a.w -= L0 * (xc.w - xc.z);
    b.x -= L1 * (xc.x - ym.x);
    b.y -= L1 * (xc.y - ym.y);
    b.z -= L1 * (xc.z - ym.z);
    b.w -= L1 * (xc.w - ym.w);
    d.x -= L2 * (xc.x - zm.x);
    d.y -= L2 * (xc.y - zm.y);
    d.z -= L2 * (xc.z - zm.z);
    d.w -= L2 * (xc.w - zm.w);
    // corrected velocity one cell up in x, y, z (what the owner of that cell computes)
    float ar = __shfl_down_sync(FULLMASK, a.x, 1);
    if (f.lane == 31 || f.lastgrp) ar = uxr - L0 * (xr - xc.w);
    float4 bu = uyp, du = uzp;
    bu.x -= L1 * (yp.x - xc.x);
    bu.y -= L1 * (yp.y - xc.y);
    bu.z -= L1 * (yp.z - xc.z);
    bu.w -= L1 * (yp.w - xc.w);
    du.x -= L2 * (xzp.x - xc.x);
    du.y -= L2 * (xzp.y - xc.y);
    du.z -= L2 * (xzp.z - xc.z);
    du.w -= L2 * (xzp.w - xc.w);
    if (f.on) {
      st4(uo + o, a);
      st4(uo + g.sc + o, b);
      st4(uo + 2 * g.sc + o, d);
      ra.add4(a);
      ra.add4(b);
      ra.add4(d);
      st4(p + o, make_float4(xc.x / dt, xc.y / dt, xc.z / dt, xc.w / dt));
      float4 s = f4zero();  // flux_out (f_cfl)
      if (CFLQ) {
      s.x = 0.f + (fmaxf(0.f, a.y) + fmaxf(0.f, -a.x));
      s.x += fmaxf(0.f, bu.x) + fmaxf(0.f, -b.x);
      s.x += fmaxf(0.f, du.x) + fmaxf(0.f, -d.x);
      s.y = 0.f + (fmaxf(0.f, a.z) + fmaxf(0.f, -a.y));
      s.y += fmaxf(0.f, bu.y) + fmaxf(0.f, -b.y);
      s.y += fmaxf(0.f, du.y) + fmaxf(0.f, -d.y);
      s.z = 0.f + (fmaxf(0.f, a.w) + fmaxf(0.f, -a.z));
      s.z += fmaxf(0.f, bu.z) + fmaxf(0.f, -b.z);
      s.z += fmaxf(0.f, du.z) + fmaxf(0.f, -d.z);
      s.w = 0.f + (fmaxf(0.f, ar) + fmaxf(0.f, -a.w));
      s.w += fmaxf(0.f, bu.w) + fmaxf(0.f, -b.w);
      s.w += fmaxf(0.f, du.w) + fmaxf(0.f, -d.w);
      m = fmax(m, (double)fmaxf(fmaxf(s.x, s.y), fmaxf(s.z, s.w)));
      }
    }
    zm = xc;
    xc = xzp;
    uz = uzp;
  }
  ra.publish(flags);
  if (!CFLQ) return;  // plain out-of-place correction (predictor): measurably faster than correcting u in place
  double v[1] = {m}, fin[1];
  if (grid_reduce<RED_MAX, 1>(v, R, slot, fin) && finalize) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      const float mm = (float)fmax(fin[0], R.out[slot_ghost]);
      *dt_out = fminf(10.f, 1.f / (mm + 5.f * nu));
    }
  }
}

// Semi-uniform flags (Coef::semi): one CUDA block per march block (grid = the level's march grid).  The march kernels of that
// block read L[I,d] on the block's cells and L[I+δ_d,d] one cell further in direction d; the flag is 1 iff all of them equal the
// fluid value Lc[d], or 0 on a wall face — exactly what a body-free region holds after BC!(L,0) (src/Flow.jl:145).
__global__ void __launch_bounds__(256) k_semi_flags(const __grid_constant__ Grid g, const float* __restrict__ L, float L0, float L1, float L2, int zchunk,
                                                    unsigned char* __restrict__ flags) {
  pdl_wait();
  const int xlo = 1 + 128 * blockIdx.x, ylo = 1 + FTY * blockIdx.y, zlo = 1 + zchunk * blockIdx.z;
  const int xhi = min(xlo + 128, g.N[0] - 1), yhi = min(ylo + FTY, g.N[1] - 1), zhi = min(zlo + zchunk, g.N[2] - 1);  // exclusive, core
  const float Lc[3] = {L0, L1, L2};
  auto expect = [&](int d, int idx) -> float {
    const bool wall = !g.per[d] && (idx <= 1 ? !(d == 2 && g.zopen[0]) : (idx >= g.N[d] - 1 ? !(d == 2 && g.zopen[1]) : false));
    return wall ? 0.f : Lc[d];
  };
  const int nx = xhi - xlo + 1, ny = yhi - ylo + 1, nz = zhi - zlo + 1;  // one extra layer in each direction
  int ok = 1;
  for (int q = threadIdx.x; q < nx * ny * nz; q += blockDim.x) {
    const int i = q % nx, j = (q / nx) % ny, k = q / (nx * ny);
    const int x = xlo + i, y = ylo + j, z = zlo + k;
    const bool ex = x == xhi, ey = y == yhi, ez = z == zhi;
    if ((int)ex + (int)ey + (int)ez > 1) continue;  // edges / corners of the extra layers are never read
    const i64 o = (i64)g.xo + x + g.s[1] * y + g.s[2] * z;
    const int I[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const bool ext = ex || ey || ez;
      const bool mine = (d == 0 && ex) || (d == 1 && ey) || (d == 2 && ez);
      if (ext && !mine) continue;  // an extra layer is read for its own direction only
      if (L[o + g.sc * d] != expect(d, I[d])) ok = 0;
    }
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0) flags[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)] = (unsigned char)ok;
}

// Device check that the uniform-coefficient specialisation is legal: μ₀ ≡ 1 on every cell the operator reads, μ₁ ≡ 0, V ≡ 0.
// Writes the number of offending values (as a max-reduced 0/1 flag) to out[slot].
__global__ void __launch_bounds__(256) k_check_uniform(const float* __restrict__ mu0, const float* __restrict__ mu1, const float* __restrict__ V, Grid g,
                                                       RedBuf R, int slot) {
  pdl_wait();
  // whole allocation incl. padding is zero-initialised: test only real cells
  const i64 ncell = (i64)g.N[0] * g.N[1] * g.N[2];
  double bad = 0.0;
  for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += (i64)gridDim.x * blockDim.x) {
    const int i = (int)(q % g.N[0]);
    const i64 t = q / g.N[0];
    const int j = (int)(t % g.N[1]);
    const int k = (int)(t / g.N[1]);
    const i64 o = (i64)g.xo + i + g.s[1] * j + g.s[2] * k;
    for (int d = 0; d < 3; d++) {
      if (mu0[o + g.sc * d] != 1.f) bad = 1.0;
      if (V[o + g.sc * d] != 0.f) bad = 1.0;
    }
    for (int d = 0; d < 9; d++)
      if (mu1[o + g.sc * d] != 0.f) bad = 1.0;
  }
  double v[1] = {bad}, fin[1];
  grid_reduce<RED_MAX, 1>(v, R, slot, fin);
}

// ================================================================================================
// Momentum: conv_diff! in gather form with shared fluxes (src/Flow.jl:38-62), fused with BDIM
// (src/Flow.jl:176-180) and scale_u! (src/Flow.jl:211-214).
//
// Geometry: blockDim = (32, CTY); a thread owns one (x,y) column and marches over a z chunk.  The three
// velocity components of the six planes z-2 … z+3 live in a shared-memory ring with an in-plane halo of 2,
// loaded with periodic wrap (ghost cells are exact periodic copies after BC!, so the wrapped loads make the
// reference's lowerBoundary!/upperBoundary! periodic variants identical to the inner formula).  Each face flux
// is evaluated once where that is free: the lower z flux is carried over from the previous plane and the upper x
// flux comes from the next lane by shuffle; the upper y flux is recomputed.
//
// FUSE = uniform mode (no body, periodic): u_new is produced directly (predictor: u = u⁰+Δt·r; corrector:
// u = (u + u⁰+Δt·r)/2) on interior cells, nothing else is read or written; the stale-Φ values that the reference
// leaves on the upper ghost cells of σ (they enter maximum(σ) in CFL, App. A.9-1) are max-reduced into out[slot]
// from their periodic images instead of being stored.
// !FUSE writes f = u⁰ + Δt·r − V on every cell with all indices ≥ 1 (interior and upper ghost rows, App. A.9-2)
// and the stale Φ on upper ghost cells of σ; lower ghost planes of f are filled by k_f_lowghost.
// ================================================================================================
#define CTY 8
#define CRING 6
#define CW 36
#define CH (CTY + 4)

struct ConvTile {
  float v[CRING][3][CH][CW];
};

template <int LAM>
__device__ __forceinline__ float flux_from(float uf, float um2, float um1, float u0c, float up1, float nu, int variant, int* flag) {
  // variant 0: ϕu (inner / periodic), 1: ϕuL (lower non-periodic boundary), 2: ϕuR (upper non-periodic boundary)
  const float diff = nu * (u0c - um1);
  const bool pos = uf > 0.f;
  float conv;
  if (variant == 1)
    conv = pos ? uf * ((u0c + um1) / 2.f) : uf * limiter_f<LAM>(up1, u0c, um1, flag);
  else if (variant == 2)
    conv = uf < 0.f ? uf * ((u0c + um1) / 2.f) : uf * limiter_f<LAM>(um2, um1, u0c, flag);
  else
    conv = uf * limiter_f<LAM>(pos ? um2 : up1, pos ? um1 : u0c, pos ? u0c : um1, flag);
  return conv - diff;
}

// PER3 = no walls: every face is periodic (local wrap) or, in z, open to a neighbouring slab.  `uext` holds the second halo
// planes of ua beyond open z faces: [side][component][plane] (z = -1 and z = N2).
template <int LAM, bool FUSE, bool PER3>
__global__ void __launch_bounds__(32 * CTY) fm_conv(Grid g, const float* __restrict__ ua, const float* __restrict__ u0, const float* __restrict__ V,
                                                    float* __restrict__ out, float* __restrict__ sigma, const float* __restrict__ dtp, float nu, int zchunk,
                                                    int corrector, RedBuf R, int slot, const float* __restrict__ uext, int* __restrict__ flag, const Force fc) {
  pdl_wait();
  extern __shared__ float smem_raw[];
  float* const T = smem_raw;  // [CRING][3][CH][CW]
  constexpr int PL = CH * CW;  // one component plane
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tid = lane + 32 * ty;
  const int xb = 1 + 32 * blockIdx.x, yb = 1 + CTY * blockIdx.y;
  const int x = xb + lane, y = yb + ty;
  const int XM = FUSE ? g.N[0] - 2 : g.N[0] - 1, YM = FUSE ? g.N[1] - 2 : g.N[1] - 1, ZM = FUSE ? g.N[2] - 2 : g.N[2] - 1;
  const int z0 = 1 + zchunk * blockIdx.z, z1 = min(z0 + zchunk, ZM + 1);
  const bool on = x <= XM && y <= YM;
  const float dt = *dtp;

  auto wrap = [&](int v, int d) -> int {  // periodic image inside [0, N-1]; clamp otherwise (clamped values are never used)
    const int N = g.N[d];
    if (g.per[d]) {
      // FUSE (uniform mode) also maps the ghost cells themselves to their interior images, so the tile never depends on the
      // periodic ghosts of u being current (deferred BC!); otherwise a ghost is read as stored: with exitBC! it legitimately
      // differs from its image (the exit plane is rewritten after BC!, src/Flow.jl:194-195)
      if (v < (FUSE ? 1 : 0)) v += N - 2;
      else if (v > N - (FUSE ? 2 : 1)) v -= N - 2;
    }
    return max(0, min(N - 1, v));
  };
  // tile elements this thread loads on every plane (fixed across planes): in-plane global offset and shared index
  constexpr int NE = (PL + 32 * CTY - 1) / (32 * CTY);
  i64 go[NE];
  int so[NE];
#pragma unroll
  for (int k = 0; k < NE; k++) {
    const int e = tid + k * 32 * CTY;
    const int ry = e / CW, rx = e - ry * CW;
    so[k] = e < PL ? e : -1;
    go[k] = (i64)g.xo + wrap(xb - 2 + rx, 0) + g.s[1] * wrap(yb - 2 + min(ry, CH - 1), 1);
  }
  auto load_plane = [&](int zz) {
    int slotp = zz % CRING;
    if (slotp < 0) slotp += CRING;
    float* dst = T + slotp * 3 * PL;
    const float* src = ua;
    i64 pz = g.s[2] * wrap(zz, 2), cs = g.sc;
    if (zz < 0 && g.zopen[0]) {  // second halo plane below an open face
      src = uext;
      pz = 0;
      cs = g.s[2];
    } else if (zz > g.N[2] - 1 && g.zopen[1]) {
      src = uext + 3 * g.s[2];
      pz = 0;
      cs = g.s[2];
    }
#pragma unroll
    for (int k = 0; k < NE; k++) {
      if (so[k] >= 0) {
        const i64 o = go[k] + pz;
        dst[so[k]] = src[o];
        dst[PL + so[k]] = src[o + cs];
        dst[2 * PL + so[k]] = src[o + 2 * cs];
      }
    }
  };
  // po[k]: offset of this thread's column on plane z-2+k (k = 0..5) inside the ring; refreshed every step
  int po[6] = {0, 0, 0, 0, 0, 0};
  const int colbase = (ty + 2) * CW + lane + 2;
  // U(c, dx, dy, dz): component c at (x+dx, y+dy) on plane z+dz, dz ∈ [-2, 3]
  auto U = [&](int c, int dx, int dy, int dz) -> float { return T[po[dz + 2] + c * PL + dy * CW + dx]; };
  int zc = 0;
  // lower-face flux of component i in direction j for the cell at offset (ox,oy,oz) from this thread's cell on plane zc
  auto flux = [&](int i, int j, int ox, int oy, int oz) -> float {
    int variant = 0;
    if (!PER3) {
      const int I[3] = {x + ox, y + oy, zc + oz};
      if (!g.per[j]) {
        const bool walllo = !(j == 2 && g.zopen[0]), wallhi = !(j == 2 && g.zopen[1]);
        variant = (I[j] == 1 && walllo) ? 1 : ((I[j] == g.N[j] - 1 && wallhi) ? 2 : 0);
      }
    }
    const int dx = (j == 0), dy = (j == 1), dz = (j == 2);
    const int ix = (i == 0), iy = (i == 1), iz = (i == 2);
    const float uf = (U(j, ox, oy, oz) + U(j, ox - ix, oy - iy, oz - iz)) / 2.f;
    const float um2 = U(i, ox - 2 * dx, oy - 2 * dy, oz - 2 * dz);
    const float um1 = U(i, ox - dx, oy - dy, oz - dz);
    const float u0c = U(i, ox, oy, oz);
    const float up1 = U(i, ox + dx, oy + dy, oz + dz);
    return flux_from<LAM>(uf, um2, um1, u0c, up1, nu, variant, flag);
  };

  // runtime-component x flux at the cell `ox` columns to the right (the one extra face a warp needs, computed by 3 lanes at once)
  auto flux_x_rt = [&](int i, int ox) -> float {
    int variant = 0;
    if (!PER3 && !g.per[0]) variant = (x + ox == 1) ? 1 : ((x + ox == g.N[0] - 1) ? 2 : 0);
    const int pc = po[2] + ox;
    const int pn = (i == 2 ? po[1] : po[2]) + ox - (i == 0) - (i == 1) * CW;  // I − δ_i
    const float uf = (T[pc] + T[pn]) / 2.f;
    const int pi = pc + i * PL;
    return flux_from<LAM>(uf, T[pi - 2], T[pi - 1], T[pi], T[pi + 1], nu, variant, flag);
  };
  auto set_planes = [&](int z) {
    zc = z;
    int sl = (z - 2) % CRING;
    if (sl < 0) sl += CRING;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      po[k] = sl * 3 * PL + colbase;
      sl = (sl + 1 == CRING) ? 0 : sl + 1;
    }
  };
  // lower y fluxes are computed one plane ahead and shared between the rows (warps) of the block through shared memory
  float* const Fy = T + CRING * 3 * PL;  // [2][3][CTY][32]
  auto fy_at = [&](int buf, int i, int row) -> float& { return Fy[((buf * 3 + i) * CTY + row) * 32 + lane]; };

  for (int zz = z0 - 2; zz <= z0 + 1; zz++) load_plane(zz);
  __syncthreads();
  set_planes(z0);
  float Fyl[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Fyl[i] = flux(i, 1, 0, 0, 0);
    fy_at(z0 & 1, i, ty) = Fyl[i];
  }
  float Fz[3] = {0.f, 0.f, 0.f};
  bool haveFz = false;
  double gmax = 0.0;
  for (int z = z0; z < z1; z++) {
    load_plane(z + 2);
    set_planes(z);
    __syncthreads();
    // which directions contribute to this cell (all for interior cells; upper ghost rows only get the others)
    const bool ax = FUSE || x <= g.N[0] - 2, ay = FUSE || y <= g.N[1] - 2, az = FUSE || z <= g.N[2] - 2;
    float Fxlo[3], Fxhi[3];
#pragma unroll
    for (int i = 0; i < 3; i++) Fxlo[i] = flux(i, 0, 0, 0, 0);
    {
      // the face beyond the warp's last cell: lanes 0-2 compute one component each, lane 31 collects them
      float fex = 0.f;
      if (lane < 3) fex = flux_x_rt(lane, 32 - lane);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const float e = __shfl_sync(FULLMASK, fex, i);
        Fxhi[i] = __shfl_down_sync(FULLMASK, Fxlo[i], 1);
        if (lane == 31) Fxhi[i] = e;
      }
    }
    if (x == XM && lane != 31) {  // partial warp at the end of a row
#pragma unroll
      for (int i = 0; i < 3; i++) Fxhi[i] = (FUSE || x + 1 <= g.N[0] - 1) ? flux(i, 0, 1, 0, 0) : 0.f;
    }
    if (!haveFz) {
#pragma unroll
      for (int i = 0; i < 3; i++) Fz[i] = flux(i, 2, 0, 0, 0);
      haveFz = true;
    }
    float Fzhi[3] = {0.f, 0.f, 0.f};
    if (FUSE || z + 1 <= g.N[2] - 1) {
#pragma unroll
      for (int i = 0; i < 3; i++) Fzhi[i] = flux(i, 2, 0, 0, 1);
    }
    // upper y flux = the next row's lower flux (shared); the last row of the block computes its own
    float Fyhi[3], Fyn[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Fyhi[i] = (ty + 1 < CTY) ? fy_at(z & 1, i, ty + 1) : flux(i, 1, 0, 1, 0);
      Fyn[i] = flux(i, 1, 0, 0, 1);  // lower y flux of plane z+1, for the next step
      fy_at((z + 1) & 1, i, ty) = Fyn[i];
    }
    float F2lo_y = 0.f;
    if (on) {
      const i64 o = (i64)g.xo + x + g.s[1] * y + g.s[2] * z;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float r = 0.f;
        const float Fylo = Fyl[i];
        if (i == 2) F2lo_y = Fylo;
        if (ax) {
          r += Fxlo[i];
          r -= Fxhi[i];
        }
        if (ay) {
          r += Fylo;
          r -= Fyhi[i];
        }
        if (az) {
          r += Fz[i];
          r -= Fzhi[i];
        }
        if (fc.on) r += fc.a[i];  // accelerate! (src/Flow.jl:64-73)
        const i64 oc = o + (i64)i * g.sc;
        if (FUSE) {
          const float f = u0[oc] + dt * r;  // − V with V ≡ 0; then X = 0/2 + 0 + 1·f
          out[oc] = corrector ? (U(i, 0, 0, 0) + f) * 0.5f : f;
        } else {
          out[oc] = u0[oc] + dt * r - V[oc];
        }
      }
      if (FUSE) {
        // periodic images of the stale Φ on upper ghost cells (see header): candidates by T = {k : I_k == 1}
        const bool t0 = x == 1, t1 = y == 1, t2 = (z + g.zoff) == 1;
        if (t0 || t1) gmax = fmax(gmax, (double)Fz[2]);
        if (t2) gmax = fmax(gmax, (double)F2lo_y);
        if (t2 && t1) gmax = fmax(gmax, (double)Fxlo[2]);
      } else {
        const bool ghost = x == g.N[0] - 1 || y == g.N[1] - 1 || (z == g.N[2] - 1 && !g.zopen[1]);
        if (ghost) {
          // last Φ written by the reference's (i=D, j) loops: the largest j whose range contains the cell
          const int lo0 = g.per[0] ? 1 : 2, lo1 = g.per[1] ? 1 : 2, lo2 = (g.per[2] || g.zopen[0]) ? 1 : 2;
          if (az && z >= lo2) sigma[o] = Fz[2];
          else if (ay && y >= lo1) sigma[o] = F2lo_y;
          else if (ax && x >= lo0) sigma[o] = Fxlo[2];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Fz[i] = Fzhi[i];
      Fyl[i] = Fyn[i];
    }
  }
  if (FUSE) {
    double v[1] = {gmax}, fin[1];
    grid_reduce<RED_MAX, 1>(v, R, slot, fin);
  }
}

// f on the lower ghost planes (any index 0): r = 0 there, so f = u⁰ + Δt·0 − V  (src/Flow.jl:178 over CartesianIndices(f))
__global__ void k_f_lowghost(Grid g, const float* __restrict__ u0, const float* __restrict__ V, float* __restrict__ f, const float* __restrict__ dtp,
                             const Force fc) {
  pdl_wait();
  const int j = blockIdx.z;  // plane I_j = 0
  if (j == 2 && g.zopen[0]) return;  // slab-internal face: that plane of f arrives by halo exchange
  const int da = (j == 0) ? 1 : 0, db = (j == 2) ? 1 : 2;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y * blockDim.y + threadIdx.y;
  if (t0 >= g.N[da] || t1 >= g.N[db]) return;
  int I[3];
  I[j] = 0;
  I[da] = t0;
  I[db] = t1;
  const i64 o = cell_off(g, I);
  const float dt = *dtp;
  for (int i = 0; i < 3; i++) f[o + g.sc * i] = u0[o + g.sc * i] + dt * (fc.on ? 0.f + fc.a[i] : 0.f) - V[o + g.sc * i];
}

// max of σ over the ghost cells (where the reference's stale Φ lives) → out[slot]; planes selected by blockIdx.z
__global__ void __launch_bounds__(256) k_sigma_ghostmax(Grid g, const float* __restrict__ sigma, RedBuf R, int slot) {
  pdl_wait();
  const int plane = blockIdx.z;
  const int j = plane / 2, which = plane % 2;
  const int da = (j == 0) ? 1 : 0, db = (j == 2) ? 1 : 2;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y * blockDim.y + threadIdx.y;
  double v[1] = {0.0}, fin[1];
  if (t0 < g.N[da] && t1 < g.N[db]) {
    int I[3];
    I[j] = which ? g.N[j] - 1 : 0;
    I[da] = t0;
    I[db] = t1;
    v[0] = (double)sigma[cell_off(g, I)];
  }
  grid_reduce<RED_MAX, 1>(v, R, slot, fin);
}

// CFL (src/Flow.jl:234-244) on the interior with the march geometry: σ = flux_out(I,u) (stored unless UNI), max-reduced,
// combined with the ghost maximum in out[slot_ghost]; the last block stores Δt = min(10, 1/(max+5ν)) to dt_out.
template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, MARCH_MINB) f_cfl(Grid g, const float* __restrict__ u, float* __restrict__ sigma, float nu, float* __restrict__ dt_out,
                                                  int zchunk, RedBuf R, int slot, int slot_ghost, int finalize) {
  pdl_wait();
  const Frame f = make_frame(g, zchunk);
  double m = 0.0;
  for (int z = f.z0; z < f.z1; z++) {
    const i64 pz = g.s[2] * z;
    const i64 o = f.row + pz + f.x0;
    float4 ux = f4zero();
    float e = 0.f;
    if (f.on) {
      ux = ld4(u + o);
      if (f.lane == 31 || f.lastgrp) e = u[f.row + pz + f.xr];
    }
    float uxr = __shfl_down_sync(FULLMASK, ux.x, 1);
    if (f.lane == 31 || f.lastgrp) uxr = e;
    if (f.on) {
      // upper neighbours through the periodic wrap (see f_div_residual)
      const float4 uy = ld4(u + g.sc + o), uyp = ld4(u + g.sc + f.rowp + pz + f.x0);
      const float4 uz = ld4(u + 2 * g.sc + o), uzp = ld4(u + 2 * g.sc + f.row + g.s[2] * zwrap_hi(g, z) + f.x0);
      float4 s;
      s.x = 0.f + (fmaxf(0.f, ux.y) + fmaxf(0.f, -ux.x));
      s.x += fmaxf(0.f, uyp.x) + fmaxf(0.f, -uy.x);
      s.x += fmaxf(0.f, uzp.x) + fmaxf(0.f, -uz.x);
      s.y = 0.f + (fmaxf(0.f, ux.z) + fmaxf(0.f, -ux.y));
      s.y += fmaxf(0.f, uyp.y) + fmaxf(0.f, -uy.y);
      s.y += fmaxf(0.f, uzp.y) + fmaxf(0.f, -uz.y);
      s.z = 0.f + (fmaxf(0.f, ux.w) + fmaxf(0.f, -ux.z));
      s.z += fmaxf(0.f, uyp.z) + fmaxf(0.f, -uy.z);
      s.z += fmaxf(0.f, uzp.z) + fmaxf(0.f, -uz.z);
      s.w = 0.f + (fmaxf(0.f, uxr) + fmaxf(0.f, -ux.w));
      s.w += fmaxf(0.f, uyp.w) + fmaxf(0.f, -uy.w);
      s.w += fmaxf(0.f, uzp.w) + fmaxf(0.f, -uz.w);
      if (!UNI) st4(sigma + o, s);
      m = fmax(m, (double)fmaxf(fmaxf(s.x, s.y), fmaxf(s.z, s.w)));
    }
  }
  double v[1] = {m}, fin[1];
  if (grid_reduce<RED_MAX, 1>(v, R, slot, fin) && finalize) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      const float mm = (float)fmax(fin[0], R.out[slot_ghost]);
      *dt_out = fminf(10.f, 1.f / (mm + 5.f * nu));
    }
  }
}

// ================================================================================================
// GaussSeidelRB!(p;it=4,ω) (src/Poisson.jl:141-148) in three march kernels instead of six launches:
//   f_gs_a : ϵ⁰ = r·iD, sweep 1                      (reads r, writes ϵ)
//   f_gs_b : sweeps 2 and 3                           (reads ϵ, r; writes ϵ)
//   f_gs_c : sweep 4, increment!(ω), L₂               (reads ϵ, r, x; writes r, x)
// Colours: sweep k₀ updates the cells with (x+y+z) ≡ k₀ (mod 2) in 0-based indices (Σ 1-based ≡ 1+k₀, :125).
// "A" cells (odd) move in sweeps 1 and 3, "B" cells (even) in sweeps 2 and 4.  Within one kernel the second
// half-sweep needs first-half-sweep values of neighbouring cells; they are RECOMPUTED from the stored field
// (same operations, same order ⇒ same bits) instead of being exchanged, so no kernel needs a grid-wide
// dependency and every field is streamed once per kernel.
// The reference fills the periodic ghosts of ϵ once, before the sweeps (:143), so during all four sweeps a
// neighbour across a periodic face is the stale ϵ⁰ = r·iD of the wrapped cell; increment! refreshes the ghosts
// (:101), so the final Aϵ sees current values.  `stale` below implements exactly that.
// Requires the march geometry, (N2-2) even (half_rangek reaches every interior plane).
// ================================================================================================
// lanes of colour A (odd x+y+z) in a group starting at x0 on row (y,z): returns the parity of the first cell
__device__ __forceinline__ bool first_is_A(int x0, int y, int z) { return ((x0 + y + z) & 1) != 0; }
__device__ __forceinline__ float4 select_A(bool firstA, const float4& a, const float4& b) {  // A lanes from a, B lanes from b
  return firstA ? make_float4(a.x, b.y, a.z, b.w) : make_float4(b.x, a.y, b.z, a.w);
}

template <bool UNI>
struct Gs {
  const Grid& g;
  const Coef& c;
  const float* eps;
  const float* r;
  int lane, x0, xl, xr;
  bool on, lastgrp, wrapl, wrapr;  // wrapl/wrapr: the row ends of this lane are periodic faces
  bool semi, wx0, wx1, wy0, wy1;   // semi-uniform block and its wall faces (Frame)

  __device__ __forceinline__ Gs(const Grid& g_, const Coef& c_, const float* eps_, const float* r_, const Frame& f)
      : g(g_), c(c_), eps(eps_), r(r_), lane(f.lane), x0(f.x0), xl(f.xl), xr(f.xr), on(f.on), lastgrp(f.lastgrp),
        wrapl(g_.per[0] && f.x0 == 1), wrapr(g_.per[0] && f.x0 + 3 == g_.N[0] - 2), semi(f.semi), wx0(f.wx0), wx1(f.wx1), wy0(f.wy0), wy1(f.wy1) {}

  __device__ __forceinline__ i64 off(int y, int z) const { return (i64)g.xo + g.s[1] * y + g.s[2] * z; }
  __device__ __forceinline__ float iD1(i64 o) const { return UNI ? c.iDc : c.iD[o]; }
  __device__ __forceinline__ float4 iD4(i64 o) const { return UNI ? make_float4(c.iDc, c.iDc, c.iDc, c.iDc) : ld4(c.iD + o); }
  __device__ __forceinline__ float stale1(i64 o) const { return r[o] * iD1(o); }
  __device__ __forceinline__ float4 stale4(i64 o) const { return mul4(ld4(r + o), iD4(o)); }

  // gauss(I,r,L,iD,x) (src/Poisson.jl:116-122) for the 4 cells of the lane: (r − Σ_d (lo·L[I,d] + hi·L[I+δ_d,d]))·iD
  __device__ __forceinline__ float4 gauss4(i64 o, int z, const float4& rr, const float4& ec, float left, float right, const float4& ym, const float4& yp,
                                           const float4& zm, const float4& zp) const {
    float4 s = rr;
    if (UNI) {
      const float L0 = c.Lc[0], L1 = c.Lc[1], L2 = c.Lc[2];
      s.x -= left * L0 + ec.y * L0;
      s.x -= ym.x * L1 + yp.x * L1;
      s.x -= zm.x * L2 + zp.x * L2;
      s.y -= ec.x * L0 + ec.z * L0;
      s.y -= ym.y * L1 + yp.y * L1;
      s.y -= zm.y * L2 + zp.y * L2;
      s.z -= ec.y * L0 + ec.w * L0;
      s.z -= ym.z * L1 + yp.z * L1;
      s.z -= zm.z * L2 + zp.z * L2;
      s.w -= ec.z * L0 + right * L0;
      s.w -= ym.w * L1 + yp.w * L1;
      s.w -= zm.w * L2 + zp.w * L2;
    } else {
      float4 A0, A1, B1, A2, B2;
      float A0r;
      if (semi) {  // no body near this block: the fluid value on every face, 0 on wall faces
        A0 = splat4(c.Lc[0]);
        if (wx0) A0.x = 0.f;
        A0r = wx1 ? 0.f : c.Lc[0];
        A1 = splat4(wy0 ? 0.f : c.Lc[1]);
        B1 = splat4(wy1 ? 0.f : c.Lc[1]);
        A2 = splat4(wall_zlo(g, z) ? 0.f : c.Lc[2]);
        B2 = splat4(wall_zhi(g, z) ? 0.f : c.Lc[2]);
      } else {
        A0 = ld4(c.L + o);
        A0r = c.L[o + 4];
        A1 = ld4(c.L + g.sc + o);
        B1 = ld4(c.L + g.sc + o + g.s[1]);
        A2 = ld4(c.L + 2 * g.sc + o);
        B2 = ld4(c.L + 2 * g.sc + o + g.s[2]);
      }
      s.x -= left * A0.x + ec.y * A0.y;
      s.x -= ym.x * A1.x + yp.x * B1.x;
      s.x -= zm.x * A2.x + zp.x * B2.x;
      s.y -= ec.x * A0.y + ec.z * A0.z;
      s.y -= ym.y * A1.y + yp.y * B1.y;
      s.y -= zm.y * A2.y + zp.y * B2.y;
      s.z -= ec.y * A0.z + ec.w * A0.w;
      s.z -= ym.z * A1.z + yp.z * B1.z;
      s.z -= zm.z * A2.z + zp.z * B2.z;
      s.w -= ec.z * A0.w + right * A0r;
      s.w -= ym.w * A1.w + yp.w * B1.w;
      s.w -= zm.w * A2.w + zp.w * B2.w;
    }
    return mul4(s, iD4(o));
  }
  // One half-sweep update of ALL four cells of the lane's group on the absolute interior row (y,z), reading the STORED field
  // `eps` (the caller keeps only the lanes of the colour that moves).  Periodic faces read the stale r·iD.  All lanes must call.
  __device__ __forceinline__ float4 row_update(int y, int z, float4& stored) const {
    const i64 ro = off(y, z);
    const i64 o = ro + x0;
    float4 ec = f4zero(), rr = f4zero(), vym = f4zero(), vyp = f4zero(), vzm = f4zero(), vzp = f4zero();
    float el = 0.f, er = 0.f;
    if (on) {
      ec = ld4(eps + o);
      rr = ld4(r + o);
      if (lane == 0) el = wrapl ? stale1(ro + xl) : eps[ro + xl];
      if (lane == 31 || lastgrp) er = wrapr ? stale1(ro + xr) : eps[ro + xr];
      const bool wym = g.per[1] && y == 1, wyp = g.per[1] && y == g.N[1] - 2;
      // stale across the global periodic z boundary: a local wrap (per[2]) or, in a z-slab, the exchanged ghost plane of r
      const bool wzm = (g.per[2] || g.zstale[0]) && z == 1, wzp = (g.per[2] || g.zstale[1]) && z == g.N[2] - 2;
      const i64 oym = off(wym ? g.N[1] - 2 : y - 1, z) + x0, oyp = off(wyp ? 1 : y + 1, z) + x0;
      const i64 ozm = off(y, (g.per[2] && z == 1) ? g.N[2] - 2 : z - 1) + x0, ozp = off(y, (g.per[2] && z == g.N[2] - 2) ? 1 : z + 1) + x0;
      vym = wym ? stale4(oym) : ld4(eps + oym);
      vyp = wyp ? stale4(oyp) : ld4(eps + oyp);
      vzm = wzm ? stale4(ozm) : ld4(eps + ozm);
      vzp = wzp ? stale4(ozp) : ld4(eps + ozp);
    }
    stored = ec;
    float left = __shfl_up_sync(FULLMASK, ec.w, 1), right = __shfl_down_sync(FULLMASK, ec.x, 1);
    if (lane == 0) left = el;
    if (lane == 31 || lastgrp) right = er;
    if (!on) return f4zero();
    return gauss4(o, z, rr, ec, left, right, vym, vyp, vzm, vzp);
  }
};

// f_gs_a: ϵ⁰ = r·iD everywhere; A cells take sweep 1 (their neighbours are ϵ⁰, stale and fresh coincide).
template <bool UNI>
__device__ __forceinline__ void b_f_gs_a(const Grid& g, const Coef& c, const float* r,
                                                   float* eps, int zchunk, const int3 vb) {
  const Frame f = make_frame(g, zchunk, vb, c);
  const Gs<UNI> G(g, c, eps, r, f);
  SField F;
  F.q = r;
  F.w = c.iD;
  F.s = c.iDc;
  F.kind = UNI ? 1 : 2;
  march7(g, f, F, [&](int z, i64 pz, i64 o, const float4& e, float left, float right, const float4& ym, const float4& yp, const float4& zm,
                      const float4& zp) {
    if (f.on) {
      const float4 up = G.gauss4(o, z, ld4(r + o), e, left, right, ym, yp, zm, zp);
      st4(eps + o, select_A(first_is_A(f.x0, f.y, z), up, e));
    }
  });
}
template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, (UNI ? MARCH_MINB : MARCH_MINB_GEN)) f_gs_a(const __grid_constant__ Grid g, const __grid_constant__ Coef c, const float* __restrict__ r,
                                                   float* __restrict__ eps, int zchunk) {
  pdl_wait();
  b_f_gs_a<UNI>(g, c, r, eps, zchunk, real_block());
}

// f_gs_half: one red/black half-sweep (src/Poisson.jl:145) in place: the cells with (x+y+z) ≡ k₀ (mod 2) move.  In-place is race-free:
// a moving cell reads only cells of the other colour (or stale r·iD across periodic faces), which nobody writes in this launch; the
// vector store rewrites the other colour's lanes with the values just loaded.
template <bool UNI>
__device__ __forceinline__ void b_f_gs_half(const Grid& g, const Coef& c, const float* r,
                                                      float* eps, int k0, int zchunk, const int3 vb) {
  const Frame f = make_frame(g, zchunk, vb, c);
  const Gs<UNI> G(g, c, eps, r, f);
  const int y = min(f.y, g.N[1] - 2);
  const bool moveA = (k0 & 1) != 0;
  for (int z = f.z0; z < f.z1; z++) {
    float4 st;
    const float4 up = G.row_update(y, z, st);
    if (f.on) {
      const bool firstA = first_is_A(f.x0, y, z);
      // select_A(firstA, a, b): A lanes from a, B lanes from b
      st4(eps + G.off(y, z) + f.x0, moveA ? select_A(firstA, up, st) : select_A(firstA, st, up));
    }
  }
}
template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, (UNI ? MARCH_MINB : MARCH_MINB_GEN)) f_gs_half(const __grid_constant__ Grid g, const __grid_constant__ Coef c, const float* __restrict__ r,
                                                      float* eps, int k0, int zchunk) {
  pdl_wait();
  b_f_gs_half<UNI>(g, c, r, eps, k0, zchunk, real_block());
}

// ================================================================================================
// k_small_levels — the coarse end of a V-cycle in ONE cooperative launch.
// Levels of a few hundred thousand cells and fewer are launch-latency bound: nine launches per level per V-cycle, each a few
// microseconds of work.  This kernel runs the same kernel BODIES (b_f_jacobi, b_f_gs_a, … — identical arithmetic, identical
// bits) over the virtual blocks of each small grid and separates the operations with grid-wide barriers instead of launches.
// The host flattens the recursion of Vcycle! (src/MultiLevelPoisson.jl:88-101) into a list of SmallOp.
// ================================================================================================
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

enum { OP_F_JACOBI = 0, OP_F_GSA, OP_F_GSHALF, OP_F_INC, OP_F_PROLONG, OP_K_JACOBI, OP_K_RESTRICT, OP_K_GSINIT, OP_K_GSSWEEP, OP_K_INC, OP_K_PROLONG };

struct SmallOp {
  int type, k0, x_is_zero, do_restrict, zoffc, zchunk;
  int vg[3];  // virtual grid of 32×8(×1)-thread blocks
  int cm[3];  // coarsening mask towards `gc`
  Grid g;     // the level the op runs on
  Coef c;
  Grid gc;    // the other level of a restriction / prolongation
  Lvl lvl;    // pointer bundle for the general bodies
  Box box;
  float* other;  // coarse r (restriction target) or coarse x (prolongation source)
};

template <bool UNI>
__global__ void __launch_bounds__(32 * FTY, 4) k_small_levels(const SmallOp* __restrict__ ops, int nops, const float* wp) {
  pdl_wait();
  cg::grid_group grid = cg::this_grid();
  __shared__ SmallOp op;
  const int tid = threadIdx.x + 32 * threadIdx.y;
  RedBuf nored{nullptr, nullptr, nullptr};
  // the descriptor of the NEXT operation is fetched into a register while the current one runs (one int per thread: the global
  // load's latency hides behind the work and the grid barrier) and stored to shared memory at the head of its own iteration
  constexpr int NINT = (int)(sizeof(SmallOp) / sizeof(int));
  static_assert(NINT <= 32 * FTY, "one descriptor word per thread");
  int pre = 0;
  if (nops > 0 && tid < NINT) pre = reinterpret_cast<const int*>(ops)[tid];
  for (int o = 0; o < nops; o++) {
    __syncthreads();
    if (tid < NINT) reinterpret_cast<int*>(&op)[tid] = pre;
    __syncthreads();
    if (o + 1 < nops && tid < NINT) pre = reinterpret_cast<const int*>(ops + o + 1)[tid];
    const int nvb = op.vg[0] * op.vg[1] * op.vg[2];
    for (int v = blockIdx.x; v < nvb; v += gridDim.x) {
      const int3 vb = make_int3(v % op.vg[0], (v / op.vg[0]) % op.vg[1], v / (op.vg[0] * op.vg[1]));
      switch (op.type) {
        case OP_F_JACOBI:
          b_f_jacobi<UNI>(op.g, op.c, op.lvl.r, op.lvl.r2, op.lvl.x, op.x_is_zero, op.zchunk, op.gc, op.other, op.do_restrict, op.zoffc, vb);
          break;
        case OP_F_GSA:
          b_f_gs_a<UNI>(op.g, op.c, op.lvl.r, op.lvl.eps, op.zchunk, vb);
          break;
        case OP_F_GSHALF:
          b_f_gs_half<UNI>(op.g, op.c, op.lvl.r, op.lvl.eps, op.k0, op.zchunk, vb);
          break;
        case OP_F_INC: {
          ProlongSrc ps{nullptr, op.g, 0, 0, 0};
          b_f_increment<UNI, false>(op.g, op.c, op.lvl.eps, ps, op.lvl.r, op.lvl.x, wp, op.x_is_zero, op.zchunk, 0, nored, 0, vb);
        } break;
        case OP_F_PROLONG: {
          ProlongSrc ps{op.other, op.gc, 0, 0, 0};
          b_f_increment<UNI, true>(op.g, op.c, nullptr, ps, op.lvl.r, op.lvl.x, wp, 0, op.zchunk, 0, nored, 0, vb);
        } break;
        case OP_K_JACOBI:
          b_k_jacobi<3>(op.lvl, op.box, op.x_is_zero, vb);
          break;
        case OP_K_RESTRICT:
          b_k_restrict<3>(op.gc, op.g, op.box, op.other, op.lvl.r, op.cm[0], op.cm[1], op.cm[2], vb);
          break;
        case OP_K_GSINIT:
          b_k_gs_init<3>(op.lvl, op.box, vb);
          break;
        case OP_K_GSSWEEP:
          b_k_gs_sweep<3>(op.lvl, op.box, op.k0, vb);
          break;
        case OP_K_INC:
          b_k_increment<3>(op.lvl, op.box, wp, op.x_is_zero, 0, nored, 0, vb);
          break;
        case OP_K_PROLONG:
          b_k_prolong_inc<3>(op.lvl, op.gc, op.other, op.box, wp, op.cm[0], op.cm[1], op.cm[2], vb);
          break;
      }
      __syncthreads();
    }
    grid.sync();
  }
}

// ================================================================================================
// k_tiny_uni — the innermost levels of the V-cycle (≤ 8192 cells: 18³ and coarser) of uniform mode in ONE block, in shared memory.
// A grid-wide barrier of k_small_levels costs ≈2.5 µs — more than an operation on such a level — and the coarse end of a V-cycle
// is nine operations per level.  Here the levels live in shared memory as dense interior-only arrays (every direction is periodic
// in uniform mode: neighbours wrap, ghosts do not exist), one block of 1024 threads runs
//     Vcycle!(ml; l=T) ; smooth!(levels[T])            (src/MultiLevelPoisson.jl:88-101,106)
// between block barriers, reading r of level T from global memory and leaving x (and r) of level T there.  The arithmetic is that
// of the general bodies in wl_kernels.cuh (b_k_jacobi, b_k_restrict, b_k_gs_init, b_k_gs_sweep, b_k_increment, b_k_prolong_inc)
// with the level's uniform coefficients — same operations, same order, same bits.
// ================================================================================================
#define TINY_MAXLEV 6
struct TinyArgs {
  int nlev;
  int n[TINY_MAXLEV][3];  // interior sizes, all even
  float L[TINY_MAXLEV][3], D[TINY_MAXLEV], iD[TINY_MAXLEV];
  Grid g0;         // layout of level T in global memory
  const float* r0; // its residual (input)
  float* x0;       // its solution (output)
  float* r0out;    // its residual after the smoother (output, for observers)
  const float* wp; // ω
};
struct TinyLvl {
  float *X, *R, *E;
  int n0, n1, n2, cells;
  unsigned m0, m1, mh;  // ⌈2³²/n0⌉, ⌈2³²/n1⌉, ⌈2³²/(n0/2)⌉: c / n = umulhi(c, m) exactly for c·n < 2³² (cells ≤ 8192)
  float L0, L1, L2, D, iD;
  int sh0, shp;  // log2(n0), log2(n0·n1) when both are powers of two and the plane has ≤ 1024 cells, else −1 (see each_cell)
  // cell index → (i, j, k) with two multiply-high divisions
  __device__ __forceinline__ void ijk(int c, int& i, int& j, int& k) const {
    const int row = (int)__umulhi((unsigned)c, m0);
    i = c - row * n0;
    k = (int)__umulhi((unsigned)row, m1);
    j = row - k * n1;
  }
  // body(c, i, j, k) for the cells of the level this thread owns.  Power-of-two planes of ≤ 1024 cells: a thread keeps its (i, j) and
  // walks k — the index arithmetic and the x / y neighbour wraps are loop invariants; otherwise cells are dealt out linearly.
  template <class F>
  __device__ __forceinline__ void each_cell(int tid, F body) const {
    if (shp >= 0) {
      const int c0 = tid & ((1 << shp) - 1), i = c0 & (n0 - 1), j = c0 >> sh0;
      const int kstep = 1024 >> shp;
      for (int k = tid >> shp; k < n2; k += kstep) body(c0 + (k << shp), i, j, k);
    } else {
      for (int c = tid; c < cells; c += 1024) {
        int i, j, k;
        ijk(c, i, j, k);
        body(c, i, j, k);
      }
    }
  }
  // the same over the cells of one colour: body(c, i, j, k) with i + j + k ≡ par (mod 2)
  template <class F>
  __device__ __forceinline__ void each_half(int tid, int par, F body) const {
    const int h0 = n0 >> 1, s1 = n0, s2 = n0 * n1;
    if (shp >= 1) {
      const int hp = shp - 1;  // log2 of a half plane
      const int q0 = tid & ((1 << hp) - 1), ih = q0 & (h0 - 1), j = q0 >> (sh0 - 1);
      const int kstep = 1024 >> hp;
      for (int k = tid >> hp; k < n2; k += kstep) {
        const int i = 2 * ih + ((par + j + k) & 1);
        body(i + s1 * j + s2 * k, i, j, k);
      }
    } else {
      for (int q = tid; q < (cells >> 1); q += 1024) {
        const int row = h0 == 1 ? q : (int)__umulhi((unsigned)q, mh), ih = q - row * h0;
        const int k = n1 == 1 ? row : (int)__umulhi((unsigned)row, m1), j = row - k * n1;
        const int i = 2 * ih + ((par + j + k) & 1);
        body(i + s1 * j + s2 * k, i, j, k);
      }
    }
  }
};
__host__ __device__ inline unsigned tiny_magic(int n) { return n <= 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)n - 1) / (unsigned)n); }
__device__ __forceinline__ int tiny_wrap(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

// GaussSeidelRB!(p; it=4, ω) + increment! (src/Poisson.jl:141-148, 100-104)
__device__ __forceinline__ void tiny_gs(const TinyLvl& l, float w, int x_is_zero, int tid) {
  const int s1 = l.n0, s2 = l.n0 * l.n1;
  for (int c = tid; c < l.cells; c += 1024) l.E[c] = l.R[c] * l.iD;
  __syncthreads();
  for (int k0 = 1; k0 <= 4; k0++) {
    // Σ(1-based indices) ≡ 1+k₀ (mod 2)  ⇔  i+j+k ≡ 1+k₀
    l.each_half(tid, 1 + k0, [&](int c, int i, int j, int k) {
      float s = l.R[c];
      // across a periodic face the sweep sees the stale ϵ⁰ = r·iD of the wrapped cell (perBC! runs once, before the sweeps)
      const float xlo = i == 0 ? l.R[c + (l.n0 - 1)] * l.iD : l.E[c - 1], xhi = i == l.n0 - 1 ? l.R[c - (l.n0 - 1)] * l.iD : l.E[c + 1];
      s -= xlo * l.L0 + xhi * l.L0;
      const float ylo = j == 0 ? l.R[c + s1 * (l.n1 - 1)] * l.iD : l.E[c - s1], yhi = j == l.n1 - 1 ? l.R[c - s1 * (l.n1 - 1)] * l.iD : l.E[c + s1];
      s -= ylo * l.L1 + yhi * l.L1;
      const float zlo = k == 0 ? l.R[c + s2 * (l.n2 - 1)] * l.iD : l.E[c - s2], zhi = k == l.n2 - 1 ? l.R[c - s2 * (l.n2 - 1)] * l.iD : l.E[c + s2];
      s -= zlo * l.L2 + zhi * l.L2;
      l.E[c] = s * l.iD;
    });
    __syncthreads();
  }
  l.each_cell(tid, [&](int c, int i, int j, int k) {
    const float e = l.E[c];
    float Ae = e * l.D;
    Ae += l.E[c - i + tiny_wrap(i - 1, l.n0)] * l.L0 + l.E[c - i + tiny_wrap(i + 1, l.n0)] * l.L0;
    Ae += l.E[c + s1 * (tiny_wrap(j - 1, l.n1) - j)] * l.L1 + l.E[c + s1 * (tiny_wrap(j + 1, l.n1) - j)] * l.L1;
    Ae += l.E[c + s2 * (tiny_wrap(k - 1, l.n2) - k)] * l.L2 + l.E[c + s2 * (tiny_wrap(k + 1, l.n2) - k)] * l.L2;
    l.R[c] = l.R[c] - w * Ae;
    l.X[c] = x_is_zero ? w * e : l.X[c] + w * e;
  });
  __syncthreads();
}

// One copy of each stage's code, looped over the levels (the fully unrolled version was 11 000 instructions = 180 KB that ran once,
// straight through: the kernel spent its time fetching instructions — `no_instruction` was its top stall, 44 µs per launch).
__global__ void __launch_bounds__(1024, 1) k_tiny_uni(const __grid_constant__ TinyArgs a) {
  pdl_wait();
  extern __shared__ float tiny_sm[];
  __shared__ TinyLvl lv[TINY_MAXLEV];
  const int tid = threadIdx.x;
  if (tid == 0) {
    float* p = tiny_sm;
    for (int q = 0; q < a.nlev; q++) {
      TinyLvl& l = lv[q];
      l.n0 = a.n[q][0], l.n1 = a.n[q][1], l.n2 = a.n[q][2];
      l.cells = l.n0 * l.n1 * l.n2;
      l.m0 = tiny_magic(l.n0), l.m1 = tiny_magic(l.n1), l.mh = tiny_magic(l.n0 >> 1);
      l.L0 = a.L[q][0], l.L1 = a.L[q][1], l.L2 = a.L[q][2], l.D = a.D[q], l.iD = a.iD[q];
      const int plane = l.n0 * l.n1;
      const bool pow2 = (l.n0 & (l.n0 - 1)) == 0 && (l.n1 & (l.n1 - 1)) == 0 && plane <= 1024 && l.n0 >= 2;
      l.sh0 = pow2 ? 31 - __clz(l.n0) : -1;
      l.shp = pow2 ? 31 - __clz(plane) : -1;
      l.X = p, l.R = p + l.cells, l.E = p + 2 * l.cells;
      p += 3 * l.cells;
    }
  }
  __syncthreads();
  const float w = *a.wp;
  {  // r of level T from global memory
    const TinyLvl l = lv[0];
    l.each_cell(tid, [&](int c, int i, int j, int k) { l.R[c] = a.r0[(i64)(a.g0.xo + i + 1) + a.g0.s[1] * (j + 1) + a.g0.s[2] * (k + 1)]; });
  }
  __syncthreads();
  // ---- down-stroke: Jacobi! (x = ϵ: the level starts from x = 0) and restrict! ----
#pragma unroll 1
  for (int q = 0; q < a.nlev - 1; q++) {
    TinyLvl l = lv[q];
    const int s1 = l.n0, s2 = l.n0 * l.n1;
    l.each_cell(tid, [&](int c, int i, int j, int k) {
      const float e = l.R[c] * l.iD;
      float s = e * l.D;
      s += (l.R[c - i + tiny_wrap(i - 1, l.n0)] * l.iD) * l.L0 + (l.R[c - i + tiny_wrap(i + 1, l.n0)] * l.iD) * l.L0;
      s += (l.R[c + s1 * (tiny_wrap(j - 1, l.n1) - j)] * l.iD) * l.L1 + (l.R[c + s1 * (tiny_wrap(j + 1, l.n1) - j)] * l.iD) * l.L1;
      s += (l.R[c + s2 * (tiny_wrap(k - 1, l.n2) - k)] * l.iD) * l.L2 + (l.R[c + s2 * (tiny_wrap(k + 1, l.n2) - k)] * l.iD) * l.L2;
      l.E[c] = l.R[c] - 1.f * s;
      l.X[c] = e;
    });
    __syncthreads();
    {  // the new residual is in E: swap the roles (Jacobi! writes r out of place), also in the table the up-stroke reads
      float* t = l.R;
      l.R = l.E;
      l.E = t;
      if (tid == 0) {
        lv[q].R = l.R;
        lv[q].E = l.E;
      }
    }
    const TinyLvl cl = lv[q + 1];
    cl.each_cell(tid, [&](int c, int i, int j, int k) {
      float s = 0.f;
      for (int kk = 2 * k; kk <= 2 * k + 1; kk++)
        for (int jj = 2 * j; jj <= 2 * j + 1; jj++)
          for (int ii = 2 * i; ii <= 2 * i + 1; ii++) s += l.R[ii + s1 * jj + s2 * kk];
      cl.R[c] = s;
    });
    __syncthreads();
  }
  // ---- coarsest level: smooth! from x = 0; then the up-stroke: prolongate! + increment!, smooth! ----
#pragma unroll 1
  for (int q = a.nlev - 1; q >= 0; q--) {
    const TinyLvl l = lv[q];
    const bool coarsest = q == a.nlev - 1;
    if (!coarsest) {
      const TinyLvl cl = lv[q + 1];
      const int c1 = cl.n0, c2 = cl.n0 * cl.n1;
      l.each_cell(tid, [&](int c, int i, int j, int k) {
        // ϵ[J] = x_c[down(J)] on the periodic image of J
        auto ec = [&](int ii, int jj, int kk) -> float {
          return cl.X[(tiny_wrap(ii, l.n0) >> 1) + c1 * (tiny_wrap(jj, l.n1) >> 1) + c2 * (tiny_wrap(kk, l.n2) >> 1)];
        };
        const float e = ec(i, j, k);
        float s = e * l.D;
        s += ec(i - 1, j, k) * l.L0 + ec(i + 1, j, k) * l.L0;
        s += ec(i, j - 1, k) * l.L1 + ec(i, j + 1, k) * l.L1;
        s += ec(i, j, k - 1) * l.L2 + ec(i, j, k + 1) * l.L2;
        l.R[c] = l.R[c] - w * s;
        l.X[c] = l.X[c] + w * e;
      });
      __syncthreads();
    }
    tiny_gs(l, w, coarsest ? 1 : 0, tid);
  }
  {  // x (and r) of level T back to global memory
    const TinyLvl l = lv[0];
    l.each_cell(tid, [&](int c, int i, int j, int k) {
      const i64 o = (i64)(a.g0.xo + i + 1) + a.g0.s[1] * (j + 1) + a.g0.s[2] * (k + 1);
      a.x0[o] = l.X[c];
      a.r0out[o] = l.R[c];
    });
  }
}

// ================================================================================================
// k_tiny_gen — the innermost levels of the V-cycle in GENERAL mode (bodies, walls, semi-coarsened levels) on ONE block in shared
// memory: the counterpart of k_tiny_uni for variable coefficients.  The levels' arrays (L, D, iD, x, ϵ, r, r2; ghost cells
// included) live in a shared-memory arena as dense copies; the operations are the SAME SmallOp list k_small_levels would run, with
// the general one-thread-per-cell bodies (b_k_jacobi, b_k_restrict, b_k_gs_init, b_k_gs_sweep, b_k_increment, b_k_prolong_inc —
// identical arithmetic, identical bits), whose Grid and pointers the host has redirected to the dense copies (pointers are arena
// offsets, encoded (offset+1)·4), separated by block barriers instead of grid barriers (≈ 0.3 µs instead of ≈ 4.4 µs per operation:
// the sphere wake's coarse end is 63 operations per V-cycle, 36 of them on levels of ≤ 1 800 cells).
// ================================================================================================
#define TINYG_MAXLEV 8
struct TinyGenLvl {
  Grid g;        // the level's layout in global memory
  const float* L;   // D components
  const float* Dg;
  const float* iD;
  int off;       // arena offset (floats) of the level's 9 dense arrays: L0 L1 L2 Dg iD x eps r r2
  int cells;     // N0·N1·N2
  float* x;      // the level's x and r in global memory: written back at the end (level T: the result; deeper levels: for observers)
  float* r;
  int r_out_off; // arena offset of the level's residual at exit (Jacobi! writes it out of place and the roles swap)
};
struct TinyGenArgs {
  int nlev;
  TinyGenLvl lv[TINYG_MAXLEV];
  const float* r0;  // residual of the first tiny level in global memory (input)
  float* x0;        // its solution (output)
  float* r0out;     // its residual after the smoother (output)
  int r_in_off;      // arena offset of level T's r at entry
  int arena_floats;  // the operation list follows the arena in shared memory (16-byte aligned)
};
__device__ __forceinline__ float* tinyg_fix(float* arena, const float* p) {
  return p ? arena + ((reinterpret_cast<uintptr_t>(p) >> 2) - 1) : nullptr;
}
__global__ void __launch_bounds__(1024, 1) k_tiny_gen(const SmallOp* __restrict__ ops, int nops, const float* wp, const __grid_constant__ TinyGenArgs a) {
  pdl_wait();
  extern __shared__ float tinyg_arena[];
  float* const arena = tinyg_arena;
  SmallOp* const sops = reinterpret_cast<SmallOp*>(arena + ((a.arena_floats + 3) & ~3));
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  for (int q = tid; q < a.arena_floats; q += 1024) arena[q] = 0.f;
  {  // the whole operation list into shared memory once (a descriptor fetched from global memory per operation was ≈ 1 µs each)
    const int* src = reinterpret_cast<const int*>(ops);
    int* dst = reinterpret_cast<int*>(sops);
    for (int q = tid; q < nops * (int)(sizeof(SmallOp) / sizeof(int)); q += 1024) dst[q] = src[q];
  }
  __syncthreads();
  // coefficients of every tiny level and r of the first one: dense copies, ghost cells included
  for (int l = 0; l < a.nlev; l++) {
    const TinyGenLvl& t = a.lv[l];
    const int N0 = t.g.N[0], N01 = t.g.N[0] * t.g.N[1];
    for (int c = tid; c < t.cells; c += 1024) {
      const int k = c / N01, rem = c - k * N01, j = rem / N0, i = rem - j * N0;
      const i64 o = (i64)(t.g.xo + i) + t.g.s[1] * j + t.g.s[2] * k;
      float* d = arena + t.off + c;
      d[0] = t.L[o];
      d[t.cells] = t.L[o + t.g.sc];
      d[2 * t.cells] = t.L[o + 2 * t.g.sc];
      d[3 * t.cells] = t.Dg[o];
      d[4 * t.cells] = t.iD[o];
      if (l == 0) arena[a.r_in_off + c] = a.r0[o];
    }
  }
  RedBuf nored{nullptr, nullptr, nullptr};
  for (int o = 0; o < nops; o++) {
    __syncthreads();
    const SmallOp& op = sops[o];
    // the block is 16 × 8 × 8 threads (the first tiny level of a wake domain is 16 × 8 × 8 cells: one pass, every lane busy);
    // threads beyond the operation's box in y or z have no cell in any virtual block and skip it (the kernel is issue-bound:
    // on the 4 × 2 × 2 and 2 × 2 × 2 levels 28 of the 32 warps have nothing to do)
    if ((int)threadIdx.y >= op.box.n[1] || (int)threadIdx.z >= op.box.n[2]) continue;
    Lvl lv = op.lvl;
    lv.L = tinyg_fix(arena, lv.L);
    lv.Dg = tinyg_fix(arena, lv.Dg);
    lv.iD = tinyg_fix(arena, lv.iD);
    lv.x = tinyg_fix(arena, lv.x);
    lv.eps = tinyg_fix(arena, lv.eps);
    lv.r = tinyg_fix(arena, lv.r);
    lv.r2 = tinyg_fix(arena, lv.r2);
    lv.z = nullptr;
    float* const other = tinyg_fix(arena, op.other);
    const int bx = (int)blockDim.x, by = (int)blockDim.y, bz = (int)blockDim.z;
    const int nx = op.type == OP_K_GSSWEEP ? (op.box.n[0] + 1) / 2 : op.box.n[0];  // (a half-sweep's threads own every other cell)
    const int vg0 = (nx + bx - 1) / bx, vg1 = (op.box.n[1] + by - 1) / by, vg2 = (op.box.n[2] + bz - 1) / bz;
    const int nvb = vg0 * vg1 * vg2;
    for (int v = 0; v < nvb; v++) {
      const int3 vb = make_int3(v % vg0, (v / vg0) % vg1, v / (vg0 * vg1));
      switch (op.type) {
        case OP_K_JACOBI:
          b_k_jacobi<3>(lv, op.box, op.x_is_zero, vb);
          break;
        case OP_K_RESTRICT:
          b_k_restrict<3>(op.gc, op.g, op.box, other, lv.r, op.cm[0], op.cm[1], op.cm[2], vb);
          break;
        case OP_K_GSINIT:
          b_k_gs_init<3>(lv, op.box, vb);
          break;
        case OP_K_GSSWEEP:
          b_k_gs_sweep<3>(lv, op.box, op.k0, vb);
          break;
        case OP_K_INC:
          b_k_increment<3>(lv, op.box, wp, op.x_is_zero, 0, nored, 0, vb);
          break;
        case OP_K_PROLONG:
          b_k_prolong_inc<3>(lv, op.gc, other, op.box, wp, op.cm[0], op.cm[1], op.cm[2], vb);
          break;
        default:
          break;
      }
    }
  }
  __syncthreads();
  // x and r of every tiny level back to global memory (interior cells): level T's are the result, the deeper ones keep the levels
  // observable (wl_download_level)
  for (int l = 0; l < a.nlev; l++) {
    const TinyGenLvl& t = a.lv[l];
    const int n0 = t.g.N[0] - 2, n1 = t.g.N[1] - 2, n2 = t.g.N[2] - 2;
    for (int c = tid; c < n0 * n1 * n2; c += 1024) {
      const int k = c / (n0 * n1), rem = c - k * n0 * n1, j = rem / n0, i = rem - j * n0;
      const int cc = (i + 1) + t.g.N[0] * ((j + 1) + t.g.N[1] * (k + 1));
      const i64 o = (i64)(t.g.xo + i + 1) + t.g.s[1] * (j + 1) + t.g.s[2] * (k + 1);
      t.x[o] = arena[t.off + 5 * t.cells + cc];
      t.r[o] = arena[t.r_out_off + cc];
    }
  }
}
