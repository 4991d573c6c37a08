// wl_conv4.cuh — fm_conv4: the momentum flux kernel of uniform mode (no body, no walls), four cells per thread.
//
// Same arithmetic, same association order and same results as fm_conv<LAM, FUSE=true, PER3=true> (wl_fast.cuh), i.e.
// conv_diff! in gather form with shared face fluxes (src/Flow.jl:38-62) fused with BDIM! for μ₀≡1, μ₁≡0, V≡0
// (src/Flow.jl:176-180) and scale_u! (src/Flow.jl:211-214).  fm_conv is issue-bound (713 instructions per cell, half of them
// shared-memory loads, integer and control overhead of a one-cell-per-thread tile); here a lane owns one aligned float4 of x,
// so a stencil row is one LDS.128 instead of four LDS.32 and all index arithmetic is amortised over four cells.
//
// Geometry: blockDim = (32, C4TY); a warp owns a 128-cell row segment, a block a 128 × C4TY tile, marching over a z chunk.
// Shared memory holds the three velocity components of planes z, z+1, z+2 (and the plane being fetched with cp.async) with
// an in-plane halo of 2, loaded through the periodic wrap so the ghost cells of u are never read.  Plane z-1 of a lane's own
// column lives in registers.  Face fluxes are evaluated once: the lower z flux is carried from the previous plane, the upper
// x flux of a lane's last cell comes from the next lane by shuffle, the lower y fluxes are computed one plane ahead and shared
// between the rows (warps) of a block through shared memory.
#pragma once
#include "wl_fast.cuh"

#define C4TY 8
#define C4RING 4
#define C4W 136                // row pitch: 4 + 128 + 4 floats, the x halo of 2 sits inside the pads
#define C4H (C4TY + 4)         // rows incl. the y halo of 2
#define C4CS (C4H * C4W)       // one component plane
#define C4PS (3 * C4CS)        // one ring slot
#define C4Q (C4W / 4)          // float4 per row
#define C4FILL ((C4H * C4Q + 32 * C4TY - 1) / (32 * C4TY))
#define C4SMEM ((C4RING * C4PS + 2 * 3 * (C4TY + 1) * 32 * 4) * 4)

// x/6.  Fast form (EXACT=false): the two FMAs of wl_kernels.cuh, proven equal to the IEEE division for 2^-100 ≤ |x| ≤ 3e38 by
// the exhaustive on-device test (wl_selftest_div6); x = ±0 gives 0.  The kernel never sees an input outside that range: the
// velocity field it reads was range-checked by the kernel that wrote it (f_correct_cfl, k_range_check, the halo push): every
// value is 0 or 2^-77 ≤ |v| ≤ 1e37.  Then 5c+2d−u — sums of multiples of the quantum 2^(−77−23) — is 0 or at least 2^-100 in
// magnitude, and below 8e37.  A field that fails the check (the 1e-34 far-field velocities of a flow starting from rest, denormals)
// is routed, whole, to the EXACT=true instance of the kernel, which divides; see fm_conv4 below.
template <bool EXACT>
__device__ __forceinline__ float div6_c4(float x) {
  if (EXACT) return x / 6.f;
  const float C = 0.16666667163372039794921875f;
  const float q0 = x * C;
  const float r = __fmaf_rn(-6.f, q0, x);
  return __fmaf_rn(r, C, q0);
}
template <bool EXACT>
__device__ __forceinline__ float2 div6_c4(const float2 x) {
  if (EXACT) return make_float2(x.x / 6.f, x.y / 6.f);
  const float2 C = splat2(0.16666667163372039794921875f);
  const float2 q0 = mul2(x, C);
  const float2 r = fma2(splat2(-6.f), q0, x);
  return fma2(r, C, q0);
}
__device__ __forceinline__ float sign_one(float v) { return __uint_as_float((__float_as_uint(v) & 0x80000000u) | 0x3f800000u); }

// ϕu(j, CI(I,i), u, û, λ) − ν ∂(j, CI(I,i), u) for an inner / periodic face (src/Flow.jl:8-11,52)
//
// quick(u,c,d) = median((5c+2d−u)/6, c, median(10c−9u, c, d))  (src/Flow.jl:6) is evaluated as a clamp: with a = (5c+2d−u)/6 and
// b = 10c−9u the nested median equals min(max(min(a,b), c), d) when c ≤ d and max(min(max(a,b), c), d) when d ≤ c (lattice
// identities on ordered values: same value, three min/max instead of eight; min(max(X,c),d) = max(min(X,d),c) for c ≤ d).  The second case is the first one mirrored, and
// mirroring (multiplying the three inputs by s = −1) is exact in IEEE arithmetic, so one code path serves both:
// λ = s·min(max(min(a',b'), s·c), s·d) with s = sign(d−c).  d−c is ±(u[I]−u[I−δ]) with the sign of û, so s = sign(t·û).
// (û = 0 gives conv = 0·λ = 0 whatever s is.)  min/max run on the half-rate ALU pipe, the multiplications on the FMA pipe.
// (±0: a clamp may return −0 where the nested median returns +0 and x/6 of −0 is +0 here; no later operation distinguishes
// the two zeros, and they compare equal.)
template <int LAM, bool EXACT>
__device__ __forceinline__ float flux_p(float uf, float um2, float um1, float u0c, float up1, float nu) {
  const float t = u0c - um1;
  const float diff = nu * t;
  const bool pos = uf > 0.f;
  const float u = pos ? um2 : up1, c = pos ? um1 : u0c, d = pos ? u0c : um1;
  float lam;
  if (LAM == 0) {
    const float s = sign_one(t * uf);
    const float cs = c * s, ds = d * s, us = u * s;
    const float a = div6_c4<EXACT>(5.f * cs + 2.f * ds - us);
    const float b = 10.f * cs - 9.f * us;
    lam = s * fmaxf(fminf(fminf(a, b), ds), cs);  // clamp(min(a,b), cs, ds) with cs ≤ ds, as max(min3(a,b,ds), cs): FMNMX3 + FMNMX
  } else if (LAM == 1) {
    lam = (c + d) / 2.f;  // cds
  } else {
    lam = (c <= fminf(u, d) || c >= fmaxf(u, d)) ? c : c + (d - c) * (c - u) / (d - u);  // vanLeer
  }
  return uf * lam - diff;
}
// The same for two cells at once: every Float32 operation of flux_p on both halves of a register pair (add2/mul2/fma2 are the
// scalar IEEE operations, issued once for two cells); selections, min/max and the sign stay scalar (no packed forms exist).
template <int LAM, bool EXACT>
__device__ __forceinline__ float2 flux_p2(const float2 uf, const float2 um2, const float2 um1, const float2 u0c, const float2 up1, const float2 nu2) {
  const float2 t = sub2(u0c, um1);
  const float2 diff = mul2(nu2, t);
  const bool px = uf.x > 0.f, py = uf.y > 0.f;
  const float2 u = make_float2(px ? um2.x : up1.x, py ? um2.y : up1.y);
  const float2 c = make_float2(px ? um1.x : u0c.x, py ? um1.y : u0c.y);
  const float2 d = make_float2(px ? u0c.x : um1.x, py ? u0c.y : um1.y);
  float2 lam;
  if (LAM == 0) {
    const float2 tu = mul2(t, uf);
    const float2 s = make_float2(sign_one(tu.x), sign_one(tu.y));
    const float2 cs = mul2(c, s), ds = mul2(d, s), us = mul2(u, s);
    const float2 a = div6_c4<EXACT>(sub2(add2x(mul2(splat2(5.f), cs), mul2(splat2(2.f), ds)), us));  // add2x / sub2x: see the contraction hazard
    const float2 b = sub2x(mul2(splat2(10.f), cs), mul2(splat2(9.f), us));                             // in wl_common.cuh
    // clamp(min(a,b), cs, ds) with cs ≤ ds written as max(min3(a, b, ds), cs): one three-input FMNMX3 and one FMNMX
    lam = mul2(s, make_float2(fmaxf(fminf(fminf(a.x, b.x), ds.x), cs.x), fmaxf(fminf(fminf(a.y, b.y), ds.y), cs.y)));
  } else if (LAM == 1) {
    lam = mul2(add2(c, d), splat2(0.5f));  // cds: (c+d)/2
  } else {
    lam.x = (c.x <= fminf(u.x, d.x) || c.x >= fmaxf(u.x, d.x)) ? c.x : c.x + (d.x - c.x) * (c.x - u.x) / (d.x - u.x);  // vanLeer
    lam.y = (c.y <= fminf(u.y, d.y) || c.y >= fmaxf(u.y, d.y)) ? c.y : c.y + (d.y - c.y) * (c.y - u.y) / (d.y - u.y);
  }
  return sub2x(mul2(uf, lam), diff);
}
template <int LAM, bool EXACT>
__device__ __forceinline__ float4 flux_p4(const float4& uf, const float4& um2, const float4& um1, const float4& u0c, const float4& up1, const float2 nu2) {
  return cat2(flux_p2<LAM, EXACT>(lo2(uf), lo2(um2), lo2(um1), lo2(u0c), lo2(up1), nu2),
              flux_p2<LAM, EXACT>(hi2(uf), hi2(um2), hi2(um1), hi2(u0c), hi2(up1), nu2));
}
// (a + b)/2 on four cells (û of a face, src/Flow.jl:3): /2 is the exact multiplication by 0.5
__device__ __forceinline__ float4 avg4(const float4& a, const float4& b) {
  const float2 h = splat2(0.5f);
  return cat2(mul2(add2(lo2(a), lo2(b)), h), mul2(add2(hi2(a), hi2(b)), h));
}
// û of the x component's flux: the x-shifted average ((a.x + left)/2, (a.y + a.x)/2, (a.z + a.y)/2, (a.w + a.z)/2)
__device__ __forceinline__ float4 avg4_shift(const float4& a, float left) {
  const float2 h = splat2(0.5f);
  return cat2(mul2(add2(lo2(a), make_float2(left, a.x)), h), mul2(add2(hi2(a), make_float2(a.y, a.z)), h));
}
__device__ __forceinline__ void acc_add(float4& r, const float4& f) {
  const float2 a = add2(lo2(r), lo2(f)), b = add2(hi2(r), hi2(f));
  r = cat2(a, b);
}
__device__ __forceinline__ void acc_sub(float4& r, const float4& f) {
  const float2 a = sub2(lo2(r), lo2(f)), b = sub2(hi2(r), hi2(f));
  r = cat2(a, b);
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}

// rflag[0]: the velocity field `ua` holds a value outside the fast division's input range (set by the kernel that wrote the field,
// cleared here).  The two instances of the kernel are launched back to back on the same grid: the one whose EXACT does not match the
// flag returns at once (≈3 µs), so the host never has to look at the flag.  rflag[1]: block ticket of the EXACT instance.
template <int LAM, bool EXACT>
__global__ void __launch_bounds__(32 * C4TY, 2) fm_conv4(const __grid_constant__ Grid g, const float* __restrict__ ua, const float* __restrict__ u0,
                                                         float* __restrict__ out, const float* __restrict__ dtp, float nu, int zchunk, int corrector, RedBuf R,
                                                         int slot, const float* __restrict__ uext, int* __restrict__ rflag, const Force fc) {
  pdl_wait();
  if ((*reinterpret_cast<volatile int*>(rflag) != 0) != EXACT) return;
  const int vbx = blockIdx.x, vby = blockIdx.y, vbz = blockIdx.z;
  extern __shared__ float4 smem4[];
  float* const T = reinterpret_cast<float*>(smem4);      // [C4RING][3][C4H][C4W]
  float* const Fy = T + C4RING * C4PS;                    // [2][3][C4TY+1][32] float4: lower y fluxes of planes z, z+1 (row C4TY: the block's upper edge)
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tid = lane + 32 * ty;
  const int xb = 1 + 128 * vbx, yb = 1 + C4TY * vby;
  const int x0 = xb + 4 * lane, y = yb + ty;
  const int z0 = 1 + zchunk * vbz, z1 = min(z0 + zchunk, g.N[2] - 1);
  const bool on = x0 <= g.N[0] - 2 && y <= g.N[1] - 2;
  const float dt = *dtp;
  const float2 nu2 = splat2(nu), dt2 = splat2(dt);

  // ---- tile fill: every plane is fetched with the same per-thread float4 elements ----
  int gof[C4FILL], sof[C4FILL];  // in-plane global offset / offset inside a component plane of the tile (-1: none)
#pragma unroll
  for (int k = 0; k < C4FILL; k++) {
    const int e = tid + k * 32 * C4TY;
    const int row = e / C4Q, q = e - row * C4Q;
    int xx = xb - 4 + 4 * q;  // first cell of the float4; interior extents are multiples of 4, so a float4 never straddles the wrap
    if (xx < 1) xx += g.N[0] - 2;
    else if (xx > g.N[0] - 2) xx -= g.N[0] - 2;
    xx = max(1, min(g.N[0] - 5, xx));
    int yy = yb - 2 + row;
    if (yy < 1) yy += g.N[1] - 2;
    else if (yy > g.N[1] - 2) yy -= g.N[1] - 2;
    yy = max(1, min(g.N[1] - 2, yy));
    sof[k] = e < C4H * C4Q ? row * C4W + 4 * q : -1;
    gof[k] = g.xo + xx + g.px * yy;
  }
  // global source of plane zz: periodic wrap, or (z slabs) the exchanged ghost plane / the second halo plane in uext
  auto plane_src = [&](int zz, const float*& src, i64& cs) {
    if (zz < 0 && g.zopen[0]) {
      src = uext;
      cs = g.s[2];
    } else if (zz > g.N[2] - 1 && g.zopen[1]) {
      src = uext + 3 * g.s[2];
      cs = g.s[2];
    } else {
      if (g.per[2]) {
        if (zz < 1) zz += g.N[2] - 2;
        else if (zz > g.N[2] - 2) zz -= g.N[2] - 2;
      }
      zz = max(0, min(g.N[2] - 1, zz));
      src = ua + g.s[2] * zz;
      cs = g.sc;
    }
  };
  auto fill = [&](int zz) {
    const float* src;
    i64 cs;
    plane_src(zz, src, cs);
    float* dst = T + ((zz + 1024) & (C4RING - 1)) * C4PS;
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
      for (int k = 0; k < C4FILL; k++)
        if (sof[k] >= 0) cp_async16(dst + c * C4CS + sof[k], src + c * cs + gof[k]);
    }
  };
  // own column of plane zz straight from global memory (prologue only)
  auto own_global = [&](int zz, int c) -> float4 {
    const float* src;
    i64 cs;
    plane_src(zz, src, cs);
    const int xx = min(x0, g.N[0] - 5), yy = min(y, g.N[1] - 2);
    return ld4(src + c * cs + g.xo + xx + (i64)g.px * yy);
  };

  const int col = (ty + 2) * C4W + 4 + 4 * lane;  // own float4 inside a component plane of the tile
  auto P = [&](int zz) -> const float* { return T + ((zz + 1024) & (C4RING - 1)) * C4PS + col; };
  auto fy4 = [&](int buf, int c, int row) -> float4* { return reinterpret_cast<float4*>(Fy) + ((buf * 3 + c) * (C4TY + 1) + row) * 32 + lane; };

  // prologue: planes z0-1, z0, z0+1 in the ring, plane z0-2 of the own column in registers; the loop starts one plane early
  // (z = z0-1, nothing stored) to produce the carried fluxes of plane z0
  fill(z0 - 1);
  fill(z0);
  fill(z0 + 1);
  float4 m1[3];
#pragma unroll
  for (int c = 0; c < 3; c++) m1[c] = own_global(z0 - 2, c);
  cp_async_wait_all();
  __syncthreads();

  float4 Fz[3] = {f4zero(), f4zero(), f4zero()};  // lower z fluxes of the current plane
  float gmax = 0.f;
  float e0prev = 0.f;  // u_x of the cell beyond the warp's segment (xb+128, y) on the previous plane: û of lane 2's extra z-momentum face
  const bool lastrow = ty == C4TY - 1;
  for (int z = z0 - 1; z < z1; z++) {
    const bool live = z >= z0;
    if (z + 1 < z1) fill(z + 3);  // into the slot plane z-1 left; read as plane z+2 of the next step
    float4 ub[3] = {f4zero(), f4zero(), f4zero()};  // corrector: u⁰ of the own cells, requested a whole plane of work ahead of its use
    if (corrector && live && on) {
      const i64 o = (i64)g.xo + x0 + g.s[1] * y + g.s[2] * z;
#pragma unroll
      for (int c = 0; c < 3; c++) ub[c] = ld4(u0 + o + c * g.sc);
    }
    const float* p0 = P(z);
    const float* p1 = P(z + 1);
    const float* p2 = P(z + 2);
    float4 r[3], own[3];
    float4 Fx2lo = f4zero();
    // ---- x fluxes on plane z (only needed for stored planes) ----
    if (live) {
      const float4 a0 = ld4(p0);                      // u_x on the own cells
      const float4 b1 = ld4(p0 - C4W);                // u_x one row down   (û of the y-momentum flux)
      float fex = 0.f;                                // the face beyond the warp's last cell: lanes 0-2 compute one component each
      if (lane < 3) {
        const float* e = T + ((z + 1024) & (C4RING - 1)) * C4PS + (ty + 2) * C4W + 4 + 128;  // cell xb+128 of u_x
        // (plane z-1 has left the ring: its value was kept in a register when it was plane z)
        const float other = lane == 0 ? e[-1] : (lane == 1 ? e[-C4W] : e0prev);
        const float* ei = e + lane * C4CS;
        fex = flux_p<LAM, EXACT>((e[0] + other) / 2.f, ei[-2], ei[-1], ei[0], ei[1], nu);
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 a = c == 0 ? a0 : ld4(p0 + c * C4CS);
        const float2 l2 = *reinterpret_cast<const float2*>(p0 + c * C4CS - 2);
        const float rr = p0[c * C4CS + 4];
        own[c] = a;
        float4 uf;
        if (c == 0) uf = avg4_shift(a0, l2.y);
        else if (c == 1) uf = avg4(a0, b1);
        else uf = avg4(a0, m1[0]);
        // the x stencil of the four cells out of six distinct pairs: (l2), (l2.y,a.x), (a.x,a.y), (a.y,a.z), (a.z,a.w), (a.w,rr)
        const float2 sA = make_float2(l2.y, a.x), sB = make_float2(a.y, a.z), sC = make_float2(a.w, rr);
        const float2 f01 = flux_p2<LAM, EXACT>(lo2(uf), l2, sA, lo2(a), sB, nu2);
        const float2 f23 = flux_p2<LAM, EXACT>(hi2(uf), lo2(a), sB, hi2(a), sC, nu2);
        const float4 lo = cat2(f01, f23);
        float hi = __shfl_down_sync(FULLMASK, lo.x, 1);
        const float e = __shfl_sync(FULLMASK, fex, c);
        if (lane == 31) hi = e;
        r[c] = make_float4(0.f + lo.x, 0.f + lo.y, 0.f + lo.z, 0.f + lo.w);
        r[c].x -= lo.y;
        r[c].y -= lo.z;
        r[c].z -= lo.w;
        r[c].w -= hi;
        if (c == 2) Fx2lo = lo;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; c++) own[c] = ld4(p0 + c * C4CS);
    }
    e0prev = T[((z + 1024) & (C4RING - 1)) * C4PS + (ty + 2) * C4W + 4 + 128];
    // ---- y fluxes: lower flux of the own row from the previous step, upper flux = next row's lower flux ----
    float4 Fy2lo = f4zero();
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 lo = *fy4(z & 1, c, ty);
        const float4 hi = *fy4(z & 1, c, lastrow ? C4TY : ty + 1);
        acc_add(r[c], lo);
        acc_sub(r[c], hi);
        if (c == 2) Fy2lo = lo;
      }
    }
    // ---- z fluxes: lower flux carried, upper flux = lower flux of plane z+1 ----
    {
      const float4 w1 = ld4(p1 + 2 * C4CS);           // u_z on plane z+1
      const float wl = p1[2 * C4CS - 1];              //   … and one cell to the left
      const float4 wd = ld4(p1 + 2 * C4CS - C4W);     //   … one row down
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 a1 = c == 2 ? w1 : ld4(p1 + c * C4CS);
        const float4 a2 = ld4(p2 + c * C4CS);
        float4 uf;
        if (c == 0) uf = avg4_shift(w1, wl);
        else if (c == 1) uf = avg4(w1, wd);
        else uf = avg4(w1, own[2]);
        const float4 hi = flux_p4<LAM, EXACT>(uf, m1[c], own[c], a1, a2, nu2);
        if (live) {
          acc_add(r[c], Fz[c]);
          acc_sub(r[c], hi);
        }
        if (live && c == 2) {
          // periodic images of the stale Φ the reference leaves on the upper ghost cells of σ (they enter maximum(σ) in CFL,
          // SURVEY App. A.9-1), see fm_conv: candidates by {k : I_k == 1}
          if (on) {
            const bool t1 = y == 1, t2 = (z + g.zoff) == 1;
            const float4 fz = Fz[2];
            if (t1) gmax = fmaxf(gmax, fmaxf(fmaxf(fz.x, fz.y), fmaxf(fz.z, fz.w)));
            else if (x0 == 1) gmax = fmaxf(gmax, fz.x);
            if (t2) gmax = fmaxf(gmax, fmaxf(fmaxf(Fy2lo.x, Fy2lo.y), fmaxf(Fy2lo.z, Fy2lo.w)));
            if (t2 && t1) gmax = fmaxf(gmax, fmaxf(fmaxf(Fx2lo.x, Fx2lo.y), fmaxf(Fx2lo.z, Fx2lo.w)));
          }
        }
        Fz[c] = hi;
      }
    }
    // ---- u_new = u⁰ + Δt·r  (predictor)  or  (u + u⁰ + Δt·r)/2  (corrector) ----
    if (live && on) {
      const i64 o = (i64)g.xo + x0 + g.s[1] * y + g.s[2] * z;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        if (fc.on) acc_add(r[c], splat4(fc.a[c]));      // accelerate! (src/Flow.jl:64-73)
        const float4 b = corrector ? ub[c] : own[c];  // predictor: ua is u⁰ itself, already in the tile
        float2 f01 = add2x(lo2(b), mul2(dt2, lo2(r[c]))), f23 = add2x(hi2(b), mul2(dt2, hi2(r[c])));
        if (corrector) {
          f01 = mul2(add2(lo2(own[c]), f01), splat2(0.5f));
          f23 = mul2(add2(hi2(own[c]), f23), splat2(0.5f));
        }
        st4(out + o + c * g.sc, cat2(f01, f23));
      }
    }
    // ---- lower y fluxes of plane z+1 for the next step (the block's last row also computes its upper flux) ----
    // The upper flux of the block's last row (row C4TY of Fy) is one more row of work: its three components go to the last three
    // warps, one each, so that no warp carries a whole extra row into the plane barrier.
    if (z + 1 < z1) {
      const int cex = ty - (C4TY - 3);  // component of the edge row this warp computes (< 0: none)
      const int nrow = cex >= 0 ? 2 : 1;
      for (int k = 0; k < nrow; k++) {
        const int dr = k * (C4TY - ty);  // rows above the own row
        const float* q1 = p1 + dr * C4W;
        const float* q0 = p0 + dr * C4W;
        const float4 v1 = ld4(q1 + C4CS);             // u_y on plane z+1, row y+k
        const float vl = q1[C4CS - 1];                //   … one cell to the left
        const float4 vd = ld4(q1 + C4CS - C4W);       //   … one row down
        const float4 vz = ld4(q0 + C4CS);             //   … one plane down (plane z)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (k == 1 && c != cex) continue;
          const float4 s2 = ld4(q1 + c * C4CS - 2 * C4W);
          const float4 s1 = c == 1 ? vd : ld4(q1 + c * C4CS - C4W);
          const float4 s0 = c == 1 ? v1 : ld4(q1 + c * C4CS);
          const float4 sp = ld4(q1 + c * C4CS + C4W);
          float4 uf;
          if (c == 0) uf = avg4_shift(v1, vl);
          else if (c == 1) uf = avg4(v1, vd);
          else uf = avg4(v1, vz);
          *fy4((z + 1) & 1, c, ty + dr) = flux_p4<LAM, EXACT>(uf, s2, s1, s0, sp, nu2);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) m1[c] = own[c];
    cp_async_wait_all();
    __syncthreads();
  }
  {
    double v[1] = {(double)gmax}, fin[1];
    grid_reduce<RED_MAX, 1>(v, R, slot, fin);
  }
  if (EXACT && tid == 0) {  // the last block of the EXACT instance re-arms the fast path for the next field
    __threadfence();
    const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(reinterpret_cast<unsigned*>(rflag + 1), 1u) == nb - 1) {
      rflag[1] = 0;
      rflag[0] = 0;
    }
  }
}
