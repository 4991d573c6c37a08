// wl_conv4g.cuh — fm_conv4g: the four-cells-per-thread flux kernel (wl_conv4.cuh) for GENERAL mode: walls, exit plane, bodies.
//
// Same arithmetic and results as fm_conv<LAM, FUSE=false, ·> (wl_fast.cuh): conv_diff! in gather form (src/Flow.jl:38-62) fused with
// BDIM-1, f = u⁰ + Δt·r − V (src/Flow.jl:178), on every cell with all indices ≥ 1 — the upper ghost rows included, where r holds
// the partial sums of the directions whose loops reach them — and the stale Φ the reference leaves on the upper
// ghost cells of σ (App. A.9-1).  Lower ghost planes of f: k_f_lowghost.  Differences from fm_conv4:
//   * faces on a wall take the one-sided forms of lowerBoundary!/upperBoundary! (ϕuL at index 1, ϕuR at index N−1, src/Flow.jl:47,56-62);
//   * the tile is filled by clamping at walls; periodic directions wrap, but a ghost ROW or PLANE is read as stored (with exitBC!
//     the periodic ghosts of the exit plane legitimately differ from their images, src/Flow.jl:194-195);
//   * the compute region is one cell larger in every direction (x, y, z up to N−1), with per-direction masks.
#pragma once
#include "wl_conv4.cuh"

// one face flux with the boundary variant: 0 inner / periodic (ϕu), 1 lower wall (ϕuL), 2 upper wall (ϕuR)
template <int LAM, int SAFE>
__device__ __forceinline__ float flux_v(int variant, float uf, float um2, float um1, float u0c, float up1, float nu, bool& tiny) {
  if (variant == 0) return flux_p<LAM, SAFE>(uf, um2, um1, u0c, up1, nu, tiny);
  const float diff = nu * (u0c - um1);
  // the limiter in the one-sided forms: λ(u,c,d) with (u,c,d) = (up1,u0c,um1) for ϕuL, (um2,um1,u0c) for ϕuR; evaluated through flux_p's
  // clamp form by handing it a face velocity of the matching sign (its result is û·λ − ν∂: recover λ·û by adding the diffusion back
  // would round twice, so the limiter is evaluated directly here)
  float lam;
  const bool upw = variant == 1 ? !(uf > 0.f) : !(uf < 0.f);  // the branch that uses the limiter
  if (!upw) {
    lam = (u0c + um1) / 2.f;
  } else {
    const float u = variant == 1 ? up1 : um2, c = variant == 1 ? u0c : um1, d = variant == 1 ? um1 : u0c;
    if (LAM == 0) {
      const float a = div6_chk<SAFE>(5.f * c + 2.f * d - u, tiny);
      const float b = 10.f * c - 9.f * u;
      lam = median3(a, c, median3(b, c, d));
    } else if (LAM == 1) {
      lam = (c + d) / 2.f;
    } else {
      lam = (c <= fminf(u, d) || c >= fmaxf(u, d)) ? c : c + (d - c) * (c - u) / (d - u);
    }
  }
  return uf * lam - diff;
}
template <int LAM, int SAFE>
__device__ __forceinline__ float4 flux_v4(int variant, const float4& uf, const float4& um2, const float4& um1, const float4& u0c, const float4& up1, float nu,
                                          bool& tiny) {
  if (variant == 0) return flux_p4<LAM, SAFE>(uf, um2, um1, u0c, up1, nu, tiny);
  return make_float4(flux_v<LAM, SAFE>(variant, uf.x, um2.x, um1.x, u0c.x, up1.x, nu, tiny), flux_v<LAM, SAFE>(variant, uf.y, um2.y, um1.y, u0c.y, up1.y, nu, tiny),
                     flux_v<LAM, SAFE>(variant, uf.z, um2.z, um1.z, u0c.z, up1.z, nu, tiny), flux_v<LAM, SAFE>(variant, uf.w, um2.w, um1.w, u0c.w, up1.w, nu, tiny));
}

template <int LAM>
__global__ void __launch_bounds__(32 * C4TY, 2) fm_conv4g(const __grid_constant__ Grid g, const float* __restrict__ ua, const float* __restrict__ u0,
                                                          const float* __restrict__ V, float* __restrict__ out, float* __restrict__ sigma,
                                                          const float* __restrict__ dtp, float nu, int zchunk, const float* __restrict__ uext,
                                                          int* __restrict__ flag) {
  constexpr int SAFE = 2;  // x/6 exact in place (see div6_chk)
  const int vbx = blockIdx.x, vby = blockIdx.y, vbz = blockIdx.z;
  extern __shared__ float4 smem4[];
  float* const T = reinterpret_cast<float*>(smem4);
  float* const Fy = T + C4RING * C4PS;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int tid = lane + 32 * ty;
  const int N0 = g.N[0], N1 = g.N[1], N2 = g.N[2];
  const int xb = 1 + 128 * vbx, yb = 1 + C4TY * vby;
  const int x0 = xb + 4 * lane, y = yb + ty;
  const int z0 = 1 + zchunk * vbz, z1 = min(z0 + zchunk, N2);  // planes z0 … z1-1 ≤ N2-1
  const bool rowon = y <= N1 - 1;
  const float dt = *dtp;
  bool bad = false, tiny = false;  // (tiny is not raised in this mode)
  // walls: a face of direction j at index 1 / N−1 is one-sided unless the direction is periodic (or, in z, open to a neighbouring slab)
  const bool wx = !g.per[0], wy = !g.per[1], wz0 = !g.per[2] && !g.zopen[0], wz1 = !g.per[2] && !g.zopen[1];
  auto varx = [&](int x) -> int { return wx ? (x == 1 ? 1 : (x == N0 - 1 ? 2 : 0)) : 0; };
  auto vary = [&](int yy) -> int { return wy ? (yy == 1 ? 1 : (yy == N1 - 1 ? 2 : 0)) : 0; };
  auto varz = [&](int zz) -> int { return (zz == 1 && wz0) ? 1 : ((zz == N2 - 1 && wz1) ? 2 : 0); };

  // ---- tile fill ----
  int gof[C4FILL], sof[C4FILL];
#pragma unroll
  for (int k = 0; k < C4FILL; k++) {
    const int e = tid + k * 32 * C4TY;
    const int row = e / C4Q, q = e - row * C4Q;
    int xx = xb - 4 + 4 * q;  // float4 granule (x ≡ 1 mod 4): periodic x wraps whole granules (ghost column = its image after BC!)
    if (g.per[0]) {
      if (xx < 1) xx += N0 - 2;
      else if (xx > N0 - 2) xx -= N0 - 2;
    }
    xx = max(-3, min(N0 - 1, xx));  // the granule −3…0 holds the ghost column 0 (the row's leading pad is readable)
    int yy = yb - 2 + row;
    if (g.per[1]) {  // a ghost row is read as stored; beyond it: the periodic image
      if (yy < 0) yy += N1 - 2;
      else if (yy > N1 - 1) yy -= N1 - 2;
    }
    yy = max(0, min(N1 - 1, yy));
    sof[k] = e < C4H * C4Q ? row * C4W + 4 * q : -1;
    gof[k] = g.xo + xx + g.px * yy;
  }
  auto plane_src = [&](int zz, const float*& src, i64& cs) {
    if (zz < 0 && g.zopen[0]) {
      src = uext;
      cs = g.s[2];
    } else if (zz > N2 - 1 && g.zopen[1]) {
      src = uext + 3 * g.s[2];
      cs = g.s[2];
    } else {
      if (g.per[2]) {
        if (zz < 0) zz += N2 - 2;
        else if (zz > N2 - 1) zz -= N2 - 2;
      }
      zz = max(0, min(N2 - 1, zz));
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
  auto own_global = [&](int zz, int c) -> float4 {
    const float* src;
    i64 cs;
    plane_src(zz, src, cs);
    const int xx = min(x0, N0 - 1), yy = min(y, N1 - 1);
    return ld4(src + c * cs + g.xo + xx + (i64)g.px * yy);
  };
  const int col = (ty + 2) * C4W + 4 + 4 * lane;
  auto P = [&](int zz) -> const float* { return T + ((zz + 1024) & (C4RING - 1)) * C4PS + col; };
  auto fy4 = [&](int buf, int c, int row) -> float4* { return reinterpret_cast<float4*>(Fy) + ((buf * 3 + c) * (C4TY + 1) + row) * 32 + lane; };

  fill(z0 - 1);
  fill(z0);
  fill(z0 + 1);
  float4 m1[3];
#pragma unroll
  for (int c = 0; c < 3; c++) m1[c] = own_global(z0 - 2, c);
  cp_async_wait_all();
  __syncthreads();

  // per-cell masks of the lane's four cells (constant over z)
  const bool cv0 = x0 <= N0 - 1, cv1 = x0 + 1 <= N0 - 1, cv2 = x0 + 2 <= N0 - 1, cv3 = x0 + 3 <= N0 - 1;          // inside the compute region
  const bool ax0 = x0 <= N0 - 2, ax1 = x0 + 1 <= N0 - 2, ax2 = x0 + 2 <= N0 - 2, ax3 = x0 + 3 <= N0 - 2;          // the x loops reach the cell
  const bool ay = y <= N1 - 2;
  const int vx0 = varx(x0);            // only the first cell of a group can sit on a wall face (x = 1, x = N0−1 ≡ 1 mod 4)
  const int lo0 = g.per[0] ? 1 : 2, lo1 = g.per[1] ? 1 : 2, lo2 = (g.per[2] || g.zopen[0]) ? 1 : 2;

  float4 Fz[3] = {f4zero(), f4zero(), f4zero()};
  const bool lastrow = ty == C4TY - 1;
  for (int z = z0 - 1; z < z1; z++) {
    const bool live = z >= z0;
    const bool az = z <= N2 - 2;
    if (z + 1 < z1) fill(z + 3);
    const float* p0 = P(z);
    const float* p1 = P(z + 1);
    const float* p2 = P(z + 2);
    float4 r[3], own[3];
    float4 Fx2lo = f4zero();
    // ---- x fluxes on plane z ----
    if (live) {
      const float4 a0 = ld4(p0);
      const float4 b1 = ld4(p0 - C4W);
      float fex = 0.f;
      if (lane < 3) {  // the face beyond the warp's last cell (x = xb+128), one component per lane
        const float* e = T + ((z + 1024) & (C4RING - 1)) * C4PS + (ty + 2) * C4W + 4 + 128;
        float other;
        if (lane == 0) other = e[-1];
        else if (lane == 1) other = e[-C4W];
        else {
          const float* src;
          i64 cs;
          plane_src(z - 1, src, cs);
          int xx = xb + 128;
          if (g.per[0] && xx > N0 - 2) xx -= N0 - 2;
          other = src[g.xo + min(xx, N0 - 1) + (i64)g.px * min(y, N1 - 1)];
        }
        const float* ei = e + lane * C4CS;
        fex = flux_v<LAM, SAFE>(varx(xb + 128), (e[0] + other) / 2.f, ei[-2], ei[-1], ei[0], ei[1], nu, tiny);
      }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 a = c == 0 ? a0 : ld4(p0 + c * C4CS);
        const float2 l2 = *reinterpret_cast<const float2*>(p0 + c * C4CS - 2);
        const float rr = p0[c * C4CS + 4];
        own[c] = a;
        float4 uf;
        if (c == 0) uf = make_float4((a0.x + l2.y) / 2.f, (a0.y + a0.x) / 2.f, (a0.z + a0.y) / 2.f, (a0.w + a0.z) / 2.f);
        else if (c == 1) uf = avg4(a0, b1);
        else uf = avg4(a0, m1[0]);
        float4 lo = flux_p4<LAM, SAFE>(uf, make_float4(l2.x, l2.y, a.x, a.y), make_float4(l2.y, a.x, a.y, a.z), a, make_float4(a.y, a.z, a.w, rr), nu, tiny);
        if (vx0) lo.x = flux_v<LAM, SAFE>(vx0, uf.x, l2.x, l2.y, a.x, a.y, nu, tiny);  // wall face of the group's first cell
        float hi = __shfl_down_sync(FULLMASK, lo.x, 1);
        const float e = __shfl_sync(FULLMASK, fex, c);
        if (lane == 31) hi = e;
        // r = Σ_j [F_j(I) − F_j(I+δ_j)] in the reference's order; the x terms only where the x loops reach the cell
        r[c].x = ax0 ? (0.f + lo.x) - lo.y : 0.f;
        r[c].y = ax1 ? (0.f + lo.y) - lo.z : 0.f;
        r[c].z = ax2 ? (0.f + lo.z) - lo.w : 0.f;
        r[c].w = ax3 ? (0.f + lo.w) - hi : 0.f;
        if (c == 2) Fx2lo = lo;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; c++) own[c] = ld4(p0 + c * C4CS);
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
      bad = bad || (cv0 && !(fabsf(own[c].x) <= 1e37f)) || (cv1 && !(fabsf(own[c].y) <= 1e37f)) || (cv2 && !(fabsf(own[c].z) <= 1e37f)) ||
            (cv3 && !(fabsf(own[c].w) <= 1e37f));
    // ---- y fluxes ----
    float4 Fy2lo = f4zero();
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 lo = *fy4(z & 1, c, ty);
        const float4 hi = *fy4(z & 1, c, lastrow ? C4TY : ty + 1);
        if (ay) {
          r[c].x += lo.x;
          r[c].y += lo.y;
          r[c].z += lo.z;
          r[c].w += lo.w;
          r[c].x -= hi.x;
          r[c].y -= hi.y;
          r[c].z -= hi.z;
          r[c].w -= hi.w;
        }
        if (c == 2) Fy2lo = lo;
      }
    }
    // ---- z fluxes ----
    float4 Fz2 = Fz[2];
    {
      const int vz = varz(z + 1);
      const float4 w1 = ld4(p1 + 2 * C4CS);
      const float wl = p1[2 * C4CS - 1];
      const float4 wd = ld4(p1 + 2 * C4CS - C4W);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float4 a1 = c == 2 ? w1 : ld4(p1 + c * C4CS);
        const float4 a2 = ld4(p2 + c * C4CS);
        float4 uf;
        if (c == 0) uf = make_float4((w1.x + wl) / 2.f, (w1.y + w1.x) / 2.f, (w1.z + w1.y) / 2.f, (w1.w + w1.z) / 2.f);
        else if (c == 1) uf = avg4(w1, wd);
        else uf = avg4(w1, own[2]);
        const float4 hi = flux_v4<LAM, SAFE>(vz, uf, m1[c], own[c], a1, a2, nu, tiny);
        if (live && az) {
          r[c].x += Fz[c].x;
          r[c].y += Fz[c].y;
          r[c].z += Fz[c].z;
          r[c].w += Fz[c].w;
          r[c].x -= hi.x;
          r[c].y -= hi.y;
          r[c].z -= hi.z;
          r[c].w -= hi.w;
        }
        Fz[c] = hi;
      }
    }
    // ---- f = u⁰ + Δt·r − V on the cells of the compute region; stale Φ on the upper ghost cells of σ ----
    if (live && rowon && cv0) {
      const i64 o = (i64)g.xo + x0 + g.s[1] * y + g.s[2] * z;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const i64 oc = o + c * g.sc;
        if (cv3) {
          const float4 b = ld4(u0 + oc), v = ld4(V + oc);
          st4(out + oc, make_float4(b.x + dt * r[c].x - v.x, b.y + dt * r[c].y - v.y, b.z + dt * r[c].z - v.z, b.w + dt * r[c].w - v.w));
        } else {  // the ghost column: only the group's first cell exists
          out[oc] = u0[oc] + dt * r[c].x - V[oc];
        }
      }
      const bool gyz = y == N1 - 1 || (z == N2 - 1 && !g.zopen[1]);
      if (gyz || x0 + 3 >= N0 - 1) {
        const float fz[4] = {Fz2.x, Fz2.y, Fz2.z, Fz2.w}, fy[4] = {Fy2lo.x, Fy2lo.y, Fy2lo.z, Fy2lo.w}, fx[4] = {Fx2lo.x, Fx2lo.y, Fx2lo.z, Fx2lo.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int x = x0 + j;
          if (x > N0 - 1 || !(gyz || x == N0 - 1)) continue;
          const bool axj = x <= N0 - 2;
          // the last Φ written by the reference's (i=D, j) loops: the largest j whose range contains the cell
          if (az && z >= lo2) sigma[o + j] = fz[j];
          else if (ay && y >= lo1) sigma[o + j] = fy[j];
          else if (axj && x >= lo0) sigma[o + j] = fx[j];
        }
      }
    }
    // ---- lower y fluxes of plane z+1 for the next step ----
    if (z + 1 < z1) {
      const int nrow = lastrow ? 2 : 1;
      for (int k = 0; k < nrow; k++) {
        const int vy = vary(y + k);
        const float* q1 = p1 + k * C4W;
        const float* q0 = p0 + k * C4W;
        const float4 v1 = ld4(q1 + C4CS);
        const float vl = q1[C4CS - 1];
        const float4 vd = ld4(q1 + C4CS - C4W);
        const float4 vz = ld4(q0 + C4CS);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float4 s2 = ld4(q1 + c * C4CS - 2 * C4W);
          const float4 s1 = c == 1 ? vd : ld4(q1 + c * C4CS - C4W);
          const float4 s0 = c == 1 ? v1 : ld4(q1 + c * C4CS);
          const float4 sp = ld4(q1 + c * C4CS + C4W);
          float4 uf;
          if (c == 0) uf = make_float4((v1.x + vl) / 2.f, (v1.y + v1.x) / 2.f, (v1.z + v1.y) / 2.f, (v1.w + v1.z) / 2.f);
          else if (c == 1) uf = avg4(v1, vd);
          else uf = avg4(v1, vz);
          *fy4((z + 1) & 1, c, ty + k) = flux_v4<LAM, SAFE>(vy, uf, s2, s1, s0, sp, nu, tiny);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) m1[c] = own[c];
    cp_async_wait_all();
    __syncthreads();
  }
  if (bad) *flag = 1;
}
