// wl_vsmooth.cuh — f_vsmooth: the up-stroke of one multigrid level in ONE pass over the level's fields (uniform mode).
//
// Fuses, for a level with a coarser level below it (src/MultiLevelPoisson.jl:99-100,106; src/Poisson.jl:100-104,141-148):
//     prolongate!(fine.ϵ, coarse.x); increment!(fine; ω)            r¹ = r − ω·A ϵc ,  x¹ = x + ω·ϵc
//     GaussSeidelRB!(fine; it=4, ω):  ϵ⁰ = r¹·iD ; four red/black half-sweeps ; increment!   r² = r¹ − ω·A ϵ⁴ ,  x² = x¹ + ω·ϵ⁴
//     (level 1 only) L₂ = Σ r²·r²
// which the march kernels run as six launches moving 80.5 B/cell (f_increment<PROLONG>, f_gs_a, 3 × f_gs_half, f_increment);
// here r and x are read once and written once: 16 B/cell (+ ⅛ cell of the coarse solution).
//
// Temporal blocking: a block owns an x-y tile, marches over a z chunk and keeps a ring of planes of ϵ and r¹ in shared memory.
// The six stages run as a software pipeline on different planes of the ring (newest first):
//     step t, phase α:  P (r¹, ϵ⁰) on plane t      sweep 2 on plane t−3      sweep 4 on plane t−6
//             phase β:  sweep 1 on plane t−1       sweep 3 on plane t−4      increment on plane t−7
// (two block barriers per plane; a stage reads planes q−1, q, q+1 of the previous stage).  r² depends on r within a distance
// of 5 cells, so a tile is loaded and processed with a halo of 5 rows in y and 5 planes in z — 8 cells in x, the granule of the
// layout below — and only its core is stored; the values computed in the halo get progressively wrong towards the tile edge,
// exactly one cell per stage, and never reach the core.  r² goes to a second array (neighbouring tiles still read r).
//
// Layout of a plane in the ring: cells are split by the parity of their x position into two arrays; a thread owns a group of
// 8 consecutive cells of a row = one float4 in each array.  A red/black half-sweep on row y moves the cells of ONE of the two
// arrays (which one alternates with y+z), reading the other array for the x neighbours and the same array one row / one plane
// away for the y / z neighbours, so every lane does useful work on full float4 vectors and the update is race-free in place.
// Warps are made of rows of equal parity, so the choice of the array is warp-uniform.
//
// The reference's periodic-ghost quirk (App. A.9-4: perBC! runs once before the sweeps, so across a periodic face a sweep sees
// the stale ϵ⁰ = r¹·iD of the wrapped cell; increment! refreshes the ghosts) is reproduced: a neighbour across the domain's
// periodic face is read from the r¹ ring and multiplied by iD during the sweeps, from the ϵ ring in the increment.
// Arithmetic and association order are those of mult_uni / Gs::gauss4 (wl_fast.cuh): results are bit-identical to the unfused path.
#pragma once
#include "wl_fast.cuh"

#define VS_NGX 8                       // groups of 8 cells per tile row: 6 core groups + one halo group on each side (8 float4 = one quarter-warp: row breaks cause no bank conflicts)
#define VS_CX ((VS_NGX - 2) * 8)       // core width  (48)
#ifndef VS_CY
#define VS_CY 42                       // core height
#endif
#ifndef VS_BPS
#define VS_BPS 1                       // blocks per SM the shared-memory footprint allows
#endif
#define VS_HALO 5
#define VS_TH (VS_CY + 2 * VS_HALO)    // tile rows (52)
#define VS_DE 9                        // ring depth of ϵ   (planes t … t−8)
#define VS_DR 8                        // ring depth of r¹  (planes t … t−7)
#define VS_AF (VS_TH * VS_NGX * 4)     // floats of one parity array of a plane
#define VS_PF (2 * VS_AF)              // floats of one plane (both arrays)
#define VS_NW (2 * ((VS_TH / 2 * VS_NGX + 31) / 32))  // warps: the rows of each parity are spread over NW/2 warps
#define VS_NT (32 * VS_NW)
#define VS_PAD 64                       // floats of padding before and after the rings (edge elements read neighbours out of the tile)
#define VS_SMEM (((VS_DE + VS_DR) * VS_PF + 2 * VS_PAD) * 4)

struct VsArgs {
  Grid g, gc;
  const float* xc;   // coarse solution
  const float* r;    // residual in
  float* r2;         // residual out
  float* x;          // solution, updated in place
  const float* wp;   // ω
  float L0, L1, L2, D, iD;
  int zchunk;
  int cy, th;  // core rows of a tile (≤ VS_CY: the level's rows split evenly over the tiles) and tile rows cy + 2·VS_HALO
  // z slab (multi-GPU): the level holds planes 1 … n2 of a globally periodic extent of n2g planes starting after global plane
  // zoff; planes 0 and n2+1 of r are the exchanged ghost planes, the four planes beyond them on each side are in rext
  // ([side][4] planes, nearest-to-farthest from the slab on the lower side reversed: index t+4 for t = −4…−1, t−n2−2 above).
  // The coarse solution is either a slab too (cslab: local planes 0 … n2/2+1 in xc, two more per side in cxext) or replicated
  // (whole periodic extent in xc, global indices).
  int slab, cslab, zoff, n2g;
  const float* rext;
  const float* cxext;
};

__device__ __forceinline__ float4 mul4s(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// Ring slot offsets (floats from the start of the ϵ ring) by (slot of plane t, planes back): a constant-bank lookup instead
// of modular arithmetic at every use.
struct VsLut {
  int e[VS_DE][VS_DE];
  int r[VS_DR][VS_DR];
};
constexpr VsLut vs_make_lut() {
  VsLut l{};
  for (int s = 0; s < VS_DE; s++)
    for (int k = 0; k < VS_DE; k++) l.e[s][k] = ((s - k + VS_DE) % VS_DE) * VS_PF;
  for (int s = 0; s < VS_DR; s++)
    for (int k = 0; k < VS_DR; k++) l.r[s][k] = VS_DE * VS_PF + ((s - k + VS_DR) % VS_DR) * VS_PF;
  return l;
}
__constant__ VsLut vs_lut = vs_make_lut();

// One red/black half-sweep of the thread's 4 target cells on plane q.  `sb` = ring base + the thread's element; eq/eqm/eqp and
// rq/rqm/rqp = slot offsets of planes q, q−1, q+1 in the ϵ and r¹ rings.  TE: the target is the array of even positions (its
// left x neighbours are (S, V.x, V.y, V.z), right V); otherwise odd positions (left V, right (V.y, V.z, V.w, S)), V = the other
// array.  The neighbour offsets are immediates: the rings are padded so that the tile's edge elements read garbage in bounds.
// stS/mS: the x neighbour S lies across the periodic face (read r¹, times mS = iD; mS = 1 otherwise); anyYZ: some y or z
// neighbour does (rare: one branch around the general form).
// Σ of the two face terms of one direction, ϵ₋·L + ϵ₊·L with the level's uniform face coefficient L.
// LM = 1: L == 1 (the finest level): the products are the operands themselves, ϵ₋ + ϵ₊ is the same Float32 value.
// LM = 2: L a power of two ≥ 1 (fully coarsened uniform levels, L = 2^level): scaling by 2^k commutes with rounding
//         (no overflow at these magnitudes; sums in the subnormal range are exact), so (ϵ₋ + ϵ₊)·L has the bits of ϵ₋·L + ϵ₊·L.
// LM = 0: the expression as the reference writes it (src/Poisson.jl:63-69,116-122).
template <int LM>
__device__ __forceinline__ float vs_pair(float a, float b, float L) {
  if (LM == 1) return a + b;
  if (LM == 2) return (a + b) * L;
  return a * L + b * L;
}

// oS: offset of the x neighbour beyond the group inside the other array (−1: last cell of the group to the left, for TE; +4: first
// cell of the group to the right); oYm / oYp: offsets of the rows below / above (∓ one row of groups).  The tile's edge threads
// (first / last group of a row, first / last row) have no such neighbour inside the tile: they pass 0 and re-read an element of their
// own — what they compute is halo garbage by design, but it must not come from an element another thread writes in the same
// phase (compute-sanitizer racecheck is clean with this; out-of-tile reads raced with the neighbouring rows' stores before).
template <bool TE, int LM>
__device__ __forceinline__ void vs_sweep(float* sb, int eq, int eqm, int eqp, int rq, int rqm, int rqp, bool stS, float mS, bool anyYZ, bool stYm,
                                         bool stYp, bool stZm, bool stZp, const VsArgs& a, int oS, int oYm, int oYp) {
  constexpr int tg = TE ? 0 : VS_AF, ot = TE ? VS_AF : 0;
  const float iD = a.iD;
  const float4 V = ld4(sb + eq + ot);
  const float S = sb[(stS ? rq : eq) + ot + oS] * mS;
  float4 ym, yp, zm, zp;
  if (!anyYZ) {
    ym = ld4(sb + eq + tg + oYm);
    yp = ld4(sb + eq + tg + oYp);
    zm = ld4(sb + eqm + tg);
    zp = ld4(sb + eqp + tg);
  } else {
    ym = stYm ? mul4s(ld4(sb + rq + tg + oYm), iD) : ld4(sb + eq + tg + oYm);
    yp = stYp ? mul4s(ld4(sb + rq + tg + oYp), iD) : ld4(sb + eq + tg + oYp);
    zm = stZm ? mul4s(ld4(sb + rqm + tg), iD) : ld4(sb + eqm + tg);
    zp = stZp ? mul4s(ld4(sb + rqp + tg), iD) : ld4(sb + eqp + tg);
  }
  float4 s = ld4(sb + rq + tg);
  const float L0 = a.L0, L1 = a.L1, L2 = a.L2;
  float4 lf, rt;
  if (TE) {
    lf = make_float4(S, V.x, V.y, V.z);
    rt = V;
  } else {
    lf = V;
    rt = make_float4(V.y, V.z, V.w, S);
  }
  s.x -= vs_pair<LM>(lf.x, rt.x, L0);
  s.x -= vs_pair<LM>(ym.x, yp.x, L1);
  s.x -= vs_pair<LM>(zm.x, zp.x, L2);
  s.y -= vs_pair<LM>(lf.y, rt.y, L0);
  s.y -= vs_pair<LM>(ym.y, yp.y, L1);
  s.y -= vs_pair<LM>(zm.y, zp.y, L2);
  s.z -= vs_pair<LM>(lf.z, rt.z, L0);
  s.z -= vs_pair<LM>(ym.z, yp.z, L1);
  s.z -= vs_pair<LM>(zm.z, zp.z, L2);
  s.w -= vs_pair<LM>(lf.w, rt.w, L0);
  s.w -= vs_pair<LM>(ym.w, yp.w, L1);
  s.w -= vs_pair<LM>(zm.w, zp.w, L2);
  st4(sb + eq + tg, mul4s(s, iD));
}

// A ϵ at the 4 cells of one parity array (mult_uni order: ϵ·D, + x pair, + y pair, + z pair)
template <int LM>
__device__ __forceinline__ float4 vs_mult(const float4& c, const float4& lf, const float4& rt, const float4& ym, const float4& yp, const float4& zm,
                                          const float4& zp, const VsArgs& a) {
  const float L0 = a.L0, L1 = a.L1, L2 = a.L2, D = a.D;
  float4 s;
  s.x = c.x * D;
  s.x += vs_pair<LM>(lf.x, rt.x, L0);
  s.x += vs_pair<LM>(ym.x, yp.x, L1);
  s.x += vs_pair<LM>(zm.x, zp.x, L2);
  s.y = c.y * D;
  s.y += vs_pair<LM>(lf.y, rt.y, L0);
  s.y += vs_pair<LM>(ym.y, yp.y, L1);
  s.y += vs_pair<LM>(zm.y, zp.y, L2);
  s.z = c.z * D;
  s.z += vs_pair<LM>(lf.z, rt.z, L0);
  s.z += vs_pair<LM>(ym.z, yp.z, L1);
  s.z += vs_pair<LM>(zm.z, zp.z, L2);
  s.w = c.w * D;
  s.w += vs_pair<LM>(lf.w, rt.w, L0);
  s.w += vs_pair<LM>(ym.w, yp.w, L1);
  s.w += vs_pair<LM>(zm.w, zp.w, L2);
  return s;
}

template <bool WITH_L2, bool SLAB, int LM>
__global__ void __launch_bounds__(VS_NT, VS_BPS) f_vsmooth(const __grid_constant__ VsArgs a, RedBuf R, int slot) {
  pdl_wait();
  extern __shared__ float4 vs_smem4[];
  // ϵ ring [VS_DE][2][VS_TH][VS_NGX] float4, then the r¹ ring [VS_DR][2][VS_TH][VS_NGX] float4
  float* const ER = reinterpret_cast<float*>(vs_smem4) + VS_PAD;
  const Grid& g = a.g;
  const Grid& gc = a.gc;
  const int n0 = g.N[0] - 2, n1 = g.N[1] - 2, n2 = g.N[2] - 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int par = warp & 1;  // parity class of this warp's rows
  const int qi = (warp >> 1) * 32 + lane;
  const int rc = qi / VS_NGX, gx = qi - rc * VS_NGX;
  const int ry = min(2 * rc + par, a.th - 1);
  const bool act = 2 * rc + par < a.th;
  const int xb = 1 + VS_CX * blockIdx.x, yb = 1 + a.cy * blockIdx.y;
  const int z0 = 1 + a.zchunk * blockIdx.z, z1 = min(z0 + a.zchunk, n2 + 1);
  const int xu = xb - 8 + 8 * gx, yu = yb - VS_HALO + ry;  // unwrapped coordinates of the group's first cell / of the row
  const int ypar = yu & 1;
  auto wrap = [](int v, int n) -> int {
    if (v < 1) v += n;
    else if (v > n) v -= n;
    return v;
  };
  const int xs = max(1, min(n0 - 7, wrap(xu, n0)));
  const int yr = max(1, min(n1, wrap(yu, n1)));
  // neighbours across the domain's periodic faces (see header): the group's left / right x neighbour, the rows below / above
  const bool stL = xu == 1 || xu == n0 + 1, stR = xu + 8 == 1 || xu + 8 == n0 + 1;
  const bool stYm = yu == 1 || yu == n1 + 1, stYp = yu == 0 || yu == n1;
  float* const sb = ER + (ry * VS_NGX + gx) * 4;  // the thread's element inside a parity array, relative to slot offsets
  const bool stY = stYm || stYp;
  const float mL = stL ? a.iD : 1.f, mR = stR ? a.iD : 1.f;
  constexpr int oY = VS_NGX * 4;
  // neighbour offsets of the sweeps; 0 where the neighbour would lie outside the tile (see vs_sweep)
  const int oSl = gx == 0 ? 0 : -1, oSr = gx == VS_NGX - 1 ? 0 : 4;
  const int oYm = ry == 0 ? 0 : -oY, oYp = ry == a.th - 1 ? 0 : oY;
  const bool core = act && gx >= 1 && gx <= VS_NGX - 2 && ry >= VS_HALO && ry < a.th - VS_HALO && xu <= n0 && yu <= n1;
  // global offsets (in-plane)
  const int gin = g.xo + xs + g.px * yr;
  const int cy = (yr + 1) >> 1;
  const int cyo = (wrap((yr & 1) ? yr - 1 : yr + 1, n1) + 1) >> 1;  // coarse row of the fine y neighbour that is not in the own coarse cell
  const int cx = (xs + 1) >> 1;
  const int cinx = gc.xo + cx;
  const int dcl = ((wrap(xs - 1, n0) + 1) >> 1) - cx, dcr = ((wrap(xs + 8, n0) + 1) >> 1) - cx;
  const float w = *a.wp;
  const float iD = a.iD;
  double l2 = 0.0;

  // global loads are issued one step ahead of their use (the block's warps run in lockstep between barriers, nothing else
  // would hide the DRAM latency): plane t+1 of r and the coarse solution around it for P, plane t−7 of x for the increment
  struct PIn {
    float4 f0, f1, C, Cy, Cz;
    float cl, cr;
  };
  // plane t of r (t may lie up to 5 planes outside the level) and the coarse plane under fine plane t
  auto rplane = [&](int t) -> const float* {
    if (!SLAB) return a.r + g.s[2] * wrap(t, n2);
    if (t < 0) return a.rext + g.s[2] * (t + 4);
    if (t > n2 + 1) return a.rext + g.s[2] * (4 + t - n2 - 2);
    return a.r + g.s[2] * t;
  };
  auto cplane = [&](int t) -> const float* {
    if (!SLAB) return a.xc + gc.s[2] * ((wrap(t, n2) + 1) >> 1);
    if (!a.cslab) return a.xc + gc.s[2] * ((wrap(t + a.zoff, a.n2g) + 1) >> 1);
    const int c = (t + 1) >> 1, nc = n2 >> 1;  // arithmetic shift = floor: planes below the slab map to coarse planes ≤ 0
    if (c < 0) return a.cxext + gc.s[2] * (c + 2);
    if (c > nc + 1) return a.cxext + gc.s[2] * (2 + c - nc - 2);
    return a.xc + gc.s[2] * c;
  };
  auto loadP = [&](int t) -> PIn {
    PIn p;
    const int tz = ((t + (SLAB ? a.zoff : 0)) & 1) ? t - 1 : t + 1;  // the fine z neighbour outside the own coarse cell (zoff and n2 are even)
    const float* rp = rplane(t) + gin;
    p.f0 = ld4(rp);
    p.f1 = ld4(rp + 4);
    const float* cp = cplane(t) + cinx;
    p.C = ld4(cp + gc.px * cy);
    p.cl = cp[gc.px * cy + dcl];
    p.cr = cp[gc.px * cy + dcr];
    p.Cy = ld4(cp + gc.px * cyo);
    p.Cz = ld4(cplane(tz) + cinx + gc.px * cy);
    return p;
  };
  PIn pin;
  if (act) pin = loadP(z0 - VS_HALO);
  float4 xi0 = f4zero(), xi1 = f4zero(), xiC = f4zero();  // x and the coarse solution for the increment of the NEXT step
  auto loadX = [&](int q) {
    if (core && q >= z0 && q <= z1 - 1) {
      const i64 o = g.s[2] * q + gin;
      xi0 = ld4(a.x + o);
      xi1 = ld4(a.x + o + 4);
      xiC = ld4(cplane(q) + cinx + gc.px * cy);
    }
  };

  int se = 0, sr = 0;  // ring slots of plane t
  for (int t = z0 - VS_HALO; t <= z1 + 6; t++) {
    const int* const le = vs_lut.e[se];  // slot offsets of planes t, t−1, …
    const int* const lr = vs_lut.r[sr];
    const bool doI = core && t - 7 >= z0 && t - 7 <= z1 - 1;
    // =============== phase α ===============
    if (act && t <= z1 + 4) {
      // ---- P: r¹ = r − ω·A ϵc with ϵc = coarse.x[down(·)], ϵ⁰ = r¹·iD, all 8 cells of the group on plane t ----
      const float4 f0 = pin.f0, f1 = pin.f1, C = pin.C, Cy = pin.Cy, Cz = pin.Cz;
      const float cl = pin.cl, cr = pin.cr;
      const float L0 = a.L0, L1 = a.L1, L2 = a.L2, D = a.D;
      // fine cell j of the group sits in coarse cell j/2; its x neighbour outside that coarse cell is on the left for even j
      auto Aec = [&](float c0, float xo, float yo, float zo) -> float {
        float s = c0 * D;
        s += vs_pair<LM>(c0, xo, L0);
        s += vs_pair<LM>(c0, yo, L1);
        s += vs_pair<LM>(c0, zo, L2);
        return s;
      };
      float4 re, ro;  // r¹ at even / odd positions
      re.x = f0.x - w * Aec(C.x, cl, Cy.x, Cz.x);
      ro.x = f0.y - w * Aec(C.x, C.y, Cy.x, Cz.x);
      re.y = f0.z - w * Aec(C.y, C.x, Cy.y, Cz.y);
      ro.y = f0.w - w * Aec(C.y, C.z, Cy.y, Cz.y);
      re.z = f1.x - w * Aec(C.z, C.y, Cy.z, Cz.z);
      ro.z = f1.y - w * Aec(C.z, C.w, Cy.z, Cz.z);
      re.w = f1.z - w * Aec(C.w, C.z, Cy.w, Cz.w);
      ro.w = f1.w - w * Aec(C.w, cr, Cy.w, Cz.w);
      st4(sb + lr[0], re);
      st4(sb + lr[0] + VS_AF, ro);
      st4(sb + le[0], mul4s(re, iD));
      st4(sb + le[0] + VS_AF, mul4s(ro, iD));
      // next plane's loads only now: a scoreboard counts all loads in flight from one instruction, so the loads of plane t+1
      // must not be issued before the values of plane t have been consumed (the fence keeps the compiler from hoisting them)
      asm volatile("" ::: "memory");
      if (t + 1 <= z1 + 4) pin = loadP(t + 1);
    }
    // ---- sweeps 2 and 4 move the even cells ((x+y+z) even; the array of even positions holds odd x) ----
#pragma unroll
    for (int k = 3; k <= 6; k += 3) {
      const int q = t - k;
      const int h = k == 3 ? 3 : 1;  // sweep 2 is valid on planes z0−3 … z1+2, sweep 4 on z0−1 … z1
      if (act && q >= z0 - h && q <= z1 - 1 + h) {
        const int zr = SLAB ? wrap(q + a.zoff, a.n2g) : wrap(q, n2);
        const bool zsm = zr == 1, zsp = zr == (SLAB ? a.n2g : n2);
        const bool te = ((1 + ypar + q) & 1) == 0;  // colour of the even-position array on this row and plane
        const bool anyYZ = stY || zsm || zsp;
        if (te)
          vs_sweep<true, LM>(sb, le[k], le[k + 1], le[k - 1], lr[k], lr[k + 1], lr[k - 1], stL, mL, anyYZ, stYm, stYp, zsm, zsp, a, oSl, oYm, oYp);
        else
          vs_sweep<false, LM>(sb, le[k], le[k + 1], le[k - 1], lr[k], lr[k + 1], lr[k - 1], stR, mR, anyYZ, stYm, stYp, zsm, zsp, a, oSr, oYm, oYp);
      }
    }
    __syncthreads();
    // =============== phase β ===============
    // ---- sweeps 1 and 3 move the odd cells ----
#pragma unroll
    for (int k = 1; k <= 4; k += 3) {
      const int q = t - k;
      const int h = k == 1 ? 4 : 2;  // sweep 1 is valid on planes z0−4 … z1+3, sweep 3 on z0−2 … z1+1
      if (act && q >= z0 - h && q <= z1 - 1 + h) {
        const int zr = SLAB ? wrap(q + a.zoff, a.n2g) : wrap(q, n2);
        const bool zsm = zr == 1, zsp = zr == (SLAB ? a.n2g : n2);
        const bool te = ((1 + ypar + q) & 1) == 1;
        const bool anyYZ = stY || zsm || zsp;
        if (te)
          vs_sweep<true, LM>(sb, le[k], le[k + 1], le[k - 1], lr[k], lr[k + 1], lr[k - 1], stL, mL, anyYZ, stYm, stYp, zsm, zsp, a, oSl, oYm, oYp);
        else
          vs_sweep<false, LM>(sb, le[k], le[k + 1], le[k - 1], lr[k], lr[k + 1], lr[k - 1], stR, mR, anyYZ, stYm, stYp, zsm, zsp, a, oSr, oYm, oYp);
      }
    }
    // ---- increment!: r² = r¹ − ω·A ϵ⁴ ; x² = (x + ω·ϵc) + ω·ϵ⁴ on plane t−7, core cells only ----
    {
      const int q = t - 7;
      if (doI) {
        const float* Eq = sb + le[7];
        const float* Eqm = sb + le[8];
        const float* Eqp = sb + le[6];
        const float* Rq = sb + lr[7];
        const float4 ce = ld4(Eq), co = ld4(Eq + VS_AF);
        const float sl = Eq[VS_AF - 1], sr_ = Eq[4];
        const float4 yme = ld4(Eq - oY), ymo = ld4(Eq + VS_AF - oY), zme = ld4(Eqm), zmo = ld4(Eqm + VS_AF);
        const float4 Ae = vs_mult<LM>(ce, make_float4(sl, co.x, co.y, co.z), co, yme, ld4(Eq + oY), zme, ld4(Eqp), a);
        const float4 Ao = vs_mult<LM>(co, ce, make_float4(ce.y, ce.z, ce.w, sr_), ymo, ld4(Eq + VS_AF + oY), zmo, ld4(Eqp + VS_AF), a);
        const float4 re = ld4(Rq), ro = ld4(Rq + VS_AF);
        const float4 ne = make_float4(re.x - w * Ae.x, re.y - w * Ae.y, re.z - w * Ae.z, re.w - w * Ae.w);
        const float4 no = make_float4(ro.x - w * Ao.x, ro.y - w * Ao.y, ro.z - w * Ao.z, ro.w - w * Ao.w);
        const i64 o = g.s[2] * q + gin;  // core rows and planes are inside the domain: no wrap
        st4(a.r2 + o, make_float4(ne.x, no.x, ne.y, no.y));
        st4(a.r2 + o + 4, make_float4(ne.z, no.z, ne.w, no.w));
        const float4 C = xiC;
        float4 x0 = xi0, x1 = xi1;
        x0.x = (x0.x + w * C.x) + w * ce.x;
        x0.y = (x0.y + w * C.x) + w * co.x;
        x0.z = (x0.z + w * C.y) + w * ce.y;
        x0.w = (x0.w + w * C.y) + w * co.y;
        x1.x = (x1.x + w * C.z) + w * ce.z;
        x1.y = (x1.y + w * C.z) + w * co.z;
        x1.z = (x1.z + w * C.w) + w * ce.w;
        x1.w = (x1.w + w * C.w) + w * co.w;
        st4(a.x + o, x0);
        st4(a.x + o + 4, x1);
        if (WITH_L2) {
          l2 += (double)ne.x * ne.x + (double)no.x * no.x + (double)ne.y * ne.y + (double)no.y * no.y;
          l2 += (double)ne.z * ne.z + (double)no.z * no.z + (double)ne.w * ne.w + (double)no.w * no.w;
        }
      }
    }
    asm volatile("" ::: "memory");
    loadX(t + 1 - 7);
    __syncthreads();
    se = se + 1 == VS_DE ? 0 : se + 1;
    sr = sr + 1 == VS_DR ? 0 : sr + 1;
  }
  if (WITH_L2) {
    double v[1] = {l2}, fin[1];
    grid_reduce<RED_SUM, 1>(v, R, slot, fin);
  }
}
