// wl_kernels.cuh — sm_100a kernels for WaterLily.jl's mom_step! path (general, variable-coefficient forms).
// Each kernel cites the reference @loop(s) it fuses (paths relative to the WaterLily.jl tree).
// Indices are 0-based here: Julia index I ↔ I-1; interior = 1 .. N-2; ghosts 0 and N-1.
#pragma once
#include "wl_common.cuh"

// ------------------------------------------------------------------------------------
// limiters  (src/Flow.jl:4-6, :27-36)
// ------------------------------------------------------------------------------------
// median(a,b,c) of src/Flow.jl:27-36 returns the middle value; max(min(a,b), min(max(a,b),c)) is the same value
// (ties return an equal value) without branches.
__device__ __forceinline__ float median3(float a, float b, float c) { return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c)); }
// x/6 correctly rounded without the division sequence: q0 = x·RN(1/6); the exact remainder r = x − 6·q0 (one FMA) corrects it,
// q = q0 + r·RN(1/6) (one FMA).  Bit-identical to IEEE x/6.f for every finite x with |x| ≥ 2^-100 — verified exhaustively over
// all 2^32 inputs on the device by wl_selftest_div6 (tests/test_gpu_parity.py); tiny and non-finite inputs take the real division.
// The rare path (tiny, zero, non-finite inputs).  Through double: x/6.0 is correctly rounded in 53 bits, and rounding that to Float32
// equals the Float32 division (double rounding is innocuous for a quotient when the wider format has ≥ 2·24+2 bits; ties at .5 of a
// subnormal ulp are exact in double) — without the Float32 division's denormal subroutine, which a wake's far field would hit on every
// step.  Covered by the same exhaustive test as the fast form.
__device__ __noinline__ float div6_slow(float x) { return (float)((double)x / 6.0); }
__device__ __forceinline__ float div6(float x) {
  const float C = 0.16666667163372039794921875f;  // RN(1/6)
  const float q0 = x * C;
  const float r = __fmaf_rn(-6.f, q0, x);
  float q = __fmaf_rn(r, C, q0);
  const float ax = fabsf(x);
  if (!(ax >= 7.888609052210118e-31f && ax <= 3.0e38f)) q = (ax == 0.f) ? q0 : div6_slow(x);  // ±0 keeps its sign; the rest is never physical
  return q;
}
// Variant for the flux kernels: the same two-FMA form; the rare inputs outside its proven range (tiny non-zero values such as the
// 1e-34 far-field velocities of the first steps of a wake, ±0, non-finite) take the IEEE division, so every input is exact.
// (`flag` is kept in the signature of the callers; it is no longer raised here.)
__device__ __forceinline__ float div6_flag(float x, int* flag) {
  const float C = 0.16666667163372039794921875f;
  const float q0 = x * C;
  const float r = __fmaf_rn(-6.f, q0, x);
  float q = __fmaf_rn(r, C, q0);
  const float ax = fabsf(x);
  const bool inr = ax >= 7.888609052210118e-31f && ax <= 3.0e38f;
  if (!inr) q = (ax == 0.f) ? q0 : div6_slow(x);
  (void)flag;
  return q;
}
template <int LAM>
__device__ __forceinline__ float limiter_f(float u, float c, float d, int* flag) {
  if (LAM == 0) return median3(div6_flag(5.f * c + 2.f * d - u, flag), c, median3(10.f * c - 9.f * u, c, d));  // quick
  if (LAM == 1) return (c + d) / 2.f;                                                                          // cds
  return (c <= fminf(u, d) || c >= fmaxf(u, d)) ? c : c + (d - c) * (c - u) / (d - u);                        // vanLeer
}
template <int LAM>
__device__ __forceinline__ float limiter(float u, float c, float d) {
  if (LAM == 0) return median3(div6(5.f * c + 2.f * d - u), c, median3(10.f * c - 9.f * u, c, d));  // quick
  if (LAM == 1) return (c + d) / 2.f;                                                                  // cds
  return (c <= fminf(u, d) || c >= fmaxf(u, d)) ? c : c + (d - c) * (c - u) / (d - u);                // vanLeer
}

// Flux through the LOWER j-face of cell I for momentum component i:
//   F = ϕu*(j,CI(I,i),u,ϕ(i,CI(I,j),u),λ) − ν ∂(j,CI(I,i),u)
// with the boundary variants of lowerBoundary!/upperBoundary! (src/Flow.jl:47,56-62):
//   I_j == 1      : ϕuL (one-sided) or, periodic, ϕuP with the upwind point wrapped to N_j-3
//   I_j == N_j-1  : ϕuR, or, periodic, the flux stored at the wrapped cell I_j → 1
//   otherwise     : ϕu
template <int D, int LAM>
__device__ __forceinline__ float face_flux(const Grid& g, const float* __restrict__ ui, const float* __restrict__ uj, int i, int j, int Ij, i64 o,
                                           float nu) {
  const i64 sj = g.s[j], si = g.s[i];
  const int Nj = g.N[j];
  if (g.per[j] && Ij == Nj - 1) {
    o -= (i64)(Nj - 2) * sj;
    Ij = 1;
  }
  const float uf = (uj[o] + uj[o - si]) / 2.f;  // ϕ(i,CI(I,j),u)
  const float c = ui[o - sj], d = ui[o];
  const float diff = nu * (d - c);
  float conv;
  if (Ij == 1) {
    if (g.per[j])
      conv = uf > 0.f ? uf * limiter<LAM>(ui[o + (i64)(Nj - 4) * sj], c, d) : uf * limiter<LAM>(ui[o + sj], d, c);
    else
      conv = uf > 0.f ? uf * ((d + c) / 2.f) : uf * limiter<LAM>(ui[o + sj], d, c);
  } else if (Ij == Nj - 1) {
    conv = uf < 0.f ? uf * ((d + c) / 2.f) : uf * limiter<LAM>(ui[o - 2 * sj], c, d);
  } else {
    conv = uf > 0.f ? uf * limiter<LAM>(ui[o - 2 * sj], c, d) : uf * limiter<LAM>(ui[o + sj], d, c);
  }
  return conv - diff;
}

// conv_diff!(f,u,Φ,λ) in gather form (src/Flow.jl:38-62) fused with BDIM-1 (src/Flow.jl:178):
//   r[I,i] = Σ_j [ F_ij(I) − F_ij(I+δ_j) ]  accumulated in the reference's order (+lo, −hi for j=1,2,3)
//   f[I,i] = u⁰[I,i] + Δt·r[I,i] − V[I,i]            over ALL cells (ghost rows included, App. A.9-2)
// and the stale-Φ bookkeeping of σ on upper-ghost cells that later enters maximum(σ) in CFL (App. A.9-1).
// mode 0: write raw r (wl_conv_diff);  mode 1: write BDIM-1 f.
template <int D, int LAM>
__global__ void __launch_bounds__(512) k_conv_bdim1(Grid g, Box box, const float* __restrict__ ua, const float* __restrict__ u0,
                                                    const float* __restrict__ V, float* __restrict__ f, float* __restrict__ sigma,
                                                    const float* __restrict__ dtp, float nu, int mode, const Force fc) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  bool lowok = true;  // all I_k >= 1
#pragma unroll
  for (int d = 0; d < D; d++) lowok = lowok && (I[d] >= 1);
  const float dt = *dtp;
  float Fd_last = 0.f;  // F_{D-1,j}(I) for the stale-Φ record
  int jphi = -1;
#pragma unroll
  for (int i = 0; i < D; i++) {
    const float* ui = ua + (i64)i * g.sc;
    float r = 0.f;
    if (lowok) {
#pragma unroll
      for (int j = 0; j < D; j++) {
        if (I[j] <= g.N[j] - 2) {
          const float* uj = ua + (i64)j * g.sc;
          const float Flo = face_flux<D, LAM>(g, ui, uj, i, j, I[j], o, nu);
          const float Fhi = face_flux<D, LAM>(g, ui, uj, i, j, I[j] + 1, o + g.s[j], nu);
          r += Flo;
          r -= Fhi;
          if (i == D - 1 && I[j] >= (g.per[j] ? 1 : 2)) {
            Fd_last = Flo;
            jphi = j;
          }
        }
      }
    }
    const i64 oc = o + (i64)i * g.sc;
    if (mode && fc.on) r += fc.a[i];  // accelerate!
    f[oc] = mode ? (u0[oc] + dt * r - V[oc]) : r;
  }
  bool ghost = false;
#pragma unroll
  for (int d = 0; d < D; d++) ghost = ghost || (I[d] == g.N[d] - 1);
  if (ghost && jphi >= 0) sigma[o] = Fd_last;
}

// BDIM-2 (src/Flow.jl:179) fused with scale_u! (src/Flow.jl:211-214):
//   X = μddn(I,μ₁,f) + V + μ₀ f ;  predictor: u = X (u was scaled by 0) ; corrector: u = (u + X)·0.5
// `nobody` (may be null): per block, 1 if the block holds no body — μ₁ ≡ 0, V ≡ 0, μ₀ ≡ 1 (0 on the wall faces BC!(μ₀,0) zeroes),
// k_nobody_flags.  There X = (0/2 + 0) + μ₀·f = 0 + μ₀·f exactly, and only f is read (24–36 B per cell instead of 90).
template <int D>
__device__ __forceinline__ bool wall_face_lo(const Grid& g, const int I[3], int i) {
  return !g.per[i] && I[i] == 1 && !(D == 3 && i == 2 && g.zopen[0]);
}
template <int D>
__global__ void __launch_bounds__(512) k_bdim2(Grid g, Box box, float* __restrict__ u, const float* __restrict__ f, const float* __restrict__ V,
                                               const float* __restrict__ mu0, const float* __restrict__ mu1, int corrector,
                                               const unsigned char* __restrict__ nobody) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  if (nobody && nobody[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)]) {
#pragma unroll
    for (int i = 0; i < D; i++) {
      const i64 oc = o + (i64)i * g.sc;
      const float X = 0.f + (wall_face_lo<D>(g, I, i) ? 0.f : 1.f) * f[oc];
      u[oc] = corrector ? (u[oc] + X) * 0.5f : X;
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < D; i++) {
    const float* fi = f + (i64)i * g.sc;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < D; j++) s += mu1[o + g.sc * (i + D * j)] * (fi[o + g.s[j]] - fi[o - g.s[j]]);
    const i64 oc = o + (i64)i * g.sc;
    const float X = s / 2.f + V[oc] + mu0[oc] * fi[o];
    u[oc] = corrector ? (u[oc] + X) * 0.5f : X;
  }
}

// ---- built-in udf: sgs! with the Smagorinsky–Lilly eddy viscosity (src/util.jl:46-76) --------------------------------------
// ∂uᵢ/∂xⱼ at the centre of cell I (src/Metrics.jl:42-44); o = offset of I.
__device__ __forceinline__ float sgs_dudx(const Grid& g, const float* __restrict__ u, i64 o, int i, int j) {
  const float* ui = u + (i64)i * g.sc;
  if (i == j) return ui[o + g.s[i]] - ui[o];
  const i64 p = o + g.s[j], m = o - g.s[j];
  return (ui[p] + ui[p + g.s[i]] - ui[m] - ui[m + g.s[i]]) / 4.f;
}
// νₜ[I] = (Cs·Δ)²·sqrt(dot(S[I,:,:],S[I,:,:])) with S(I,u) = (∂ᵢuⱼ+∂ⱼuᵢ)/2 (src/util.jl:62,68; src/Metrics.jl:140) over inside(σ).
// The reference keeps the tensor in the user's buffer S and evaluates νₜ from it at every use; the value is the same, so only νₜ is
// stored.  Ghost cells of νₜ stay 0: the reference's S is written on inside(σ) only and the flux loops below read it on upper ghosts.
template <int D>
__global__ void __launch_bounds__(512) k_sgs_nut(Grid g, Box box, const float* __restrict__ u, float* __restrict__ nut, float c2) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < D; j++)
#pragma unroll
    for (int i = 0; i < D; i++) {  // column-major, like the generic dot of the two views
      const float Sij = (sgs_dudx(g, u, o, i, j) + sgs_dudx(g, u, o, j, i)) / 2.f;
      s += Sij * Sij;
    }
  nut[o] = c2 * sqrtf(s);
}
// The flux loops of sgs! in gather form, then accelerate! and BDIM-1 (src/util.jl:69-75; src/Flow.jl:64-73,178): for every cell K and
// component i, in the reference's order  j = 1…D:  r += σᵢⱼ(K) if K ∈ inside_u(N,j);  r −= σᵢⱼ(K+δⱼ) if K+δⱼ ∈ inside_u(N,j),
// σᵢⱼ(I) = −νₜ(I)·(u[I,i] − u[I−δⱼ,i]);  f = u⁰ + Δt·r − V over ALL cells (r: conv_diff!'s raw sum, k_conv_bdim1 mode 0).
// flow.σ keeps the last value the (i=D, j) loops wrote on each upper-ghost cell (it enters maximum(σ) in CFL, App. A.9-1).
template <int D>
__global__ void __launch_bounds__(512) k_sgs_apply(Grid g, Box box, const float* __restrict__ ua, const float* __restrict__ nut,
                                                   const float* __restrict__ u0, const float* __restrict__ V, float* __restrict__ f,
                                                   float* __restrict__ sigma, const float* __restrict__ dtp, const Force fc) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  const float dt = *dtp;
  auto in_u = [&](const int* J, int j) -> bool {  // inside_u(N,j), 0-based: 2 … N_j−2 in j, 1 … N_k−1 elsewhere
    bool ok = true;
#pragma unroll
    for (int d = 0; d < D; d++) ok = ok && (d == j ? (J[d] >= 2 && J[d] <= g.N[d] - 2) : (J[d] >= 1 && J[d] <= g.N[d] - 1));
    return ok;
  };
  float last = 0.f;
  bool have = false;
#pragma unroll
  for (int i = 0; i < D; i++) {
    const float* ui = ua + (i64)i * g.sc;
    const i64 oc = o + (i64)i * g.sc;
    float r = f[oc];
#pragma unroll
    for (int j = 0; j < D; j++) {
      if (in_u(I, j)) {
        const float s = -nut[o] * (ui[o] - ui[o - g.s[j]]);
        r += s;
        if (i == D - 1) {
          last = s;
          have = true;
        }
      }
      int J[3] = {I[0], I[1], I[2]};
      J[j]++;
      if (in_u(J, j)) {
        const i64 oj = o + g.s[j];
        r -= -nut[oj] * (ui[oj] - ui[o]);
      }
    }
    if (fc.on) r += fc.a[i];  // accelerate!
    f[oc] = u0[oc] + dt * r - V[oc];
  }
  bool ghost = false;
#pragma unroll
  for (int d = 0; d < D; d++) ghost = ghost || (I[d] == g.N[d] - 1);
  if (ghost && have) sigma[o] = last;
}

// Flags for k_bdim2's body-free fast path, same launch geometry as k_bdim2.
template <int D>
__global__ void __launch_bounds__(512) k_nobody_flags(Grid g, Box box, const float* __restrict__ V, const float* __restrict__ mu0,
                                                      const float* __restrict__ mu1, unsigned char* __restrict__ flags) {
  pdl_wait();
  int I[3];
  int ok = 1;
  if (thread_cell<D>(box, I)) {
    const i64 o = cell_off(g, I);
#pragma unroll
    for (int i = 0; i < D; i++) {
      if (V[o + g.sc * i] != 0.f) ok = 0;
      if (mu0[o + g.sc * i] != (wall_face_lo<D>(g, I, i) ? 0.f : 1.f)) ok = 0;
#pragma unroll
      for (int j = 0; j < D; j++)
        if (mu1[o + g.sc * (i + D * j)] != 0.f) ok = 0;
    }
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) flags[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)] = (unsigned char)ok;
}

// BC!(a,U,saveexit,perdir) for a constant U (src/core.jl:200-219) in ONE launch.  The reference fills planes
// sequentially for i, for j; the final value of any ghost cell is found by resolving dimensions from the last to
// the first: periodic → opposite interior, normal → U, tangential → inward neighbour.
// blockIdx.z selects the plane: (j, lower ghost | upper ghost | first interior).
template <int D>
__global__ void k_bc_vec(Grid g, float* a, const float* keep, float U0, float U1, float U2, int saveexit) {
  pdl_wait();
  const int plane = blockIdx.z;
  const int j = plane / 3, which = plane % 3;
  if (which == 2 && g.per[j]) return;
  if (D == 3 && j == 2) {  // z faces owned by a neighbouring rank are filled by the halo exchange, not here
    if ((which == 0 || which == 2) && g.zopen[0]) return;
    if (which == 1 && g.zopen[1]) return;
  }
  // plane coordinates: the two (or one) dims other than j
  int I[3] = {0, 0, 0};
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int t1 = blockIdx.y * blockDim.y + threadIdx.y;
  int da = (j == 0) ? 1 : 0, db = (j == 2) ? 1 : 2;
  if (D == 2) {
    if (t1 > 0) return;
    if (t0 >= g.N[da]) return;
    I[da] = t0;
  } else {
    if (t0 >= g.N[da] || t1 >= g.N[db]) return;
    I[da] = t0;
    I[db] = t1;
  }
  I[j] = which == 0 ? 0 : (which == 1 ? g.N[j] - 1 : 1);
  const i64 o = cell_off(g, I);
  const float U[3] = {U0, U1, U2};
#pragma unroll
  for (int i = 0; i < D; i++) {
    if (which == 2 && i != j) continue;
    int S[3] = {I[0], I[1], I[2]};
    bool isconst = false;
#pragma unroll
    for (int d = D - 1; d >= 0; d--) {
      if (D == 3 && d == 2 && ((S[2] == 0 && g.zopen[0]) || (S[2] == g.N[2] - 1 && g.zopen[1]) || (S[2] == 1 && g.zopen[0]))) {
        continue;  // slab-internal z plane: behaves like an interior plane here (its ghost cells arrive by exchange)
      }
      if (g.per[d]) {
        if (S[d] == 0) S[d] = g.N[d] - 2;
        else if (S[d] == g.N[d] - 1) S[d] = 1;
      } else if (d == i) {
        if (S[d] <= 1 || (S[d] == g.N[d] - 1 && !(saveexit && i == 0))) {
          isconst = true;
          break;
        }
      } else {
        if (S[d] == 0) S[d] = 1;
        else if (S[d] == g.N[d] - 1) S[d] = g.N[d] - 2;
      }
    }
    float* ai = a + (i64)i * g.sc;
    if (isconst) {
      ai[o] = U[i];
    } else {
      // a source left on the exit plane (saveexit) holds the value from before this step: read it from `keep`
      const i64 so = cell_off(g, S);
      const bool kept = saveexit && i == 0 && !g.per[0] && S[0] == g.N[0] - 1;
      if (so != o) ai[o] = kept ? keep[so] : ai[so];
    }
  }
}

// perBC!(a,perdir) for a scalar (src/core.jl:239-243) in one launch, same plane scheme.
template <int D>
__global__ void k_perbc(Grid g, float* __restrict__ a) {
  pdl_wait();
  const int plane = blockIdx.z;
  const int j = plane / 2, which = plane % 2;
  if (!g.per[j]) return;
  int I[3] = {0, 0, 0};
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int t1 = blockIdx.y * blockDim.y + threadIdx.y;
  int da = (j == 0) ? 1 : 0, db = (j == 2) ? 1 : 2;
  if (D == 2) {
    if (t1 > 0 || t0 >= g.N[da]) return;
    I[da] = t0;
  } else {
    if (t0 >= g.N[da] || t1 >= g.N[db]) return;
    I[da] = t0;
    I[db] = t1;
  }
  I[j] = which == 0 ? 0 : g.N[j] - 1;
  int S[3] = {I[0], I[1], I[2]};
#pragma unroll
  for (int d = 0; d < D; d++)
    if (g.per[d]) {
      if (S[d] == 0) S[d] = g.N[d] - 2;
      else if (S[d] == g.N[d] - 1) S[d] = 1;
    }
  a[cell_off(g, I)] = a[cell_off(g, S)];
}

// exitBC!(u,u⁰,Δt) (src/core.jl:226-233) as three plane kernels with deterministic in-kernel means.
//  stage 0: partial sums of u[1,·,·,0]                     → out[slot]   (inflow mass flux U·len)
//  stage 1: u[N-1] = u⁰[N-1] − U·Δt·(u⁰[N-1] − u⁰[N-2]); partial sums of the new plane → out[slot+1]
//  stage 2: u[N-1] −= (mean − U)
template <int D>
__global__ void __launch_bounds__(256) k_exitbc(Grid g, float* __restrict__ u, const float* __restrict__ u0, const float* __restrict__ dtp,
                                                float dt_scale, RedBuf R, int slot, int stage, float len) {
  pdl_wait();
  int J[3] = {0, 0, 0};
  J[1] = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  J[2] = (D == 3) ? 1 + blockIdx.y * blockDim.y + threadIdx.y : 0;
  const bool ok = J[1] <= g.N[1] - 2 && (D == 2 || J[2] <= g.N[2] - 2);
  double v[1] = {0.0}, fin[1];
  if (stage == 0) {
    J[0] = 1;
    if (ok) v[0] = (double)u[cell_off(g, J)];
    grid_reduce<RED_SUM, 1>(v, R, slot, fin);
  } else if (stage == 1) {
    const float Um = (float)R.out[slot] / len;
    const float dt = (*dtp) * dt_scale;
    J[0] = g.N[0] - 1;
    if (ok) {
      const i64 o = cell_off(g, J);
      const float nv = u0[o] - Um * dt * (u0[o] - u0[o - 1]);
      u[o] = nv;
      v[0] = (double)nv;
    }
    grid_reduce<RED_SUM, 1>(v, R, slot + 1, fin);
  } else {
    const float Um = (float)R.out[slot] / len;
    const float flux = (float)R.out[slot + 1] / len - Um;
    J[0] = g.N[0] - 1;
    if (ok) u[cell_off(g, J)] -= flux;
  }
}

// ------------------------------------------------------------------------------------
// Poisson operator pieces (src/Poisson.jl)
// ------------------------------------------------------------------------------------
struct Lvl {
  Grid g;
  const float* L;  // D components (level 1: flow.μ₀)
  float* Dg;
  float* iD;
  float* x;
  float* eps;
  float* r;
  float* r2;  // ping-pong partner of r for Jacobi
  float* z;
};

// mult(I,L,D,x) (src/Poisson.jl:70-76)
template <int D>
__device__ __forceinline__ float mult_at(const Grid& g, const float* __restrict__ x, const float* __restrict__ L, const float* __restrict__ Dg, i64 o,
                                         const i64 lo[3], const i64 hi[3]) {
  float s = x[o] * Dg[o];
#pragma unroll
  for (int d = 0; d < D; d++) s += x[o + lo[d]] * L[o + g.sc * d] + x[o + hi[d]] * L[o + g.sc * d + g.s[d]];
  return s;
}

// set_diag!(D,iD,L) (src/Poisson.jl:43-55)
template <int D>
__global__ void k_set_diag(Grid g, Box box, const float* __restrict__ L, float* __restrict__ Dg, float* __restrict__ iD) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < D; d++) s -= L[o + g.sc * d] + L[o + g.sc * d + g.s[d]];
  Dg[o] = s;
  iD[o] = (s == 0.f) ? s : 1.f / s;
}

// restrictL!(a,b,c) interior part (src/MultiLevelPoisson.jl:42-46, :20-26, upL :9-11); BC!(a,0) follows via k_bc_vec.
template <int D>
__global__ void k_restrictL(Grid gc, Grid gf, Box box, float* __restrict__ a, const float* __restrict__ b, int c0, int c1, int c2, int zoffc) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const int c[3] = {c0, c1, c2};
  const i64 o = cell_off(gc, I);
  if (D == 3) I[2] -= zoffc;  // z slab restricting into a replicated level: local fine planes ↔ global coarse planes
#pragma unroll
  for (int i = 0; i < D; i++) {
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < D; j++) {
      // Julia: fine 2I-2 : 2I-1  (1-based)  ↔  0-based 2I0-1 : 2I0
      if (j == i) {
        lo[j] = hi[j] = c[i] ? 2 * I[j] - 1 : I[j];
      } else {
        lo[j] = c[j] ? 2 * I[j] - 1 : I[j];
        hi[j] = c[j] ? 2 * I[j] : I[j];
      }
    }
    float s = 0.f;
    for (int k = lo[2]; k <= hi[2]; k++)
      for (int jj = lo[1]; jj <= hi[1]; jj++)
        for (int ii = lo[0]; ii <= hi[0]; ii++) s += b[(i64)(gf.xo + ii) + gf.s[1] * jj + gf.s[2] * k + gf.sc * i];
    a[o + gc.sc * i] = c[i] ? s / 2.f : s;
  }
}

// @inside z = div(I,u) (src/Flow.jl:225, :13-19) fused with x .*= dt and residual!(p) part 1 (src/Poisson.jl:93-95):
//   z = Σ_i (u[I+δ_i,i] − u[I,i]);  x = p·dt;  r = iD==0 ? 0 : z − A x;  Σr → out[slot]
// with_div=0 skips the divergence and the scaling (standalone residual! on x, z).
template <int D>
__global__ void __launch_bounds__(512) k_div_residual(Lvl l, Box box, const float* __restrict__ u, const float* __restrict__ p,
                                                      const float* __restrict__ dtp, float w, int with_div, RedBuf R, int slot) {
  pdl_wait();
  const Grid& g = l.g;
  int I[3];
  const bool ok = thread_cell<D>(box, I);
  double v[1] = {0.0}, fin[1];
  if (ok) {
    const i64 o = cell_off(g, I);
    i64 lo[3], hi[3];
    nbr_offsets<D>(g, I, lo, hi);
    float z, xs, Ax;
    if (with_div) {
      const float dt = w * (*dtp);
      z = 0.f;
#pragma unroll
      for (int d = 0; d < D; d++) z += u[o + g.sc * d + g.s[d]] - u[o + g.sc * d];
      l.z[o] = z;
      xs = p[o] * dt;
      Ax = xs * l.Dg[o];
#pragma unroll
      for (int d = 0; d < D; d++) Ax += (p[o + lo[d]] * dt) * l.L[o + g.sc * d] + (p[o + hi[d]] * dt) * l.L[o + g.sc * d + g.s[d]];
      l.x[o] = xs;
    } else {
      z = l.z[o];
      Ax = mult_at<D>(g, l.x, l.L, l.Dg, o, lo, hi);
    }
    const float r = (l.iD[o] == 0.f) ? 0.f : z - Ax;
    l.r[o] = r;
    v[0] = (double)r;
  }
  grid_reduce<RED_SUM, 1>(v, R, slot, fin);
}

// residual! part 2 (src/Poisson.jl:95-97) fused with L₂(p) (src/Poisson.jl:189):
//   s = Σr/|inside|;  |s| > 2eps ? r -= s ;  Σ r² → out[slot_out]
template <int D>
__global__ void __launch_bounds__(512) k_resid_fix(Lvl l, Box box, float count, RedBuf R, int slot_in, int slot_out) {
  pdl_wait();
  const Grid& g = l.g;
  int I[3];
  const bool ok = thread_cell<D>(box, I);
  const float s = (float)R.out[slot_in] / count;
  double v[1] = {0.0}, fin[1];
  if (ok) {
    const i64 o = cell_off(g, I);
    float r = l.r[o];
    if (fabsf(s) > 2.f * 1.1920929e-7f) {
      r = r - s;
      l.r[o] = r;
    }
    v[0] = (double)r * (double)r;
  }
  grid_reduce<RED_SUM, 1>(v, R, slot_out, fin);
}

// Jacobi!(p) with ω=1 (src/Poisson.jl:111-114 + increment! :100-104) without materialising ϵ:
//   ϵ = r·iD (neighbours recomputed on the fly, periodic wrap = perBC!(ϵ));  r' = r − Aϵ;  x (+)= ϵ
// r' goes to the ping-pong buffer r2 because neighbours still need the old r.
template <int D>
__device__ __forceinline__ void b_k_jacobi(Lvl l, Box box, int x_is_zero, const int3 vb) {
  const Grid& g = l.g;
  int I[3];
  if (!thread_cell<D>(box, I, vb)) return;
  const i64 o = cell_off(g, I);
  i64 lo[3], hi[3];
  nbr_offsets<D>(g, I, lo, hi);
  const float e = l.r[o] * l.iD[o];
  float s = e * l.Dg[o];
#pragma unroll
  for (int d = 0; d < D; d++)
    s += (l.r[o + lo[d]] * l.iD[o + lo[d]]) * l.L[o + g.sc * d] + (l.r[o + hi[d]] * l.iD[o + hi[d]]) * l.L[o + g.sc * d + g.s[d]];
  l.r2[o] = l.r[o] - 1.f * s;
  l.x[o] = x_is_zero ? e : l.x[o] + 1.f * e;
}
template <int D>
__global__ void __launch_bounds__(512) k_jacobi(Lvl l, Box box, int x_is_zero) {
  pdl_wait();
  b_k_jacobi<D>(l, box, x_is_zero, real_block());
}

// restrict!(a,b,c) (src/MultiLevelPoisson.jl:49, :13-19): coarse r = Σ fine r over up(I,c), x fastest.
template <int D>
__device__ __forceinline__ void b_k_restrict(Grid gc, Grid gf, Box box, float* a, const float* b, int c0, int c1, int c2, const int3 vb) {
  int I[3];
  if (!thread_cell<D>(box, I, vb)) return;
  const int c[3] = {c0, c1, c2};
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
  for (int j = 0; j < D; j++) {
    lo[j] = c[j] ? 2 * I[j] - 1 : I[j];
    hi[j] = c[j] ? 2 * I[j] : I[j];
  }
  float s = 0.f;
  for (int k = lo[2]; k <= hi[2]; k++)
    for (int jj = lo[1]; jj <= hi[1]; jj++)
      for (int ii = lo[0]; ii <= hi[0]; ii++) s += b[(i64)(gf.xo + ii) + gf.s[1] * jj + gf.s[2] * k];
  a[cell_off(gc, I)] = s;
}
template <int D>
__global__ void k_restrict(Grid gc, Grid gf, Box box, float* __restrict__ a, const float* __restrict__ b, int c0, int c1, int c2) {
  pdl_wait();
  b_k_restrict<D>(gc, gf, box, a, b, c0, c1, c2, real_block());
}

// GaussSeidelRB! line 1: @inside ϵ = r·iD (src/Poisson.jl:142).  perBC!(ϵ) is NOT materialised: the sweeps
// below read r·iD of the wrapped cell across periodic faces, which is exactly the stale ghost the reference sees.
template <int D>
__device__ __forceinline__ void b_k_gs_init(Lvl l, Box box, const int3 vb) {
  int I[3];
  if (!thread_cell<D>(box, I, vb)) return;
  const i64 o = cell_off(l.g, I);
  l.eps[o] = l.r[o] * l.iD[o];
}
template <int D>
__global__ void k_gs_init(Lvl l, Box box) {
  pdl_wait();
  b_k_gs_init<D>(l, box, real_block());
}

// One red/black half-sweep gauss_rb(ϵ,r,L,iD,k₀,·) (src/Poisson.jl:116-132,145).  Thread t along x handles the cell
// x = 1 + 2t + shift of the sweep's colour: Σ(1-based idx) ≡ 1+k₀ (mod 2).  The reference's half_rangek only reaches
// last-dim indices k with Iv=(k+1+p)/2 ≤ N_d÷2 (matters for odd N_d).
template <int D>
__device__ __forceinline__ void b_k_gs_sweep(Lvl l, Box box, int k0, const int3 vb) {
  const Grid& g = l.g;
  int I[3];
  I[1] = box.lo[1] + vb.y * blockDim.y + threadIdx.y;
  I[2] = (D == 3) ? box.lo[2] + vb.z * blockDim.z + threadIdx.z : 0;
  if (I[1] >= box.lo[1] + box.n[1]) return;
  if (D == 3 && I[2] >= box.lo[2] + box.n[2]) return;
  const int t = vb.x * blockDim.x + threadIdx.x;
  // 1-based index sum parity must equal (1+k0)&1 :  (Σ0 + D) & 1 == (1+k0) & 1
  const int rest = I[1] + I[2] + D + 1 + k0;  // x0 must make (x0 + rest) even
  I[0] = 1 + 2 * t + ((1 + rest) & 1);
  if (I[0] > g.N[0] - 2) return;
  {  // half_rangek reach in the last dimension
    const int d = D - 1;
    const int k1 = I[d] + 1;  // 1-based
    int front = 0;
#pragma unroll
    for (int q = 0; q < D - 1; q++) front += I[q] + 1;
    const int pp = (front + k0) & 1;
    const int Iv = (k1 + 1 + pp) / 2;
    if (Iv > g.N[d] / 2) return;
  }
  const i64 o = cell_off(g, I);
  float s = l.r[o];
#pragma unroll
  for (int d = 0; d < D; d++) {
    float elo, ehi;
    if (g.per[d] && I[d] == 1) {
      const i64 w = o + (i64)(g.N[d] - 3) * g.s[d];
      elo = l.r[w] * l.iD[w];
    } else
      elo = l.eps[o - g.s[d]];
    if (g.per[d] && I[d] == g.N[d] - 2) {
      const i64 w = o - (i64)(g.N[d] - 3) * g.s[d];
      ehi = l.r[w] * l.iD[w];
    } else
      ehi = l.eps[o + g.s[d]];
    s -= elo * l.L[o + g.sc * d] + ehi * l.L[o + g.sc * d + g.s[d]];
  }
  l.eps[o] = s * l.iD[o];
}
template <int D>
__global__ void __launch_bounds__(512) k_gs_sweep(Lvl l, Box box, int k0) {
  pdl_wait();
  b_k_gs_sweep<D>(l, box, k0, real_block());
}

// increment!(p;ω) (src/Poisson.jl:100-104): r −= ω·Aϵ; x += ω·ϵ (periodic wrap = perBC!(ϵ)), optionally fused with
// L₂(p) = r⋅r (src/Poisson.jl:189) → out[slot].
template <int D>
__device__ __forceinline__ void b_k_increment(Lvl l, Box box, const float* wp, int x_is_zero, int with_l2, RedBuf R, int slot, const int3 vb) {
  const Grid& g = l.g;
  int I[3];
  const bool ok = thread_cell<D>(box, I, vb);
  double v[1] = {0.0}, fin[1];
  if (ok) {
    const float w = *wp;
    const i64 o = cell_off(g, I);
    i64 lo[3], hi[3];
    nbr_offsets<D>(g, I, lo, hi);
    const float Ae = mult_at<D>(g, l.eps, l.L, l.Dg, o, lo, hi);
    const float r = l.r[o] - w * Ae;
    l.r[o] = r;
    const float e = l.eps[o];
    l.x[o] = x_is_zero ? w * e : l.x[o] + w * e;
    v[0] = (double)r * (double)r;
  }
  if (with_l2) grid_reduce<RED_SUM, 1>(v, R, slot, fin);
}
template <int D>
__global__ void __launch_bounds__(512) k_increment(Lvl l, Box box, const float* __restrict__ wp, int x_is_zero, int with_l2, RedBuf R, int slot) {
  pdl_wait();
  b_k_increment<D>(l, box, wp, x_is_zero, with_l2, R, slot, real_block());
}

// prolongate!(fine.ϵ,coarse.x,c) + increment!(fine;ω) (src/MultiLevelPoisson.jl:50,99-100) without materialising ϵ:
// ϵ[I] = x_c[down(I,c)], down = (I+2)÷2 (1-based) ↔ (I0+1)/2 (0-based) in coarsened dims.
template <int D>
__device__ __forceinline__ void b_k_prolong_inc(Lvl l, Grid gc, const float* xc, Box box, const float* wp, int c0, int c1,
                                                     int c2, const int3 vb) {
  const Grid& g = l.g;
  int I[3];
  if (!thread_cell<D>(box, I, vb)) return;
  const int c[3] = {c0, c1, c2};
  const float w = *wp;
  const i64 o = cell_off(g, I);
  auto ec = [&](int a, int b, int cc) -> float {
    int J[3] = {a, b, cc};
#pragma unroll
    for (int d = 0; d < D; d++) {
      if (g.per[d]) {  // perBC!(ϵ): wrap to the opposite interior cell
        if (J[d] == 0) J[d] = g.N[d] - 2;
        else if (J[d] == g.N[d] - 1) J[d] = 1;
      }
      if (c[d]) J[d] = (J[d] + 1) / 2;
    }
    return xc[cell_off(gc, J)];
  };
  const float e = ec(I[0], I[1], I[2]);
  float s = e * l.Dg[o];
#pragma unroll
  for (int d = 0; d < D; d++) {
    int A[3] = {I[0], I[1], I[2]}, B[3] = {I[0], I[1], I[2]};
    A[d] -= 1;
    B[d] += 1;
    s += ec(A[0], A[1], A[2]) * l.L[o + g.sc * d] + ec(B[0], B[1], B[2]) * l.L[o + g.sc * d + g.s[d]];
  }
  l.r[o] = l.r[o] - w * s;
  l.x[o] = l.x[o] + w * e;
}
template <int D>
__global__ void __launch_bounds__(512) k_prolong_inc(Lvl l, Grid gc, const float* __restrict__ xc, Box box, const float* __restrict__ wp, int c0, int c1,
                                                     int c2) {
  pdl_wait();
  b_k_prolong_inc<D>(l, gc, xc, box, wp, c0, c1, c2, real_block());
}

// Velocity correction and pressure unscale of mom_project! (src/Flow.jl:227-230):
//   u[I,i] −= L[I,i]·(x[I] − x[I−δ_i]);  p = x/dt   (x keeps the scaled iterate; p is the observable pressure)
template <int D>
__global__ void __launch_bounds__(512) k_correct(Lvl l, Box box, float* __restrict__ u, float* __restrict__ p, const float* __restrict__ dtp, float w) {
  pdl_wait();
  const Grid& g = l.g;
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  i64 lo[3], hi[3];
  nbr_offsets<D>(g, I, lo, hi);
  const float dt = w * (*dtp);
  const float x = l.x[o];
#pragma unroll
  for (int d = 0; d < D; d++) u[o + g.sc * d] -= l.L[o + g.sc * d] * (x - l.x[o + lo[d]]);
  p[o] = x / dt;
}

// CFL (src/Flow.jl:234-244): σ = flux_out(I,u) on the interior; maximum(σ) over ALL cells (ghost σ holds stale Φ, lower
// ghosts 0).  The last block turns the max into Δt = min(10, 1/(max+5ν)) and stores it at dt_out.
template <int D>
__global__ void __launch_bounds__(512) k_cfl(Grid g, Box box, const float* __restrict__ u, float* __restrict__ sigma, float nu, float* __restrict__ dt_out,
                                             RedBuf R, int slot) {
  pdl_wait();
  int I[3];
  const bool ok = thread_cell<D>(box, I);
  double v[1] = {0.0}, fin[1];  // lower ghosts hold 0, so the max is at least 0
  if (ok) {
    const i64 o = cell_off(g, I);
    bool interior = true;
#pragma unroll
    for (int d = 0; d < D; d++) interior = interior && I[d] >= 1 && I[d] <= g.N[d] - 2;
    float s;
    if (interior) {
      s = 0.f;
#pragma unroll
      for (int d = 0; d < D; d++) s += fmaxf(0.f, u[o + g.sc * d + g.s[d]]) + fmaxf(0.f, -u[o + g.sc * d]);
      sigma[o] = s;
    } else
      s = sigma[o];
    v[0] = (double)s;
  }
  if (grid_reduce<RED_MAX, 1>(v, R, slot, fin)) {
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
      const float m = (float)fin[0];
      *dt_out = fminf(10.f, 1.f / (m + 5.f * nu));
    }
  }
}

// mult!(p,x) (src/Poisson.jl:63-69): z = A x on the interior (ghosts zeroed by the caller).
template <int D>
__global__ void k_mult(Lvl l, Box box, const float* __restrict__ x, float* __restrict__ z) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(l.g, I);
  i64 lo[3], hi[3];
  nbr_offsets<D>(l.g, I, lo, hi);
  z[o] = mult_at<D>(l.g, x, l.L, l.Dg, o, lo, hi);
}

// ---- pcg! pieces (src/Poisson.jl:166-186) -------------------------------------------------
// stage 0: z = ϵ = r·iD ; ρ = r⋅z                       (:168-169)
// stage 1: z = Aϵ ; σ = z⋅ϵ                              (:173-174)
// stage 2: x += αϵ ; r −= αz                            (:176-177)   α = sc[0]
// stage 3: z = r·iD ; ρ₂ = r⋅z                          (:179-180)
// stage 4: ϵ = βϵ + z                                    (:183)       β = sc[1]
// The iteration is driven ON THE DEVICE: the block that folds a dot product also forms ρ, α, β (Float32 divisions like the
// reference's) and evaluates pcg!'s early exits — abs(ρ) < 10eps, abs(α) ∉ [1e-2, 1e2] — into `ctl->active`; every later stage
// of the same pcg! call returns at once when it is 0.  The host enqueues all stages of the six iterations and never waits.
struct PcgCtl {
  float rho, alpha, beta;
  int active;
};
template <int D>
__global__ void __launch_bounds__(512) k_pcg(Lvl l, Box box, int stage, PcgCtl* __restrict__ ctl, RedBuf R, int slot) {
  pdl_wait();
  const Grid& g = l.g;
  int I[3];
  if (stage > 0 && ctl->active == 0) return;  // (written by an earlier launch only: the last block of a stage writes it after every block has started)
  const bool ok = thread_cell<D>(box, I);
  double v[1] = {0.0}, fin[1];
  if (ok) {
    const i64 o = cell_off(g, I);
    if (stage == 0) {
      const float z = l.r[o] * l.iD[o];
      l.z[o] = z;
      l.eps[o] = z;
      v[0] = (double)l.r[o] * (double)z;
    } else if (stage == 1) {
      i64 lo[3], hi[3];
      nbr_offsets<D>(g, I, lo, hi);
      const float z = mult_at<D>(g, l.eps, l.L, l.Dg, o, lo, hi);
      l.z[o] = z;
      v[0] = (double)z * (double)l.eps[o];
    } else if (stage == 2) {
      const float a = ctl->alpha;
      l.x[o] += a * l.eps[o];
      l.r[o] -= a * l.z[o];
    } else if (stage == 3) {
      const float z = l.r[o] * l.iD[o];
      l.z[o] = z;
      v[0] = (double)l.r[o] * (double)z;
    } else {
      l.eps[o] = ctl->beta * l.eps[o] + l.z[o];
    }
  }
  if (stage == 0 || stage == 1 || stage == 3) {
    const bool last = grid_reduce<RED_SUM, 1>(v, R, slot, fin);
    if (last && threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
      const float eps10 = 10.f * 1.1920929e-7f;
      const float d = (float)fin[0];
      if (stage == 0) {
        ctl->rho = d;
        ctl->active = fabsf(d) < eps10 ? 0 : 1;
      } else if (stage == 1) {
        const float alpha = ctl->rho / d;
        ctl->alpha = alpha;
        if ((double)fabsf(alpha) < 1e-2 || (double)fabsf(alpha) > 1e2) ctl->active = 0;  // alpha should be O(1) (Float64 literals in the reference)
      } else {
        if (fabsf(d) < eps10)
          ctl->active = 0;
        else {
          ctl->beta = d / ctl->rho;
          ctl->rho = d;
        }
      }
    }
  }
}

// L₂(p) = r⋅r and L∞(p) = maximum(abs,r) (src/Poisson.jl:189-190) as a standalone reduction.
template <int D>
__global__ void __launch_bounds__(512) k_norms(Lvl l, Box box, RedBuf R, int slot, int want_max) {
  pdl_wait();
  int I[3];
  const bool ok = thread_cell<D>(box, I);
  double v[1] = {0.0}, fin[1];
  if (ok) {
    const float r = l.r[cell_off(l.g, I)];
    v[0] = want_max ? (double)fabsf(r) : (double)r * (double)r;
  }
  if (want_max)
    grid_reduce<RED_MAX, 1>(v, R, slot, fin);
  else
    grid_reduce<RED_SUM, 1>(v, R, slot, fin);
}

// generic helpers
__global__ void k_fill(float* __restrict__ a, size_t n, float v) {
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) a[i] = v;
}
__global__ void k_set_scalar(float* p, float v) {
  pdl_wait();
  *p = v;
}

// Exhaustive self-test of div6: counts the float bit patterns (all 2^32) where div6(x) and x/6.f differ (NaNs compare equal).
__global__ void k_selftest_div6(unsigned long long* nbad) {
  pdl_wait();
  unsigned long long bad = 0;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned)b);
    const float a = div6(x), e = x / 6.f;
    const bool same = (__float_as_uint(a) == __float_as_uint(e)) || (a != a && e != e);
    if (!same) bad++;
  }
  if (bad) atomicAdd(nbad, bad);
}

// z slabs: Δt from the all-reduced maxima (the single-GPU path does this in f_cfl's last block)
__global__ void k_cfl_final(RedBuf R, int slot_a, int slot_b, float nu, float* dt_out) {
  pdl_wait();
  const float mm = (float)fmax(R.out[slot_a], R.out[slot_b]);
  *dt_out = fminf(10.f, 1.f / (mm + 5.f * nu));
}

// ------------------------------------------------------------------------------------------------
// Peer-to-peer halo exchange (z slabs): one launch pushes this rank's boundary planes straight into the neighbours' ghost planes
// with NVLink stores and returns only when the neighbours' planes have arrived here.  Flags live in each rank's `mbox`
// (ints: [0] ready-from-lower, [1] ready-from-upper, [2] arrived-from-lower, [3] arrived-from-upper, [4] block counter,
// [5] error) and carry the exchange's sequence number, identical on every rank.
//   A. tell both neighbours "my ghost planes may be overwritten" (everything that read them is earlier in this stream);
//   B. wait until both neighbours said so;            C. copy;            D. fence, tell them "arrived", wait for theirs.
// Every wait is bounded (≈30 s): a lost partner raises the error word on the whole ring instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
struct HaloSegs {
  int nseg;
  const float* src[16];
  float* dst[16];
};

__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_flag_sys(int* p, int v) { asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// `no_ready`: the caller knows that the neighbours finished reading the destination planes before their previous exchange with
// this rank (every exchange is a meeting of neighbours, so having passed exchange n−1 proves the neighbour had entered it): the
// "ready to receive" round trip is skipped and the kernel is copy → arrived → wait for the neighbours' arrived.
__global__ void __launch_bounds__(256) k_halo_push(HaloSegs segs, int cnt4, int seq, int* my, int* peer_lo, int* peer_hi, long long timeout,
                                                   const int* myflags, int* flags_lo, int* flags_hi, int no_ready) {
  pdl_wait();
  // A timeout is FATAL for the whole ring: the rank that gave up raises the error word of its own mailbox and of both neighbours'
  // (my[5]), never copies and never signals `arrived`, and every later exchange on a rank whose error word is set returns at once;
  // the host reads the word after every step (check_flags) and fails the call.  NCCL collectives block without a bound anyway,
  // so the bound (≈30 s) only has to catch a dead peer.
  __shared__ int ok;
  if (threadIdx.x == 0) {
    int good = ld_flag(my + 5) == 0;
    if (good && blockIdx.x == 0) {
      // velocity planes: a field that failed its range check (range_note, wl_common.cuh) fails it for the neighbours that read
      // its halo planes too — the word lands before `arrived`, i.e. before the receiver's flux kernel can start
      if (myflags && ld_flag(myflags + 2) != 0) {
        if (flags_lo) st_flag_sys(flags_lo + 2, 1);
        if (flags_hi) st_flag_sys(flags_hi + 2, 1);
      }
      __threadfence_system();
      if (!no_ready) {
        if (peer_lo) st_flag_sys(peer_lo + 1, seq);  // I am the upper neighbour of my lower neighbour
        if (peer_hi) st_flag_sys(peer_hi + 0, seq);
      }
    }
    const long long t0 = clock64();
    while (good && !no_ready && ((peer_lo && ld_flag(my + 0) < seq) || (peer_hi && ld_flag(my + 1) < seq))) {
      if (clock64() - t0 > timeout || ld_flag(my + 5) != 0) good = 0;
    }
    if (!good) {
      st_flag_sys(my + 5, 1);
      if (peer_lo) st_flag_sys(peer_lo + 5, 1);
      if (peer_hi) st_flag_sys(peer_hi + 5, 1);
    }
    ok = good;
  }
  __syncthreads();
  if (!ok) return;
  const long long total = (long long)segs.nseg * cnt4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int sgi = (int)(q / cnt4);
    const int e = (int)(q - (long long)sgi * cnt4);
    reinterpret_cast<float4*>(segs.dst[sgi])[e] = reinterpret_cast<const float4*>(segs.src[sgi])[e];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(my + 4, 1);
    if (t == (int)gridDim.x - 1) {
      atomicExch(my + 4, 0);
      __threadfence_system();
      if (peer_lo) st_flag_sys(peer_lo + 3, seq);
      if (peer_hi) st_flag_sys(peer_hi + 2, seq);
      const long long t0 = clock64();
      while ((peer_lo && ld_flag(my + 2) < seq) || (peer_hi && ld_flag(my + 3) < seq)) {
        if (clock64() - t0 > timeout || ld_flag(my + 5) != 0) {
          st_flag_sys(my + 5, 1);
          if (peer_lo) st_flag_sys(peer_lo + 5, 1);
          if (peer_hi) st_flag_sys(peer_hi + 5, 1);
          break;
        }
      }
      __threadfence_system();
    }
  }
}

// All-reduce of one or two doubles across ≤ 8 ranks over NVLink peer memory, one warp: every rank stores its values and then the
// sequence number into its own row of every rank's mailbox (two buffers alternate with the sequence's parity: a rank can be at most
// one all-reduce ahead of another, because it needs everybody's values of call n to finish call n), waits until all rows of its own
// mailbox carry the sequence, and folds them in rank order — the same order on every rank, so all ranks get the same bits.
// The result goes to out[] and, tagged, to the host-visible mirror (RedBuf::hout / hseq) like a local reduction's.
struct ArPeers {
  double* p[8];  // rank q's mailbox as mapped into this process: [2 buffers][8 source ranks][4 doubles: v0, v1, sequence, pad]
};
__global__ void __launch_bounds__(32) k_allreduce(ArPeers peers, int P, int rank, long long seq, int op, int count, double* out, double* hout,
                                                  unsigned int* hseq, unsigned int tag, int* err, long long timeout) {
  pdl_wait();
  const int t = threadIdx.x;
  const int b = (int)(seq & 1);
  const double v0 = out[0], v1 = count > 1 ? out[1] : 0.0;
  if (t < P) {
    volatile double* dst = peers.p[t] + (b * 8 + rank) * 4;
    dst[0] = v0;
    dst[1] = v1;
    __threadfence_system();
    *reinterpret_cast<volatile long long*>(dst + 2) = seq;
    volatile double* src = peers.p[rank] + (b * 8 + t) * 4;
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile long long*>(src + 2) != seq) {
      if (clock64() - t0 > timeout || ld_flag(err) != 0) {
        st_flag_sys(err, 1);
        break;
      }
    }
  }
  __syncwarp();
  if (t == 0) {
    __threadfence_system();
    volatile double* mine = peers.p[rank] + b * 8 * 4;
    double a0 = op == RED_SUM ? 0.0 : -1.0e300, a1 = a0;
    for (int q = 0; q < P; q++) {
      const double w0 = mine[q * 4], w1 = mine[q * 4 + 1];
      a0 = op == RED_SUM ? a0 + w0 : (a0 > w0 ? a0 : w0);
      a1 = op == RED_SUM ? a1 + w1 : (a1 > w1 ? a1 : w1);
    }
    out[0] = a0;
    if (count > 1) out[1] = a1;
    if (tag) {
      hout[0] = a0;
      if (count > 1) hout[1] = a1;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned int*>(hseq) = tag;
    }
  }
}

// All-gather of a replicated level's field over peer memory: every rank stores its own part (n4 float4 at the same offset on every
// rank) into the other ranks' copies, then the ranks meet at a barrier built like k_allreduce's (sequence words in a second mailbox
// region).  In place of ncclAllGather inside the V-cycle: one launch, no staging, ≈ 54 → 20 µs at 8 ranks.
struct BcastDst {
  float4* p[8];  // rank q's copy of MY part (null for this rank)
};
__global__ void __launch_bounds__(256) k_bcast_planes(const float4* __restrict__ src, BcastDst dst, long long n4, ArPeers peers, int P, int rank, long long seq,
                                                      int* counter, int* err, long long timeout) {
  pdl_wait();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[e];
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (q < P && dst.p[q]) dst.p[q][e] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x >= 32) return;
  int lastb = 0;
  if (threadIdx.x == 0) {
    lastb = atomicAdd(counter, 1) == (int)gridDim.x - 1;
    if (lastb) atomicExch(counter, 0);
  }
  lastb = __shfl_sync(0xffffffffu, lastb, 0);
  if (!lastb) return;
  __threadfence_system();
  const int t = threadIdx.x, b = (int)(seq & 1);
  if (t < P) {
    *reinterpret_cast<volatile long long*>(peers.p[t] + (b * 8 + rank) * 4 + 2) = seq;
    volatile double* srcw = peers.p[rank] + (b * 8 + t) * 4;
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile long long*>(srcw + 2) != seq) {
      if (clock64() - t0 > timeout || ld_flag(err) != 0) {
        st_flag_sys(err, 1);
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

// Range check of a velocity field the library did not write itself (uploads, wl_apply_bc, kernels without the built-in check):
// raises flags[2] / flags[0] as described at range_note (wl_common.cuh).  n4 = number of float4 to scan.
__global__ void __launch_bounds__(256) k_range_check(const float4* __restrict__ a, long long n4, int* __restrict__ flags) {
  pdl_wait();
  RangeAcc ra;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) ra.add4(a[i]);
  ra.publish(flags);
}

// ================================================================================================
// measure!(flow, body; t, ϵ) on the device (src/Body.jl:28-60, src/AutoBody.jl:29-37, src/Body.jl:88-107) for parametrised bodies:
// primitives (sphere / torus) with a rigid translation map, combined by lazy set operations.  One thread per interior cell.
// The arithmetic follows the reference expression by expression (the gradients are the closed forms ForwardDiff produces for
// these sdfs; sinpi / cospi are evaluated in double and rounded, like the oracle's restatement).
// ================================================================================================
struct BodyPrim {  // = wl_body_prim (include/wl_b200.h)
  int kind, op;
  float c[3];
  float R, r;
  float vel[3];
};
struct BodySet {
  int np;
  BodyPrim p[8];
};
struct Meas {
  float d, n[3], V[3];
};
__device__ __forceinline__ float prim_sdf(const BodyPrim& b, int D, const float* x) {
  if (b.kind == 0) {
    float s = 0.f;
    for (int d = 0; d < D; d++) s += (x[d] - b.c[d]) * (x[d] - b.c[d]);
    return sqrtf(s) - b.R;
  }
  const float y = x[1] - b.c[1], z = x[2] - b.c[2], xx = x[0] - b.c[0];
  const float q = sqrtf(y * y + z * z) - b.R;
  return sqrtf(q * q + xx * xx) - b.r;
}
__device__ __forceinline__ void prim_grad(const BodyPrim& b, int D, const float* x, float* n) {
  if (b.kind == 0) {
    float s = 0.f;
    for (int d = 0; d < D; d++) s += (x[d] - b.c[d]) * (x[d] - b.c[d]);
    const float m = sqrtf(s);
    for (int d = 0; d < D; d++) n[d] = (x[d] - b.c[d]) / m;
    return;
  }
  const float y = x[1] - b.c[1], z = x[2] - b.c[2], xx = x[0] - b.c[0];
  const float rho = sqrtf(y * y + z * z);
  const float q = rho - b.R;
  const float m = sqrtf(q * q + xx * xx);
  n[0] = xx / m;
  n[1] = (q / m) * (y / rho);
  n[2] = (q / m) * (z / rho);
}
// measure(body::AutoBody, x, t; fastd²) with map(x,t) = x − vel·t: J = I, V = −J\∂ₜmap = vel
__device__ __forceinline__ Meas prim_measure(const BodyPrim& p, int D, const float* x, float t, float fastd2) {
  float xi[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < D; k++) xi[k] = x[k] - p.vel[k] * t;
  Meas m;
  m.d = prim_sdf(p, D, xi);
  for (int k = 0; k < 3; k++) m.n[k] = m.V[k] = 0.f;
  if (m.d * m.d > fastd2) return m;
  float gk[3] = {0.f, 0.f, 0.f};
  prim_grad(p, D, xi, gk);
  for (int k = 0; k < D; k++)
    if (isnan(gk[k])) return m;
  float mm = 0.f;
  for (int k = 0; k < D; k++) mm += gk[k] * gk[k];
  mm = sqrtf(mm);
  m.d /= mm;
  for (int k = 0; k < D; k++) {
    m.n[k] = gk[k] / mm;
    m.V[k] = p.vel[k];
  }
  return m;
}
__device__ __forceinline__ bool fless(float a, float b) {  // isless(::Float32, ::Float32)
  if (isnan(a)) return false;
  if (isnan(b)) return true;
  if (a == b) return signbit(a) && !signbit(b);
  return a < b;
}
__device__ __forceinline__ bool meas_less(const Meas& a, const Meas& b, int D) {  // isless on the tuples (d, n, V)
  if (fless(a.d, b.d)) return true;
  if (fless(b.d, a.d)) return false;
  for (int k = 0; k < D; k++) {
    if (fless(a.n[k], b.n[k])) return true;
    if (fless(b.n[k], a.n[k])) return false;
  }
  for (int k = 0; k < D; k++) {
    if (fless(a.V[k], b.V[k])) return true;
    if (fless(b.V[k], a.V[k])) return false;
  }
  return false;
}
__device__ __forceinline__ Meas csg_measure(const BodySet& B, int D, const float* x, float t, float fastd2) {
  Meas acc = prim_measure(B.p[0], D, x, t, fastd2);
  for (int q = 1; q < B.np; q++) {
    Meas m = prim_measure(B.p[q], D, x, t, fastd2);
    if (B.p[q].op == 2) {
      m.d = -m.d;
      for (int k = 0; k < D; k++) m.n[k] = -m.n[k];
    }
    if (B.p[q].op == 0) {
      if (meas_less(m, acc, D)) acc = m;
    } else {
      if (!meas_less(m, acc, D)) acc = m;
    }
  }
  return acc;
}
__device__ __forceinline__ float ulp_f(float d) {  // eps(d::Float32)
  const float a = fabsf(d);
  return __uint_as_float(__float_as_uint(a) + 1u) - a;
}
__device__ __forceinline__ float sinpi_f(float x) { return (float)sinpi((double)x); }
__device__ __forceinline__ float cospi_f(float x) { return (float)cospi((double)x); }
__device__ __forceinline__ float kern0_f(float d) { return (1.f + d + sinpi_f(d) / 3.14159274101257324f) / 2.f; }
__device__ __forceinline__ float kern1_f(float d) {
  return (1.f - d * d) / 4.f - (d * sinpi_f(d) + (1.f + cospi_f(d)) / 3.14159274101257324f) / (2.f * 3.14159274101257324f);
}
__device__ __forceinline__ float mu0_f(float d, float e) { return d / e < -1.f + sqrtf(ulp_f(d)) ? 0.f : kern0_f(fminf(d / e, 1.f)); }
__device__ __forceinline__ float mu1_f(float d, float e) { return e * kern1_f(fmaxf(-1.f, fminf(d / e, 1.f))); }

template <int D>
__global__ void __launch_bounds__(256) k_measure(const __grid_constant__ Grid g, Box box, const __grid_constant__ BodySet B, float eps, float t,
                                                 float* __restrict__ sigma, float* __restrict__ V, float* __restrict__ mu0, float* __restrict__ mu1) {
  pdl_wait();
  int I[3];
  if (!thread_cell<D>(box, I)) return;
  const i64 o = cell_off(g, I);
  const float d2 = (2.f + eps) * (2.f + eps);
  // loc(0,I): cell centre, in GLOBAL coordinates (1-based index − 1.5; g.zoff = global z index of local plane 0)
  float xc[3] = {0.f, 0.f, 0.f};
  for (int d = 0; d < D; d++) xc[d] = (float)(I[d] + 1 + (d == 2 ? g.zoff : 0)) - 1.5f;
  float dI;
  if (B.np == 1) {  // sdf(body::AutoBody,x,t) = body.sdf(body.map(x,t),t)
    float xi[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < D; k++) xi[k] = xc[k] - B.p[0].vel[k] * t;
    dI = prim_sdf(B.p[0], D, xi);
  } else  // sdf(body::SetBody,…) = measure(body,x,t;fastd²)[1]
    dI = csg_measure(B, D, xc, t, d2).d;
  sigma[o] = dI;
  // V = 0, μ₀ = 1, μ₁ = 0 outside the band (the reference resets the arrays first)
  float v[3] = {0.f, 0.f, 0.f}, m0[3] = {1.f, 1.f, 1.f}, m1[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (dI * dI < d2) {
    for (int i = 0; i < D; i++) {
      float x[3] = {xc[0], xc[1], xc[2]};
      x[i] = xc[i] - 0.5f;  // loc(i,I)
      const Meas m = csg_measure(B, D, x, t, d2);
      const float di = fabsf(m.d) <= 0.5f ? m.d : copysignf(m.d, dI);
      v[i] = m.V[i];
      m0[i] = mu0_f(di, eps);
      const float k1 = mu1_f(di, eps);
      for (int j = 0; j < D; j++) m1[i + D * j] = k1 * m.n[j];
    }
  } else if (dI < 0.f) {
    for (int i = 0; i < D; i++) m0[i] = 0.f;
  }
  for (int i = 0; i < D; i++) {
    V[o + g.sc * i] = v[i];
    mu0[o + g.sc * i] = m0[i];
    for (int j = 0; j < D; j++) mu1[o + g.sc * (i + D * j)] = m1[i + D * j];
  }
}

// ================================================================================================
// Forces and moments on the body (src/Metrics.jl:111-190) as ONE fused device reduction:
//   pressure_force  Σ p[I]·nds            viscous_force  Σ −2ν·S(I,u)·nds
//   pressure_moment Σ p[I]·(x−x₀)×nds     viscous_moment Σ −2ν·(x−x₀)×(S(I,u)·nds)
// with nds(body,x,t) = n·kern(clamp(d,−1,1)), (d,n) = measure(body,x,t,fastd²=1), kern(d) = (1+cospi(d))/2, over inside(p).
// The per-cell vectors are Float32 like the reference's df; the sums are Float64 (sum(Float64, df)), deterministic.
// out[slot0 + 0:3 | 3:6 | 6:9 | 9:12].
// ================================================================================================
template <int D>
__global__ void __launch_bounds__(256) k_forces(const __grid_constant__ Grid g, Box box, const __grid_constant__ BodySet B, float t, float nu, float x00,
                                                float x01, float x02, const float* __restrict__ u, const float* __restrict__ p, RedBuf R, int slot0) {
  pdl_wait();
  int I[3];
  const bool ok = thread_cell<D>(box, I);
  double v[12], fin[12];
#pragma unroll
  for (int q = 0; q < 12; q++) v[q] = 0.0;
  if (ok) {
    const i64 o = cell_off(g, I);
    float x[3] = {0.f, 0.f, 0.f};
    for (int d = 0; d < D; d++) x[d] = (float)(I[d] + 1 + (d == 2 ? g.zoff : 0)) - 1.5f;
    const Meas m = csg_measure(B, D, x, t, 1.f);
    const float kd = (1.f + cospi_f(fmaxf(-1.f, fminf(m.d, 1.f)))) / 2.f;
    float nds[3] = {0.f, 0.f, 0.f};
    for (int d = 0; d < D; d++) nds[d] = m.n[d] * kd;
    if (nds[0] != 0.f || nds[1] != 0.f || nds[2] != 0.f) {  // (every term below is a multiple of nds: exact zeros elsewhere)
      const float nu2 = -2.f * nu;
      auto U = [&](i64 off, int i) { return u[o + off + g.sc * i]; };
      auto dd = [&](int i, int j) -> float {  // ∂(i,j,I,u)
        if (i == j) return U(g.s[i], i) - U(0, i);
        return (U(g.s[j], i) + U(g.s[j] + g.s[i], i) - U(-g.s[j], i) - U(-g.s[j] + g.s[i], i)) / 4.f;
      };
      float S[3][3], Sn[3] = {0.f, 0.f, 0.f}, fv[3] = {0.f, 0.f, 0.f}, r[3] = {0.f, 0.f, 0.f};
      for (int a = 0; a < D; a++)
        for (int b = 0; b < D; b++) S[a][b] = (dd(a, b) + dd(b, a)) / 2.f;
      for (int a = 0; a < D; a++) {
        float s1 = 0.f, s2 = 0.f;
        for (int b = 0; b < D; b++) {
          s1 = b == 0 ? S[a][b] * nds[b] : s1 + S[a][b] * nds[b];
          s2 = b == 0 ? (nu2 * S[a][b]) * nds[b] : s2 + (nu2 * S[a][b]) * nds[b];
        }
        Sn[a] = s1;
        fv[a] = s2;
      }
      const float x0[3] = {x00, x01, x02};
      for (int d = 0; d < D; d++) r[d] = x[d] - x0[d];
      const float pr = p[o];
      float pm[3] = {0.f, 0.f, 0.f}, vm[3] = {0.f, 0.f, 0.f};
      if (D == 3) {
        const float c1[3] = {r[1] * nds[2] - r[2] * nds[1], r[2] * nds[0] - r[0] * nds[2], r[0] * nds[1] - r[1] * nds[0]};
        const float c2[3] = {r[1] * Sn[2] - r[2] * Sn[1], r[2] * Sn[0] - r[0] * Sn[2], r[0] * Sn[1] - r[1] * Sn[0]};
        for (int d = 0; d < 3; d++) pm[d] = pr * c1[d], vm[d] = nu2 * c2[d];
      } else {
        const float c1 = r[0] * nds[1] - r[1] * nds[0], c2 = r[0] * Sn[1] - r[1] * Sn[0];
        pm[0] = pm[1] = pr * c1;
        vm[0] = vm[1] = nu2 * c2;
      }
      for (int d = 0; d < D; d++) {
        v[d] = (double)(pr * nds[d]);
        v[3 + d] = (double)fv[d];
        v[6 + d] = (double)pm[d];
        v[9 + d] = (double)vm[d];
      }
    }
  }
  grid_reduce<RED_SUM, 12>(v, R, slot0, fin);
}

// update!(meanflow, flow) (src/Metrics.jl:236-247): P = ε·p + (1−ε)·P, U = ε·u + (1−ε)·U, UU[i,j] = ε·(u_i·u_j) + (1−ε)·UU[i,j]
// over every cell of the arrays (ghosts included; the pitched layout's padding rides along).  n = floats of one scalar field.
__global__ void __launch_bounds__(256) k_meanflow(long long n, int D, int uu, float eps, const float* __restrict__ p, const float* __restrict__ u,
                                                  float* __restrict__ P, float* __restrict__ U, float* __restrict__ UU) {
  pdl_wait();
  const float om = 1.f - eps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    P[i] = eps * p[i] + om * P[i];
    float v[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < D; c++) {
      v[c] = u[i + n * c];
      U[i + n * c] = eps * v[c] + om * U[i + n * c];
    }
    if (uu)
      for (int a = 0; a < D; a++)
        for (int b = 0; b < D; b++) UU[i + n * (a + D * b)] = eps * (v[a] * v[b]) + om * UU[i + n * (a + D * b)];
  }
}
