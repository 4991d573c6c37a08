// wl_oracle.cpp — CPU restatement of WaterLily.jl's `mom_step!` hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle (and the CPU baseline
// "port") for the B200 library in waterlily.jl_b200/csrc.  Nothing on the product
// path may link, import or call it: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs do.
//
// PARITY PINNING: the reference is pure Julia and Julia is not installed in this
// image, so the reference cannot be executed here ("oracle/_ref" does not exist).
// The oracle is pinned by the reference's own known-answer tests instead
// (test/test_poisson.jl, test/test_flow.jl, test/test_core.jl, test/test_bodies.jl;
// see tests/test_oracle_golden.py).  Reductions (`sum`, `⋅`, `maximum`) whose
// summation order lives in Julia Base/BLAS are accumulated in double and rounded to
// Float32 ("order unpinned").  Julia's @fastmath reassociation cannot
// be reproduced bit for bit; parity against real Julia output is therefore UNPINNED
// beyond those known answers.
//
// Every loop below is one reference `@loop` (src/core.jl:125-156), in the same order,
// over the same CartesianIndices range, writing the same arrays — including the quirks
// listed in SURVEY.md App. A.9.  Indices are 1-based like Julia: N_k = n_k+2 cells per
// dimension including ghosts, x fastest, vector component slowest.  2-D fields are
// stored as N1 x N2 x 1.
//
// Build: see oracle/Makefile  (g++ -O3 -march=x86-64-v3 -fopenmp -shared -fPIC).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef float T;

namespace {

struct I3 {
  int v[3];
};
static inline I3 shift(I3 a, int d, int s = 1) {
  a.v[d] += s;
  return a;
}
static inline I3 setj(I3 a, int d, int val) {  // CIj (src/core.jl:31)
  a.v[d] = val;
  return a;
}

struct Grid {
  int D;
  int N[3];  // incl. ghosts; N[2]==1 in 2-D
  size_t n() const { return (size_t)N[0] * N[1] * N[2]; }
  size_t at(const I3& I) const {
    return (size_t)(I.v[0] - 1) + (size_t)N[0] * ((size_t)(I.v[1] - 1) + (size_t)N[1] * (size_t)(I.v[2] - 1));
  }
};

struct Range {
  int lo[3], hi[3];
  size_t count() const {
    size_t c = 1;
    for (int d = 0; d < 3; d++) c *= (size_t)std::max(0, hi[d] - lo[d] + 1);
    return c;
  }
};

// CartesianIndices(a)  — every cell incl. ghosts
static Range all_cells(const Grid& g) {
  Range R;
  for (int d = 0; d < 3; d++) {
    R.lo[d] = 1;
    R.hi[d] = g.N[d];
  }
  return R;
}
// inside(a)  src/core.jl:47
static Range inside(const Grid& g) {
  Range R = all_cells(g);
  for (int d = 0; d < g.D; d++) {
    R.lo[d] = 2;
    R.hi[d] = g.N[d] - 1;
  }
  return R;
}
// inside_u(dims,j)  src/core.jl:55-57 : 3:N_j-1 in j, 2:N_k (incl. upper ghost) elsewhere
static Range inside_u(const Grid& g, int j) {
  Range R = all_cells(g);
  for (int d = 0; d < g.D; d++) {
    R.lo[d] = (d == j) ? 3 : 2;
    R.hi[d] = (d == j) ? g.N[d] - 1 : g.N[d];
  }
  return R;
}
// slice(dims,i,j,low)  src/core.jl:188-190
static Range slice(const Grid& g, int s, int j, int low = 1) {
  Range R = all_cells(g);
  for (int d = 0; d < g.D; d++) {
    R.lo[d] = (d == j) ? s : low;
    R.hi[d] = (d == j) ? s : g.N[d];
  }
  return R;
}

// One reference `@loop`: a parallel loop over a CartesianIndices range.
template <class F>
static void loop(const Range& R, F f) {
  if (R.count() == 0) return;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = R.lo[2]; k <= R.hi[2]; k++)
    for (int j = R.lo[1]; j <= R.hi[1]; j++)
      for (int i = R.lo[0]; i <= R.hi[0]; i++) f(I3{{i, j, k}});
}

// ---- reductions (Julia Base / BLAS; order unpinned → double accumulation) ----------
static double sum_range(const Grid& g, const T* a, const Range& R) {
  double s = 0;
#pragma omp parallel for collapse(2) reduction(+ : s) schedule(static)
  for (int k = R.lo[2]; k <= R.hi[2]; k++)
    for (int j = R.lo[1]; j <= R.hi[1]; j++)
      for (int i = R.lo[0]; i <= R.hi[0]; i++) s += (double)a[g.at(I3{{i, j, k}})];
  return s;
}
static double dot_range(const Grid& g, const T* a, const T* b, const Range& R) {
  double s = 0;
#pragma omp parallel for collapse(2) reduction(+ : s) schedule(static)
  for (int k = R.lo[2]; k <= R.hi[2]; k++)
    for (int j = R.lo[1]; j <= R.hi[1]; j++)
      for (int i = R.lo[0]; i <= R.hi[0]; i++) {
        size_t o = g.at(I3{{i, j, k}});
        s += (double)a[o] * (double)b[o];
      }
  return s;
}

// ---- limiters and face fluxes: src/Flow.jl:1-36 --------------------------------------
static inline T median3(T a, T b, T c) {  // src/Flow.jl:27-36
  if (a > b) {
    if (b >= c) return b;
    if (a > c) return c;
  } else {
    if (b <= c) return b;
    if (a < c) return c;
  }
  return a;
}
static inline T quick(T u, T c, T d) { return median3((5 * c + 2 * d - u) / 6, c, median3(10 * c - 9 * u, c, d)); }
static inline T vanLeer(T u, T c, T d) {
  return (c <= std::min(u, d) || c >= std::max(u, d)) ? c : c + (d - c) * (c - u) / (d - u);
}
static inline T cds(T u, T c, T d) { return (c + d) / 2; }
enum { LAM_QUICK = 0, LAM_CDS = 1, LAM_VANLEER = 2 };
static inline T lam(int l, T u, T c, T d) { return l == LAM_QUICK ? quick(u, c, d) : (l == LAM_CDS ? cds(u, c, d) : vanLeer(u, c, d)); }

struct FluxCtx {
  const Grid& g;
  const T* f;  // the advected component u[:, i]
  int a;       // flux direction j
  int l;       // limiter
  T at(const I3& I) const { return f[g.at(I)]; }
  T phi(const I3& I) const { return (at(I) + at(shift(I, a, -1))) / 2; }  // ϕ  :3
  T phiu(const I3& I, T u) const {                                        // ϕu :8
    return u > 0 ? u * lam(l, at(shift(I, a, -2)), at(shift(I, a, -1)), at(I)) : u * lam(l, at(shift(I, a, 1)), at(I), at(shift(I, a, -1)));
  }
  T phiuP(const I3& Ip, const I3& I, T u) const {  // ϕuP :9
    return u > 0 ? u * lam(l, at(Ip), at(shift(I, a, -1)), at(I)) : u * lam(l, at(shift(I, a, 1)), at(I), at(shift(I, a, -1)));
  }
  T phiuL(const I3& I, T u) const {  // ϕuL :10
    return u > 0 ? u * phi(I) : u * lam(l, at(shift(I, a, 1)), at(I), at(shift(I, a, -1)));
  }
  T phiuR(const I3& I, T u) const {  // ϕuR :11
    return u < 0 ? u * phi(I) : u * lam(l, at(shift(I, a, -2)), at(shift(I, a, -1)), at(I));
  }
  T d(const I3& I) const { return at(I) - at(shift(I, a, -1)); }  // ∂(j,CI(I,i),u) :1
};

// conv_diff!  src/Flow.jl:38-62
static void conv_diff(const Grid& g, T* r, const T* u, T* Phi, int l, T nu, const int* per) {
  const size_t n = g.n();
  const int D = g.D;
#pragma omp parallel for schedule(static)
  for (size_t o = 0; o < n * D; o++) r[o] = 0;  // r .= 0  :39
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) {
      T* ri = r + (size_t)i * n;
      const T* uj = u + (size_t)j * n;
      FluxCtx F{g, u + (size_t)i * n, j, l};
      auto uface = [&](const I3& I) { return (uj[g.at(I)] + uj[g.at(shift(I, i, -1))]) / 2; };  // ϕ(i,CI(I,j),u)
      // lowerBoundary!  :56 / :60-61
      if (!per[j]) {
        loop(slice(g, 2, j, 2), [&](I3 I) { ri[g.at(I)] += F.phiuL(I, uface(I)) - nu * F.d(I); });
      } else {
        loop(slice(g, 2, j, 2), [&](I3 I) {
          size_t o = g.at(I);
          Phi[o] = F.phiuP(setj(I, j, g.N[j] - 2), I, uface(I)) - nu * F.d(I);
          ri[o] += Phi[o];
        });
      }
      // inner cells  :47-49
      loop(inside_u(g, j), [&](I3 I) {
        size_t o = g.at(I);
        Phi[o] = F.phiu(I, uface(I)) - nu * F.d(I);
        ri[o] += Phi[o];
      });
      loop(inside_u(g, j), [&](I3 I) { ri[g.at(shift(I, j, -1))] -= Phi[g.at(I)]; });
      // upperBoundary!  :57 / :62
      if (!per[j]) {
        loop(slice(g, g.N[j], j, 2), [&](I3 I) { ri[g.at(shift(I, j, -1))] += -F.phiuR(I, uface(I)) + nu * F.d(I); });
      } else {
        loop(slice(g, g.N[j], j, 2), [&](I3 I) { ri[g.at(shift(I, j, -1))] -= Phi[g.at(setj(I, j, 2))]; });
      }
    }
}

// BC!  src/core.jl:200-219 with a constant (tuple) uBC.  The tangential rule
// `uBC + a[I±δ] − uBC` is a plain copy (what @fastmath folds it to; ≤1 ulp otherwise).
static void BC(const Grid& g, T* a, const T* U, bool saveexit, const int* per) {
  const size_t n = g.n();
  const int D = g.D;
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) {
      T* ai = a + (size_t)i * n;
      const int Nj = g.N[j];
      if (per[j]) {
        loop(slice(g, 1, j), [&](I3 I) { ai[g.at(I)] = ai[g.at(setj(I, j, Nj - 1))]; });
        loop(slice(g, Nj, j), [&](I3 I) { ai[g.at(I)] = ai[g.at(setj(I, j, 2))]; });
      } else if (i == j) {
        for (int s = 1; s <= 2; s++) loop(slice(g, s, j), [&](I3 I) { ai[g.at(I)] = U[i]; });
        if (!saveexit || i > 0) loop(slice(g, Nj, j), [&](I3 I) { ai[g.at(I)] = U[i]; });
      } else {
        loop(slice(g, 1, j), [&](I3 I) { ai[g.at(I)] = ai[g.at(shift(I, j, 1))]; });
        loop(slice(g, Nj, j), [&](I3 I) { ai[g.at(I)] = ai[g.at(shift(I, j, -1))]; });
      }
    }
}

// exitBC!  src/core.jl:226-233
static void exitBC(const Grid& g, T* u, const T* u0, T dt) {
  Range exitR = slice(g, g.N[0], 0, 2);  // slice(N.-1, N[1], 1, 2)
  Range inR = slice(g, 2, 0, 2);
  for (int d = 1; d < g.D; d++) exitR.hi[d] = inR.hi[d] = g.N[d] - 1;
  const T len = (T)exitR.count();
  T U = (T)sum_range(g, u, inR) / len;
  loop(exitR, [&](I3 I) {
    size_t o = g.at(I);
    u[o] = u0[o] - U * dt * (u0[o] - u0[g.at(shift(I, 0, -1))]);
  });
  T flux = (T)sum_range(g, u, exitR) / len - U;
  loop(exitR, [&](I3 I) { u[g.at(I)] -= flux; });
}

// perBC!  src/core.jl:239-243
static void perBC(const Grid& g, T* a, const int* per) {
  for (int j = 0; j < g.D; j++) {
    if (!per[j]) continue;
    const int Nj = g.N[j];
    loop(slice(g, 1, j), [&](I3 I) { a[g.at(I)] = a[g.at(setj(I, j, Nj - 1))]; });
    loop(slice(g, Nj, j), [&](I3 I) { a[g.at(I)] = a[g.at(setj(I, j, 2))]; });
  }
}

// ---- Poisson  src/Poisson.jl ---------------------------------------------------------
struct Poisson {
  Grid g;
  T *L, *x, *z;  // may alias flow arrays (level 1) or own storage
  std::vector<T> Lown, xown, zown, D, iD, eps, r;
  int per[3];
  std::vector<int16_t> n;
  size_t nn() const { return g.n(); }
  const T* Lc(int i) const { return L + (size_t)i * g.n(); }
};

static void set_diag(Poisson& p) {  // :43-55
  const Grid& g = p.g;
  loop(inside(g), [&](I3 I) {
    T s = 0;
    for (int i = 0; i < g.D; i++) s -= p.Lc(i)[g.at(I)] + p.Lc(i)[g.at(shift(I, i, 1))];
    p.D[g.at(I)] = s;
  });
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    p.iD[o] = (p.D[o] == 0) ? p.D[o] : 1 / p.D[o];
  });
}
static void pois_init(Poisson& p, const Grid& g, T* x, T* L, T* z, const int* per) {  // :32-38
  p.g = g;
  p.x = x;
  p.L = L;
  p.z = z;
  for (int d = 0; d < 3; d++) p.per[d] = per[d];
  p.r.assign(g.n(), 0);
  p.eps.assign(g.n(), 0);
  p.D.assign(g.n(), 0);
  p.iD.assign(g.n(), 0);
  set_diag(p);
}
static inline T mult_at(const Poisson& p, const I3& I, const T* x) {  // mult :70-76
  const Grid& g = p.g;
  size_t o = g.at(I);
  T s = x[o] * p.D[o];
  for (int i = 0; i < g.D; i++) s += x[g.at(shift(I, i, -1))] * p.Lc(i)[o] + x[g.at(shift(I, i, 1))] * p.Lc(i)[g.at(shift(I, i, 1))];
  return s;
}
static void mult(Poisson& p, T* x) {  // mult! :63-69
  perBC(p.g, x, p.per);
  std::fill(p.z, p.z + p.nn(), (T)0);
  loop(inside(p.g), [&](I3 I) { p.z[p.g.at(I)] = mult_at(p, I, x); });
}
static void residual(Poisson& p) {  // residual! :92-98
  const Grid& g = p.g;
  perBC(g, p.x, p.per);
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    p.r[o] = (p.iD[o] == 0) ? (T)0 : p.z[o] - mult_at(p, I, p.x);
  });
  T s = (T)sum_range(g, p.r.data(), all_cells(g)) / (T)inside(g).count();
  if (std::fabs(s) <= 2 * 1.1920929e-7f) return;
  loop(inside(g), [&](I3 I) { p.r[g.at(I)] = p.r[g.at(I)] - s; });
}
static void increment(Poisson& p, T w) {  // increment! :100-104
  const Grid& g = p.g;
  perBC(g, p.eps.data(), p.per);
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    p.r[o] = p.r[o] - w * mult_at(p, I, p.eps.data());
    p.x[o] = p.x[o] + w * p.eps[o];
  });
}
static void Jacobi(Poisson& p, int it = 1, T w = 1) {  // Jacobi! :111-114
  for (int k = 0; k < it; k++) {
    loop(inside(p.g), [&](I3 I) {
      size_t o = p.g.at(I);
      p.eps[o] = p.r[o] * p.iD[o];
    });
    increment(p, w);
  }
}
static void GaussSeidelRB(Poisson& p, int it, T w) {  // :116-148
  const Grid& g = p.g;
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    p.eps[o] = p.r[o] * p.iD[o];
  });
  perBC(g, p.eps.data(), p.per);
  const int d = g.D - 1;  // last spatial dim
  Range H = inside(g);    // half_rangek :130-132
  H.lo[d] = 2;
  H.hi[d] = g.N[d] / 2;
  for (int k0 = 1; k0 <= it; k0++) {
    loop(H, [&](I3 Iv) {
      int front = 0;
      for (int q = 0; q < d; q++) front += Iv.v[q];
      int k = 2 * Iv.v[d] - 1 - (front + k0) % 2;  // gauss_rb :125
      I3 I = setj(Iv, d, k);
      size_t o = g.at(I);
      T s = p.r[o];  // gauss :116-122
      for (int i = 0; i < g.D; i++) s -= p.eps[g.at(shift(I, i, -1))] * p.Lc(i)[o] + p.eps[g.at(shift(I, i, 1))] * p.Lc(i)[g.at(shift(I, i, 1))];
      p.eps[o] = s * p.iD[o];
    });
  }
  increment(p, w);
}
static T L2(const Poisson& p) { return (T)dot_range(p.g, p.r.data(), p.r.data(), all_cells(p.g)); }  // :189
static T Linf(const Poisson& p) {
  T m = 0;
  for (size_t o = 0; o < p.nn(); o++) m = std::max(m, std::fabs(p.r[o]));
  return m;
}
static T perdot(const Poisson& p, const T* a, const T* b) {  // :156-157
  bool anyper = false;
  for (int d = 0; d < p.g.D; d++) anyper |= (p.per[d] != 0);
  return (T)dot_range(p.g, a, b, anyper ? inside(p.g) : all_cells(p.g));
}
static void pcg(Poisson& p, int it = 6) {  // pcg! :166-186
  const Grid& g = p.g;
  const T eps32 = 1.1920929e-7f;
  T *x = p.x, *r = p.r.data(), *e = p.eps.data(), *z = p.z;
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    z[o] = e[o] = r[o] * p.iD[o];
  });
  T rho = (T)dot_range(g, r, z, all_cells(g));
  if (std::fabs(rho) < 10 * eps32) return;
  for (int i = 1; i <= it; i++) {
    perBC(g, e, p.per);
    loop(inside(g), [&](I3 I) { z[g.at(I)] = mult_at(p, I, e); });
    T alpha = rho / perdot(p, z, e);
    if (std::fabs(alpha) < 1e-2 || std::fabs(alpha) > 1e2) return;
    loop(inside(g), [&](I3 I) {
      size_t o = g.at(I);
      x[o] += alpha * e[o];
      r[o] -= alpha * z[o];
    });
    if (i == it) return;
    loop(inside(g), [&](I3 I) {
      size_t o = g.at(I);
      z[o] = r[o] * p.iD[o];
    });
    T rho2 = (T)dot_range(g, r, z, all_cells(g));
    if (std::fabs(rho2) < 10 * eps32) return;
    T beta = rho2 / rho;
    loop(inside(g), [&](I3 I) {
      size_t o = g.at(I);
      e[o] = beta * e[o] + z[o];
    });
    rho = rho2;
  }
}

struct SolverLog {
  std::vector<float> rows;  // (iter, r2, omega) per line of the reference's @log
  void add(int it, T r2, T w) {
    rows.push_back((float)it);
    rows.push_back(r2);
    rows.push_back(w);
  }
};

static int solver_single(Poisson& p, double tol, double itmx, SolverLog* log) {  // solver!(::Poisson) :204-214
  residual(p);
  T r2 = L2(p);
  int np = 0;
  if (log) log->add(np, r2, 1);
  while (np < itmx) {
    pcg(p);
    r2 = L2(p);
    np++;
    if (log) log->add(np, r2, 1);
    if ((double)r2 < tol) break;
  }
  perBC(p.g, p.x, p.per);
  p.n.push_back((int16_t)np);
  return np;
}

// ---- MultiLevelPoisson  src/MultiLevelPoisson.jl -------------------------------------
static inline bool divisible(int N) { return N % 2 == 0 && N > 4; }  // :52
struct MLPoisson {
  std::vector<Poisson*> levels;
  std::vector<int16_t> n;
  int per[3];
  int smoother = 0;  // 0 = GaussSeidelRB! (default :106), 1 = pcg!
  ~MLPoisson() {
    for (auto* l : levels) delete l;
  }
};
static void coarsen_mask(const Grid& fine, const Grid& coarse, bool* c) {  // :31
  for (int d = 0; d < 3; d++) c[d] = d < fine.D && coarse.N[d] < fine.N[d];
}
// restrictL!  :42-48 (+ restrictL :20-26, upL :9-11)
static void restrictL(Poisson& a, const Poisson& b, const bool* c) {
  const Grid &ga = a.g, &gb = b.g;
  const int D = ga.D;
  for (int i = 0; i < D; i++) {
    T* ai = a.L + (size_t)i * ga.n();
    const T* bi = b.Lc(i);
    loop(inside(ga), [&](I3 I) {
      int lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
      for (int j = 0; j < D; j++) {
        if (j == i) {
          lo[j] = hi[j] = c[i] ? 2 * I.v[j] - 2 : I.v[j];
        } else {
          lo[j] = c[j] ? 2 * I.v[j] - 2 : I.v[j];
          hi[j] = c[j] ? 2 * I.v[j] - 1 : I.v[j];
        }
      }
      T s = 0;
      for (int k = lo[2]; k <= hi[2]; k++)
        for (int j = lo[1]; j <= hi[1]; j++)
          for (int ii = lo[0]; ii <= hi[0]; ii++) s += bi[gb.at(I3{{ii, j, k}})];
      ai[ga.at(I)] = c[i] ? s / 2 : s;
    });
  }
  T zero[3] = {0, 0, 0};
  BC(ga, a.L, zero, false, a.per);  // BC!(a,0,false,perdir)  :47
}
static Poisson* restrictML(const Poisson& b) {  // :33-41
  Poisson* a = new Poisson();
  Grid g = b.g;
  for (int d = 0; d < g.D; d++)
    if (divisible(b.g.N[d])) g.N[d] = 1 + b.g.N[d] / 2;
  bool c[3];
  coarsen_mask(b.g, g, c);
  a->Lown.assign(g.n() * g.D, 0);
  a->xown.assign(g.n(), 0);
  a->zown.assign(g.n(), 0);
  a->g = g;
  a->L = a->Lown.data();
  for (int d = 0; d < 3; d++) a->per[d] = b.per[d];
  restrictL(*a, b, c);
  pois_init(*a, g, a->xown.data(), a->Lown.data(), a->zown.data(), b.per);
  return a;
}
static bool level_divisible(const Poisson& l) {  // :54
  for (int d = 0; d < l.g.D; d++)
    if (divisible(l.g.N[d])) return true;
  return false;
}
static MLPoisson* ml_create(const Grid& g, T* x, T* L, T* z, const int* per, int maxlevels = 10) {  // :68-76
  MLPoisson* ml = new MLPoisson();
  for (int d = 0; d < 3; d++) ml->per[d] = per[d];
  Poisson* p = new Poisson();
  pois_init(*p, g, x, L, z, per);
  ml->levels.push_back(p);
  while (level_divisible(*ml->levels.back()) && (int)ml->levels.size() <= maxlevels) ml->levels.push_back(restrictML(*ml->levels.back()));
  if (ml->levels.size() <= 2) {  // @assert length(levels)>2
    delete ml;
    return nullptr;
  }
  return ml;
}
static void ml_update(MLPoisson& ml) {  // update! :79-86
  set_diag(*ml.levels[0]);
  for (size_t l = 1; l < ml.levels.size(); l++) {
    bool c[3];
    coarsen_mask(ml.levels[l - 1]->g, ml.levels[l]->g, c);
    restrictL(*ml.levels[l], *ml.levels[l - 1], c);
    set_diag(*ml.levels[l]);
  }
}
static void restrict_r(Poisson& coarse, const Poisson& fine, const bool* c) {  // restrict! :49 (+ :13-19, up :6)
  const Grid &ga = coarse.g, &gb = fine.g;
  loop(inside(ga), [&](I3 I) {
    int lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
    for (int j = 0; j < ga.D; j++) {
      lo[j] = c[j] ? 2 * I.v[j] - 2 : I.v[j];
      hi[j] = c[j] ? 2 * I.v[j] - 1 : I.v[j];
    }
    T s = 0;
    for (int k = lo[2]; k <= hi[2]; k++)
      for (int j = lo[1]; j <= hi[1]; j++)
        for (int ii = lo[0]; ii <= hi[0]; ii++) s += fine.r[gb.at(I3{{ii, j, k}})];
    coarse.r[ga.at(I)] = s;
  });
}
static void prolongate(Poisson& fine, const Poisson& coarse, const bool* c) {  // prolongate! :50 (+ down :7)
  const Grid &gf = fine.g, &gc = coarse.g;
  loop(inside(gf), [&](I3 I) {
    I3 J = I;
    for (int j = 0; j < gf.D; j++) J.v[j] = c[j] ? (I.v[j] + 2) / 2 : I.v[j];
    fine.eps[gf.at(I)] = coarse.x[gc.at(J)];
  });
}
static void smooth(MLPoisson& ml, Poisson& p, T w) {  // smooth! :106
  if (ml.smoother == 0)
    GaussSeidelRB(p, 4, w);
  else
    pcg(p);
}
static void Vcycle(MLPoisson& ml, size_t l, T w) {  // :88-101  (l is 0-based here)
  Poisson &fine = *ml.levels[l], &coarse = *ml.levels[l + 1];
  bool c[3];
  coarsen_mask(fine.g, coarse.g, c);
  Jacobi(fine);
  restrict_r(coarse, fine, c);
  std::fill(coarse.x, coarse.x + coarse.nn(), (T)0);
  if (l + 2 < ml.levels.size()) Vcycle(ml, l + 1, w);
  smooth(ml, coarse, w);
  prolongate(fine, coarse, c);
  increment(fine, w);
}
static int solver_ml(MLPoisson& ml, double tol, int itmx, SolverLog* log) {  // solver!(ml) :108-127
  Poisson& p = *ml.levels[0];
  residual(p);
  T r2 = L2(p);
  T w = 1;
  int np = 0;
  if (log) log->add(np, r2, w);
  while (np < itmx) {
    Vcycle(ml, 0, w);
    smooth(ml, p, w);
    T rnew = L2(p);
    np++;
    if (log) log->add(np, rnew, w);
    if (rnew >= r2)
      w = (T)std::max(0.2, 0.9 * (double)w);
    else if (rnew < r2)
      w = (T)std::min(1.0, 1.02 * (double)w);
    r2 = rnew;
    if ((double)r2 < tol) break;
  }
  perBC(p.g, p.x, p.per);
  ml.n.push_back((int16_t)np);
  return np;
}

// ---- Flow  src/Flow.jl ----------------------------------------------------------------
struct Flow {
  Grid g;
  std::vector<T> u, u0, f, p, sigma, V, mu0, mu1;
  T uBC[3];
  std::vector<T> dt;
  T nu;
  int exit;
  int per[3];
  int lam;
  MLPoisson* ml = nullptr;
  Poisson* single = nullptr;  // if the pressure solver is a single-level Poisson
  SolverLog log;
  double tol = 1e-4;
  int itmx = 32;
  // enumerated forcings standing in for the closures g(i,x,t) and uBC(i,x,t) (src/Flow.jl:64-73, src/core.jl:201-219):
  // g_i(t) = g0_i + g1_i·t;  U_i(t) = uBC_i + U1_i·t + ½·U2_i·t²  (uniform in space)
  bool forcing = false;
  T g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0}, U1[3] = {0, 0, 0}, U2[3] = {0, 0, 0};
  T ubc_t[3];  // uBC evaluated at the time the current BC! call is made at
  // built-in udf: sgs! with the Smagorinsky–Lilly νₜ of its docstring (src/util.jl:57-76); S is the user's buffer (N…,D,D), zeros outside inside(σ)
  bool sgs = false;
  T sgs_c2 = 0;  // (Cs·Δ)²
  std::vector<T> S;
  ~Flow() {
    delete ml;
    delete single;
  }
};

static void scale_u(Flow& a, T s) {  // :211-214
  const Grid& g = a.g;
  for (int i = 0; i < g.D; i++) {
    T* ui = a.u.data() + (size_t)i * g.n();
    loop(inside(g), [&](I3 I) { ui[g.at(I)] *= s; });
  }
}
static void BDIM(Flow& a) {  // :176-180
  const Grid& g = a.g;
  const size_t n = g.n();
  const int D = g.D;
  const T dt = a.dt.back();
#pragma omp parallel for schedule(static)
  for (size_t o = 0; o < n * D; o++) a.f[o] = a.u0[o] + dt * a.f[o] - a.V[o];
  for (int i = 0; i < D; i++) {
    T* ui = a.u.data() + (size_t)i * n;
    const T* fi = a.f.data() + (size_t)i * n;
    loop(inside(g), [&](I3 I) {
      size_t o = g.at(I);
      T s = 0;  // μddn :20-26
      for (int j = 0; j < D; j++) s += a.mu1[o + n * ((size_t)i + (size_t)D * j)] * (fi[g.at(shift(I, j, 1))] - fi[g.at(shift(I, j, -1))]);
      ui[o] += s / 2 + a.V[o + n * i] + a.mu0[o + n * i] * fi[o];
    });
  }
}
static int flow_solver(Flow& a) {
  if (a.ml) return solver_ml(*a.ml, a.tol, a.itmx, &a.log);
  return solver_single(*a.single, a.tol, 1e3, &a.log);
}
static void project(Flow& a, T w) {  // mom_project! :223-232
  const Grid& g = a.g;
  const size_t n = g.n();
  const T dt = w * a.dt.back();
  Poisson& b = a.ml ? *a.ml->levels[0] : *a.single;
  loop(inside(g), [&](I3 I) {
    T s = 0;  // div :13-19
    for (int i = 0; i < g.D; i++) s += a.u[g.at(shift(I, i, 1)) + n * i] - a.u[g.at(I) + n * i];
    b.z[g.at(I)] = s;
  });
#pragma omp parallel for schedule(static)
  for (size_t o = 0; o < n; o++) b.x[o] *= dt;
  flow_solver(a);
  for (int i = 0; i < g.D; i++) {
    T* ui = a.u.data() + (size_t)i * n;
    loop(inside(g), [&](I3 I) {
      size_t o = g.at(I);
      ui[o] -= b.Lc(i)[o] * (b.x[o] - b.x[g.at(shift(I, i, -1))]);
    });
  }
#pragma omp parallel for schedule(static)
  for (size_t o = 0; o < n; o++) b.x[o] /= dt;
  BC(g, a.u.data(), a.forcing ? a.ubc_t : a.uBC, a.exit, a.per);
}
static T CFL(Flow& a) {  // :234-244
  const Grid& g = a.g;
  const size_t n = g.n();
  loop(inside(g), [&](I3 I) {
    T s = 0;
    for (int i = 0; i < g.D; i++) s += std::max((T)0, a.u[g.at(shift(I, i, 1)) + n * i]) + std::max((T)0, -a.u[g.at(I) + n * i]);
    a.sigma[g.at(I)] = s;
  });
  T m = a.sigma[0];  // maximum(a.σ) over ALL cells (ghosts hold stale Φ)
#pragma omp parallel for reduction(max : m) schedule(static)
  for (size_t o = 0; o < n; o++) m = std::max(m, a.sigma[o]);
  return std::min((T)10, 1 / (m + 5 * a.nu));
}
// accelerate!(r,t,g,U): r[Ii] += g(i,x,t) + dU(i,x,t)/dt over ALL cells of r  (src/Flow.jl:64-73)
static void accelerate(Flow& a, T t) {
  if (!a.forcing) return;
  const size_t n = a.g.n();
  for (int i = 0; i < a.g.D; i++) {
    const T acc = (a.g0[i] + a.g1[i] * t) + (a.U1[i] + a.U2[i] * t);
    T* fi = a.f.data() + (size_t)i * n;
#pragma omp parallel for schedule(static)
    for (size_t o = 0; o < n; o++) fi[o] += acc;
  }
}
// ∂(i,j,I,u): ∂uᵢ/∂xⱼ at the centre of cell I  (src/Metrics.jl:42-44)
static inline T dudx_c(const Grid& g, const T* u, int i, int j, const I3& I) {
  const size_t n = g.n();
  const T* ui = u + n * i;
  if (i == j) return ui[g.at(shift(I, i, 1))] - ui[g.at(I)];
  const I3 P = shift(I, j, 1), M = shift(I, j, -1);
  return (ui[g.at(P)] + ui[g.at(shift(P, i, 1))] - ui[g.at(M)] - ui[g.at(shift(M, i, 1))]) / 4;
}
// sgs!(flow,u,t; νₜ,S,Cs,Δ)  src/util.jl:66-76, with νₜ = smagorinsky(I;S,Cs,Δ) = (Cs·Δ)²·sqrt(dot(S[I,:,:],S[I,:,:]))  (src/util.jl:62)
// and S(I,u) = (∂(i,j,I,u)+∂(j,i,I,u))/2  (src/Metrics.jl:140).  `u` is the advecting field of the phase (u⁰ / u, src/Flow.jl:192,207).
static void sgs(Flow& a, const T* u) {
  const Grid& g = a.g;
  const size_t n = g.n();
  const int D = g.D;
  loop(inside(g), [&](I3 I) {
    const size_t o = g.at(I);
    for (int j = 0; j < D; j++)
      for (int i = 0; i < D; i++) a.S[o + n * ((size_t)i + (size_t)D * j)] = (dudx_c(g, u, i, j, I) + dudx_c(g, u, j, i, I)) / 2;
  });
  auto nut = [&](size_t o) -> T {  // generic `dot` of the two views: sequential, column-major
    T s = 0;
    for (int q = 0; q < D * D; q++) s += a.S[o + n * q] * a.S[o + n * q];
    return a.sgs_c2 * std::sqrt(s);
  };
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) {
      const T* ui = u + n * i;
      T* fi = a.f.data() + n * i;
      loop(inside_u(g, j), [&](I3 I) {
        const size_t o = g.at(I);
        a.sigma[o] = -nut(o) * (ui[o] - ui[g.at(shift(I, j, -1))]);
        fi[o] += a.sigma[o];
      });
      loop(inside_u(g, j), [&](I3 I) { fi[g.at(shift(I, j, -1))] -= a.sigma[g.at(I)]; });
    }
}
static void mom_step(Flow& a) {  // mom_step! :156-167
  const Grid& g = a.g;
  a.u0 = a.u;
  scale_u(a, 0);
  T t1 = 0, t0 = 0;
  if (a.forcing) {  // t₁ = sum(a.Δt); t₀ = t₁ − a.Δt[end]  (the Float32 nearest to the exact sum, see oracle.py time())
    double s = 0;
    for (T v : a.dt) s += (double)v;
    t1 = (T)s;
    t0 = t1 - a.dt.back();
    for (int i = 0; i < 3; i++) a.ubc_t[i] = a.uBC[i] + a.U1[i] * t1 + (a.U2[i] * t1) * t1 / 2;  // BC MUST be at t₁ (src/Flow.jl:194)
  }
  // predictor  :190-196
  conv_diff(g, a.f.data(), a.u0.data(), a.sigma.data(), a.lam, a.nu, a.per);
  if (a.sgs) sgs(a, a.u0.data());  // udf!(a,udf,a.u⁰,t₀)  src/Flow.jl:192
  accelerate(a, t0);
  BDIM(a);
  BC(g, a.u.data(), a.forcing ? a.ubc_t : a.uBC, a.exit, a.per);
  if (a.exit) exitBC(g, a.u.data(), a.u0.data(), a.dt.back());
  project(a, 1);
  // corrector  :205-210
  conv_diff(g, a.f.data(), a.u.data(), a.sigma.data(), a.lam, a.nu, a.per);
  if (a.sgs) sgs(a, a.u.data());  // udf!(a,udf,a.u,t)  src/Flow.jl:207
  accelerate(a, t1);
  BDIM(a);
  scale_u(a, 0.5f);
  BC(g, a.u.data(), a.forcing ? a.ubc_t : a.uBC, a.exit, a.per);
  project(a, 0.5f);
  a.dt.push_back(CFL(a));
}

// ---- Body kernels + analytic-SDF measure!  src/Body.jl:28-60, src/AutoBody.jl:29-37 ----
static inline T ulp(T d) {  // Julia eps(d::Float32)
  T a = std::fabs(d);
  return std::nextafter(a, INFINITY) - a;
}
// sinpi / cospi (src/Body.jl:55-56 call Julia's sinpi/cospi, which are exact at multiples of ½: sinpi(1) = 0, cospi(½) = 0 — the
// kernel moments vanish EXACTLY at |d| = ϵ).  Evaluated in double with the argument reduced to [−½, ½] first (exact for Float32
// inputs), then rounded; valid for |x| ≤ 1.5, the kernels call them on [−1, 1].
static inline double sinpi_d(double x) {
  const double a = std::fabs(x);
  const double r = a <= 0.5 ? std::sin(M_PI * a) : std::sin(M_PI * (1.0 - a));
  return x < 0 ? -r : r;
}
static inline T sinpiT(T x) { return (T)sinpi_d((double)x); }
static inline T cospiT(T x) { return (T)sinpi_d(0.5 - std::fabs((double)x)); }
static inline T kern0(T d) { return (1 + d + sinpiT(d) / (T)M_PI) / 2; }                                               // :55
static inline T kern1(T d) { return (1 - d * d) / 4 - (d * sinpiT(d) + (1 + cospiT(d)) / (T)M_PI) / (2 * (T)M_PI); }  // :56
static inline T mu0f(T d, T e) { return d / e < -1 + std::sqrt(ulp(d)) ? (T)0 : kern0(std::min(d / e, (T)1)); }       // :59
static inline T mu1f(T d, T e) { return e * kern1(std::max((T)-1, std::min(d / e, (T)1))); }                          // :60

struct Body {
  int kind;  // 0 sphere/circle, 1 torus (axis ∥ x)
  T c[3];
  T R, r;
};
static T body_sdf(const Body& b, int D, const T* x) {
  if (b.kind == 0) {
    T s = 0;
    for (int d = 0; d < D; d++) s += (x[d] - b.c[d]) * (x[d] - b.c[d]);
    return std::sqrt(s) - b.R;
  }
  T y = x[1] - b.c[1], z = x[2] - b.c[2], xx = x[0] - b.c[0];
  T q = std::sqrt(y * y + z * z) - b.R;
  return std::sqrt(q * q + xx * xx) - b.r;
}
static void body_grad(const Body& b, int D, const T* x, T* n) {  // what ForwardDiff returns for these sdfs
  if (b.kind == 0) {
    T s = 0;
    for (int d = 0; d < D; d++) s += (x[d] - b.c[d]) * (x[d] - b.c[d]);
    T m = std::sqrt(s);
    for (int d = 0; d < D; d++) n[d] = (x[d] - b.c[d]) / m;
    return;
  }
  T y = x[1] - b.c[1], z = x[2] - b.c[2], xx = x[0] - b.c[0];
  T rho = std::sqrt(y * y + z * z);
  T q = rho - b.R;
  T m = std::sqrt(q * q + xx * xx);
  n[0] = xx / m;
  n[1] = (q / m) * (y / rho);
  n[2] = (q / m) * (z / rho);
}
// measure(body::AutoBody,x,t;fastd²)  src/AutoBody.jl:29-37  (static map ⇒ J=I, V=0)
static void body_measure(const Body& b, int D, const T* x, T fastd2, T* d, T* n) {
  *d = body_sdf(b, D, x);
  for (int k = 0; k < 3; k++) n[k] = 0;
  if ((*d) * (*d) > fastd2) return;
  T gk[3] = {0, 0, 0};
  body_grad(b, D, x, gk);
  for (int k = 0; k < D; k++)
    if (std::isnan(gk[k])) return;
  T m = 0;
  for (int k = 0; k < D; k++) m += gk[k] * gk[k];
  m = std::sqrt(m);
  *d /= m;
  for (int k = 0; k < D; k++) n[k] = gk[k] / m;
}
// ---- parametrised bodies with a rigid translation map and lazy set operations (src/AutoBody.jl:10-37, src/Body.jl:88-107) ----
// Primitive k is AutoBody(sdf_k, (x,t) -> x .- vel_k.*t); the body is ((p₀ op₁ p₁) op₂ p₂) … with op ∈ {∪ = min, ∩ = max, − = a ∩ (−b)}
// on the measure tuples (d, n, V) (tuple isless: lexicographic).
struct Prim {
  int kind, op;  // kind: 0 sphere/circle, 1 torus (axis ∥ x); op: 0 ∪, 1 ∩, 2 − (ignored for the first primitive)
  T c[3];
  T R, r;
  T vel[3];
};
struct Meas {
  T d, n[3], V[3];
};
static Meas prim_measure(const Prim& p, int D, const T* x, T t, T fastd2) {  // measure(body::AutoBody,x,t;fastd²)  src/AutoBody.jl:29-37
  Body b{p.kind, {p.c[0], p.c[1], p.c[2]}, p.R, p.r};
  T xi[3] = {0, 0, 0};
  for (int k = 0; k < D; k++) xi[k] = x[k] - p.vel[k] * t;  // map(x,t)
  Meas m;
  m.d = body_sdf(b, D, xi);
  for (int k = 0; k < 3; k++) m.n[k] = m.V[k] = 0;
  if (m.d * m.d > fastd2) return m;
  T gk[3] = {0, 0, 0};
  body_grad(b, D, xi, gk);
  for (int k = 0; k < D; k++)
    if (std::isnan(gk[k])) return m;
  T mm = 0;  // J = I: n = J'n
  for (int k = 0; k < D; k++) mm += gk[k] * gk[k];
  mm = std::sqrt(mm);
  m.d /= mm;
  for (int k = 0; k < D; k++) {
    m.n[k] = gk[k] / mm;
    m.V[k] = p.vel[k];  // −J\∂ₜmap = vel
  }
  return m;
}
static bool fless(T a, T b) {  // isless(::Float32, ::Float32): NaN is the largest, −0.0 < 0.0
  if (std::isnan(a)) return false;
  if (std::isnan(b)) return true;
  if (a == b) return std::signbit(a) && !std::signbit(b);
  return a < b;
}
static bool meas_less(const Meas& a, const Meas& b, int D) {  // isless on the tuples (d, n, V)
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
static Meas csg_measure(const Prim* ps, int np, int D, const T* x, T t, T fastd2) {  // measure(body::SetBody, …)  src/Body.jl:104-107
  Meas acc = prim_measure(ps[0], D, x, t, fastd2);
  for (int q = 1; q < np; q++) {
    Meas m = prim_measure(ps[q], D, x, t, fastd2);
    if (ps[q].op == 2) {  // a − b = a ∩ (−b): (−d, −n, V)
      m.d = -m.d;
      for (int k = 0; k < D; k++) m.n[k] = -m.n[k];
    }
    if (ps[q].op == 0) {  // min(x,y) = ifelse(isless(y,x), y, x)
      if (meas_less(m, acc, D)) acc = m;
    } else {  // max(x,y) = ifelse(isless(y,x), x, y)
      if (!meas_less(m, acc, D)) acc = m;
    }
  }
  return acc;
}
// ---- forces and moments on the body (src/Metrics.jl:111-190) ----
// nds(body,x,t) = n·kern(clamp(d,−1,1)) with (d,n) = measure(body,x,t,fastd²=1), kern(d) = (1+cospi(d))/2.
// out[0:3] pressure_force, out[3:6] viscous_force, out[6:9] pressure_moment(x₀), out[9:12] viscous_moment(x₀); Float64 sums of the
// Float32 per-cell vectors df[I,:] like sum(Float64, df, dims=…).  2-D moments: cross of 2-vectors is the scalar a₁b₂−a₂b₁,
// broadcast into both components of df.
static void body_forces(Flow& a, const Prim* ps, int np, T t, const T* x0, double* out) {
  const Grid& g = a.g;
  const size_t n = g.n();
  const int D = g.D;
  for (int q = 0; q < 12; q++) out[q] = 0;
  const T* u = a.u.data();
  const T nu2 = -2 * a.nu;
  auto U = [&](I3 I, int i) { return u[g.at(I) + n * i]; };
  auto dd = [&](int i, int j, I3 I) -> T {  // ∂(i,j,I,u)  src/Metrics.jl:42-44
    if (i == j) return U(shift(I, i), i) - U(I, i);
    I3 P = shift(I, j), M = shift(I, j, -1);
    return (U(P, i) + U(shift(P, i), i) - U(M, i) - U(shift(M, i), i)) / 4;
  };
  double acc[12] = {0};
  Range R = inside(g);
  for (int k = R.lo[2]; k <= R.hi[2]; k++)
    for (int j = R.lo[1]; j <= R.hi[1]; j++)
      for (int i = R.lo[0]; i <= R.hi[0]; i++) {
        I3 I{{i, j, k}};
        T x[3] = {0, 0, 0};
        for (int d = 0; d < D; d++) x[d] = (T)I.v[d] - 1.5f;
        Meas m = csg_measure(ps, np, D, x, t, (T)1);
        T kd = (1 + cospiT(std::max((T)-1, std::min(m.d, (T)1)))) / 2;
        T nds[3] = {0, 0, 0};
        for (int d = 0; d < D; d++) nds[d] = m.n[d] * kd;
        T S[3][3] = {{0}}, Sn[3] = {0, 0, 0}, fv[3] = {0, 0, 0}, r[3] = {0, 0, 0};
        for (int p = 0; p < D; p++)
          for (int q = 0; q < D; q++) S[p][q] = (dd(p, q, I) + dd(q, p, I)) / 2;
        for (int p = 0; p < D; p++) {
          T s1 = 0, s2 = 0;
          for (int q = 0; q < D; q++) {
            s1 = q == 0 ? S[p][q] * nds[q] : s1 + S[p][q] * nds[q];
            s2 = q == 0 ? (nu2 * S[p][q]) * nds[q] : s2 + (nu2 * S[p][q]) * nds[q];
          }
          Sn[p] = s1;
          fv[p] = s2;
        }
        for (int d = 0; d < D; d++) r[d] = x[d] - x0[d];
        T pr = a.p[g.at(I)];
        T pm[3], vm[3];
        if (D == 3) {
          T c1[3] = {r[1] * nds[2] - r[2] * nds[1], r[2] * nds[0] - r[0] * nds[2], r[0] * nds[1] - r[1] * nds[0]};
          T c2[3] = {r[1] * Sn[2] - r[2] * Sn[1], r[2] * Sn[0] - r[0] * Sn[2], r[0] * Sn[1] - r[1] * Sn[0]};
          for (int d = 0; d < 3; d++) pm[d] = pr * c1[d], vm[d] = nu2 * c2[d];
        } else {
          T c1 = r[0] * nds[1] - r[1] * nds[0], c2 = r[0] * Sn[1] - r[1] * Sn[0];
          pm[0] = pm[1] = pr * c1;
          vm[0] = vm[1] = nu2 * c2;
          pm[2] = vm[2] = 0;
        }
        for (int d = 0; d < D; d++) {
          acc[d] += (double)(T)(pr * nds[d]);
          acc[3 + d] += (double)fv[d];
          acc[6 + d] += (double)pm[d];
          acc[9 + d] += (double)vm[d];
        }
      }
  for (int q = 0; q < 12; q++) out[q] = acc[q];
}
static void measure_prims(Flow& a, const Prim* ps, int np, T eps, T t) {  // measure!(flow, body; t, ϵ)  src/Body.jl:28-51
  const Grid& g = a.g;
  const size_t n = g.n();
  const int D = g.D;
  std::fill(a.V.begin(), a.V.end(), (T)0);
  std::fill(a.mu0.begin(), a.mu0.end(), (T)1);
  std::fill(a.mu1.begin(), a.mu1.end(), (T)0);
  const T d2 = (2 + eps) * (2 + eps);
  loop(inside(g), [&](I3 I) {  // measure_sdf!(σ, body, t; fastd²=d²): AutoBody → its sdf (src/AutoBody.jl:19); SetBody → measure(…)[1] (src/Body.jl:67)
    T x[3];
    for (int d = 0; d < 3; d++) x[d] = (T)I.v[d] - 1.5f;
    if (np == 1) {
      Body b{ps[0].kind, {ps[0].c[0], ps[0].c[1], ps[0].c[2]}, ps[0].R, ps[0].r};
      T xi[3] = {0, 0, 0};
      for (int k = 0; k < D; k++) xi[k] = x[k] - ps[0].vel[k] * t;
      a.sigma[g.at(I)] = body_sdf(b, D, xi);
    } else
      a.sigma[g.at(I)] = csg_measure(ps, np, D, x, t, d2).d;
  });
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    T dI = a.sigma[o];
    if (dI * dI < d2) {
      for (int i = 0; i < D; i++) {
        T x[3];
        for (int d = 0; d < 3; d++) x[d] = (T)I.v[d] - 1.5f - (d == i ? 0.5f : 0.f);  // loc(i,I)  src/core.jl:177
        Meas m = csg_measure(ps, np, D, x, t, d2);
        T di = std::fabs(m.d) <= 0.5f ? m.d : std::copysign(m.d, dI);
        a.V[o + n * i] = m.V[i];
        a.mu0[o + n * i] = mu0f(di, eps);
        for (int j = 0; j < D; j++) a.mu1[o + n * ((size_t)i + (size_t)D * j)] = mu1f(di, eps) * m.n[j];
      }
    } else if (dI < 0) {
      for (int i = 0; i < D; i++) a.mu0[o + n * i] = 0;
    }
  });
  T zero[3] = {0, 0, 0};
  BC(g, a.mu0.data(), zero, false, a.per);
  BC(g, a.V.data(), zero, a.exit, a.per);
}
static void measure(Flow& a, const Body& b, T eps) {  // measure!  src/Body.jl:28-51
  const Grid& g = a.g;
  const size_t n = g.n();
  const int D = g.D;
  std::fill(a.V.begin(), a.V.end(), (T)0);
  std::fill(a.mu0.begin(), a.mu0.end(), (T)1);
  std::fill(a.mu1.begin(), a.mu1.end(), (T)0);
  const T d2 = (2 + eps) * (2 + eps);
  loop(inside(g), [&](I3 I) {  // measure_sdf! :74
    T x[3];
    for (int d = 0; d < 3; d++) x[d] = (T)I.v[d] - 1.5f;
    a.sigma[g.at(I)] = body_sdf(b, D, x);
  });
  loop(inside(g), [&](I3 I) {
    size_t o = g.at(I);
    T dI = a.sigma[o];
    if (dI * dI < d2) {
      for (int i = 0; i < D; i++) {
        T x[3];
        for (int d = 0; d < 3; d++) x[d] = (T)I.v[d] - 1.5f - (d == i ? 0.5f : 0.f);  // loc(i,I)  src/core.jl:177
        T di, ni[3];
        body_measure(b, D, x, d2, &di, ni);
        di = std::fabs(di) <= 0.5f ? di : std::copysign(di, dI);
        a.V[o + n * i] = 0;
        a.mu0[o + n * i] = mu0f(di, eps);
        for (int j = 0; j < D; j++) a.mu1[o + n * ((size_t)i + (size_t)D * j)] = mu1f(di, eps) * ni[j];
      }
    } else if (dI < 0) {
      for (int i = 0; i < D; i++) a.mu0[o + n * i] = 0;
    }
  });
  T zero[3] = {0, 0, 0};
  BC(g, a.mu0.data(), zero, false, a.per);
  BC(g, a.V.data(), zero, a.exit, a.per);
}

}  // namespace

// =========================== C ABI (ctypes) =============================================
extern "C" {

struct wlo_config {
  int D;
  int n[3];  // interior cells
  float uBC[3];
  int perdir[3];
  int exitBC;
  int lambda;  // 0 quick, 1 cds, 2 vanLeer
  float nu;
  float dt0;
};

enum { WLO_U = 0, WLO_U0, WLO_F, WLO_P, WLO_SIGMA, WLO_V, WLO_MU0, WLO_MU1 };

// Flow(N,uBC;…)  src/Flow.jl:133-147 with a constant initial condition u0 = uBC;
// callers overwrite u through wlo_field and then call wlo_init_bc for function ICs.
void* wlo_create(const wlo_config* c) {
  Flow* a = new Flow();
  a->g.D = c->D;
  for (int d = 0; d < 3; d++) {
    a->g.N[d] = d < c->D ? c->n[d] + 2 : 1;
    a->uBC[d] = c->uBC[d];
    a->per[d] = d < c->D ? c->perdir[d] : 0;
  }
  const size_t n = a->g.n();
  const int D = c->D;
  a->u.assign(n * D, 0);
  for (int i = 0; i < D; i++) std::fill(a->u.begin() + i * n, a->u.begin() + (i + 1) * n, c->uBC[i]);
  a->f.assign(n * D, 0);
  a->p.assign(n, 0);
  a->sigma.assign(n, 0);
  a->V.assign(n * D, 0);
  a->mu0.assign(n * D, 1);
  a->mu1.assign(n * D * D, 0);
  a->dt.assign(1, c->dt0);
  a->nu = c->nu;
  a->exit = c->exitBC;
  a->lam = c->lambda;
  BC(a->g, a->u.data(), a->uBC, a->exit, a->per);
  exitBC(a->g, a->u.data(), a->u.data(), 0);
  a->u0 = a->u;
  T zero[3] = {0, 0, 0};
  BC(a->g, a->mu0.data(), zero, false, a->per);
  return a;
}
void wlo_destroy(void* h) { delete (Flow*)h; }
// after overwriting u with a function IC: BC!(u,…); exitBC!(u,u,0); u⁰=copy(u)  :141-142
void wlo_init_bc(void* h) {
  Flow* a = (Flow*)h;
  BC(a->g, a->u.data(), a->uBC, a->exit, a->per);
  exitBC(a->g, a->u.data(), a->u.data(), 0);
  a->u0 = a->u;
}
float* wlo_field(void* h, int id, uint64_t* len) {
  Flow* a = (Flow*)h;
  std::vector<T>* v = nullptr;
  switch (id) {
    case WLO_U: v = &a->u; break;
    case WLO_U0: v = &a->u0; break;
    case WLO_F: v = &a->f; break;
    case WLO_P: v = &a->p; break;
    case WLO_SIGMA: v = &a->sigma; break;
    case WLO_V: v = &a->V; break;
    case WLO_MU0: v = &a->mu0; break;
    case WLO_MU1: v = &a->mu1; break;
    default: return nullptr;
  }
  if (len) *len = v->size();
  return v->data();
}
void wlo_measure_sphere(void* h, const float* c, float R, float eps) {
  Body b{0, {c[0], c[1], c[2]}, R, 0};
  measure(*(Flow*)h, b, eps);
}
void wlo_measure_torus(void* h, const float* c, float R, float r, float eps) {
  Body b{1, {c[0], c[1], c[2]}, R, r};
  measure(*(Flow*)h, b, eps);
}
// measure!(flow, body; t, ϵ) for a body made of `np` primitives (struct layout = wl_body_prim of include/wl_b200.h)
void wlo_measure_prims(void* h, const void* prims, int np, float eps, float t) { measure_prims(*(Flow*)h, (const Prim*)prims, np, eps, t); }
// enumerated forcings (see struct Flow): g_i(t) = g0_i + g1_i·t, U_i(t) = uBC_i + U1_i·t + ½·U2_i·t²
void wlo_set_forcing(void* h, const float* g0, const float* g1, const float* U1, const float* U2) {
  Flow* a = (Flow*)h;
  a->forcing = true;
  for (int i = 0; i < 3; i++) a->g0[i] = g0[i], a->g1[i] = g1[i], a->U1[i] = U1[i], a->U2[i] = U2[i];
}
// pressure_force / viscous_force / pressure_moment / viscous_moment at time t about x0 (out: 12 doubles, see body_forces)
// udf = sgs! with νₜ = smagorinsky, Cs, Δ (src/util.jl:46-76); Cs·Δ = 0 switches it off
void wlo_set_sgs(void* h, float Cs, float Delta) {
  Flow* a = (Flow*)h;
  const T c = Cs * Delta;
  a->sgs = c != 0;
  a->sgs_c2 = c * c;
  a->S.assign(a->g.n() * a->g.D * a->g.D, 0);
}
void wlo_sgs(void* h, int from_u0) {  // sgs!(flow, u, t) alone: adds the sub-grid fluxes of u (or u⁰) to flow.f
  Flow* a = (Flow*)h;
  sgs(*a, from_u0 ? a->u0.data() : a->u.data());
}
void wlo_body_forces(void* h, const void* prims, int np, float t, const float* x0, double* out) { body_forces(*(Flow*)h, (const Prim*)prims, np, t, x0, out); }
// pois_ctor(flow): MultiLevelPoisson(flow.p,flow.μ₀,flow.σ;perdir) (kind 0) or Poisson(...) (kind 1)
int wlo_init_pois(void* h, int kind) {
  Flow* a = (Flow*)h;
  delete a->ml;
  delete a->single;
  a->ml = nullptr;
  a->single = nullptr;
  if (kind == 0) {
    a->ml = ml_create(a->g, a->p.data(), a->mu0.data(), a->sigma.data(), a->per);
    return a->ml ? (int)a->ml->levels.size() : -1;
  }
  a->single = new Poisson();
  pois_init(*a->single, a->g, a->p.data(), a->mu0.data(), a->sigma.data(), a->per);
  return 1;
}
void wlo_set_solver(void* h, double tol, int itmx, int smoother) {
  Flow* a = (Flow*)h;
  a->tol = tol;
  a->itmx = itmx;
  if (a->ml) a->ml->smoother = smoother;
}
void wlo_update(void* h) {  // update!(pois)
  Flow* a = (Flow*)h;
  if (a->ml) ml_update(*a->ml);
  if (a->single) set_diag(*a->single);
}
void wlo_mom_step(void* h) { mom_step(*(Flow*)h); }
void wlo_project(void* h, float w) { project(*(Flow*)h, w); }
void wlo_conv_diff(void* h, int from_u0) {
  Flow* a = (Flow*)h;
  conv_diff(a->g, a->f.data(), from_u0 ? a->u0.data() : a->u.data(), a->sigma.data(), a->lam, a->nu, a->per);
}
void wlo_bdim(void* h) { BDIM(*(Flow*)h); }
void wlo_bc_u(void* h) {
  Flow* a = (Flow*)h;
  BC(a->g, a->u.data(), a->uBC, a->exit, a->per);
}
float wlo_cfl(void* h) { return CFL(*(Flow*)h); }
int wlo_dt_len(void* h) { return (int)((Flow*)h)->dt.size(); }
void wlo_get_dt(void* h, float* out) {
  Flow* a = (Flow*)h;
  std::copy(a->dt.begin(), a->dt.end(), out);
}
void wlo_push_dt(void* h, float dt) { ((Flow*)h)->dt.push_back(dt); }
int wlo_iters_len(void* h) {
  Flow* a = (Flow*)h;
  return (int)(a->ml ? a->ml->n.size() : (a->single ? a->single->n.size() : 0));
}
void wlo_get_iters(void* h, int16_t* out) {
  Flow* a = (Flow*)h;
  const std::vector<int16_t>& n = a->ml ? a->ml->n : a->single->n;
  std::copy(n.begin(), n.end(), out);
}
int wlo_log_len(void* h) { return (int)((Flow*)h)->log.rows.size() / 3; }
void wlo_get_log(void* h, float* out) {
  Flow* a = (Flow*)h;
  std::copy(a->log.rows.begin(), a->log.rows.end(), out);
}
int wlo_num_levels(void* h) {
  Flow* a = (Flow*)h;
  return a->ml ? (int)a->ml->levels.size() : 1;
}
// level arrays: id 0 L, 1 D, 2 iD, 3 x, 4 eps, 5 r, 6 z;  dims written to N[3]
float* wlo_level_field(void* h, int level, int id, int* N) {
  Flow* a = (Flow*)h;
  Poisson* p = a->ml ? a->ml->levels[level] : a->single;
  for (int d = 0; d < 3; d++) N[d] = p->g.N[d];
  switch (id) {
    case 0: return p->L;
    case 1: return p->D.data();
    case 2: return p->iD.data();
    case 3: return p->x;
    case 4: return p->eps.data();
    case 5: return p->r.data();
    case 6: return p->z;
  }
  return nullptr;
}
// standalone operator API on the flow's Poisson (test/test_poisson.jl Poisson_setup)
void wlo_pois_mult(void* h, float* x) {  // mult!(pois,x) → z
  Flow* a = (Flow*)h;
  mult(a->ml ? *a->ml->levels[0] : *a->single, x);
}
int wlo_pois_solve(void* h) { return flow_solver(*(Flow*)h); }
void wlo_pois_residual(void* h) {
  Flow* a = (Flow*)h;
  residual(a->ml ? *a->ml->levels[0] : *a->single);
}
float wlo_pois_L2(void* h) {
  Flow* a = (Flow*)h;
  return L2(a->ml ? *a->ml->levels[0] : *a->single);
}
float wlo_pois_Linf(void* h) {
  Flow* a = (Flow*)h;
  return Linf(a->ml ? *a->ml->levels[0] : *a->single);
}
void wlo_pois_smooth(void* h, int level, int kind, float w) {  // 0 GS-RB(it=4), 1 Jacobi, 2 pcg
  Flow* a = (Flow*)h;
  Poisson& p = a->ml ? *a->ml->levels[level] : *a->single;
  if (kind == 0) GaussSeidelRB(p, 4, w);
  if (kind == 1) Jacobi(p, 1, w);
  if (kind == 2) pcg(p);
}
void wlo_pois_vcycle(void* h, float w) {
  Flow* a = (Flow*)h;
  Vcycle(*a->ml, 0, w);
}

// free functions used by the unit tests
float wlo_quick(float u, float c, float d) { return quick(u, c, d); }
float wlo_vanleer(float u, float c, float d) { return vanLeer(u, c, d); }
float wlo_cds(float u, float c, float d) { return cds(u, c, d); }
float wlo_mu0(float d, float e) { return mu0f(d, e); }
float wlo_mu1(float d, float e) { return mu1f(d, e); }
// 1-D flux helpers on a vector f of length n, 1-based index I  (test/test_flow.jl:11-41)
float wlo_phiuL(const float* f, int n, int I, float u, int l) {
  Grid g{1, {n, 1, 1}};
  return FluxCtx{g, f, 0, l}.phiuL(I3{{I, 1, 1}}, u);
}
float wlo_phiuR(const float* f, int n, int I, float u, int l) {
  Grid g{1, {n, 1, 1}};
  return FluxCtx{g, f, 0, l}.phiuR(I3{{I, 1, 1}}, u);
}
float wlo_phiu(const float* f, int n, int I, float u, int l) {
  Grid g{1, {n, 1, 1}};
  return FluxCtx{g, f, 0, l}.phiu(I3{{I, 1, 1}}, u);
}
float wlo_phiuP(const float* f, int n, int Ip, int I, float u, int l) {
  Grid g{1, {n, 1, 1}};
  return FluxCtx{g, f, 0, l}.phiuP(I3{{Ip, 1, 1}}, I3{{I, 1, 1}}, u);
}
float wlo_phi(const float* f, int n, int I) {
  Grid g{1, {n, 1, 1}};
  return FluxCtx{g, f, 0, 0}.phi(I3{{I, 1, 1}});
}
// array-level BC helpers (test/test_core.jl:19-70); N incl. ghosts
void wlo_BC(int D, const int* N, float* a, const float* U, int saveexit, const int* per) {
  Grid g{D, {N[0], N[1], D > 2 ? N[2] : 1}};
  BC(g, a, U, saveexit, per);
}
void wlo_exitBC(int D, const int* N, float* u, const float* u0, float dt) {
  Grid g{D, {N[0], N[1], D > 2 ? N[2] : 1}};
  exitBC(g, u, u0, dt);
}
void wlo_perBC(int D, const int* N, float* a, const int* per) {
  Grid g{D, {N[0], N[1], D > 2 ? N[2] : 1}};
  perBC(g, a, per);
}
// bench.py: under torchrun the environment carries OMP_NUM_THREADS=1; the CPU baseline must use all host cores it may run on
void wlo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int wlo_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp single
    n = omp_get_num_threads();
  }
#endif
  return n;
}

}  // extern "C"
