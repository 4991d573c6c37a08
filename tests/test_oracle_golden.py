"""Pins the CPU oracle (oracle/wl_oracle.cpp) to the reference's own known-answer tests.

Each test cites the WaterLily.jl test it restates (paths relative to the reference tree).
Arrays are numpy views in C order, i.e. the reference's `a[x,y,(z),c]` is `a[c,(z),y,x]` here.
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle
from oracle import OracleSim
from oracle.oracle import _fp, _ip


def L2in(a):
    """L₂(a) = sum(abs2, a[inside(a)])  (src/Poisson.jl:188)"""
    sl = tuple(slice(1, -1) for _ in a.shape)
    return float((a[sl].astype(np.float64) ** 2).sum())


def poisson_setup(N, kind):
    """Poisson_setup (test/test_poisson.jl:1-13): soln = x-index, z = A*soln, solve, compare."""
    dims = tuple(n - 2 for n in N)
    s = OracleSim(dims, (0.0,) * len(N), pois=kind)
    s.init_pois()
    x = s.level_field(0, "x")
    soln = np.zeros_like(x)
    soln[...] = np.arange(1, N[0] + 1, dtype=np.float32)  # T(I.I[1])
    first = (1,) * len(N)
    soln -= soln[first]
    s.L.wlo_pois_mult(s.h, _fp(soln))
    n = s.L.wlo_pois_solve(s.h)
    x -= x[first]
    return L2in(x - soln) / L2in(soln), s, n


# ---------------------------------------------------------------- test/test_poisson.jl
def test_poisson_diag_5x5():  # :16-19
    err, s, _ = poisson_setup((5, 5), "single")
    D = s.level_field(0, "D")
    iD = s.level_field(0, "iD")
    ref = np.array([[0, 0, 0, 0, 0], [0, -2, -3, -2, 0], [0, -3, -4, -3, 0], [0, -2, -3, -2, 0], [0, 0, 0, 0, 0]], np.float32)
    assert np.array_equal(D, ref)
    refi = np.array([[0, 0, 0, 0, 0], [0, -1 / 2, -1 / 3, -1 / 2, 0], [0, -1 / 3, -1 / 4, -1 / 3, 0], [0, -1 / 2, -1 / 3, -1 / 2, 0], [0, 0, 0, 0, 0]], np.float32)
    assert np.allclose(iD, refi, rtol=1e-7, atol=0)
    assert err < 1e-5


def test_poisson_pcg_2d_3d():  # :20-25
    err, s, n = poisson_setup((2**6 + 2, 2**6 + 2), "single")
    assert err < 1e-6 and n < 310
    err, s, n = poisson_setup((2**4 + 2,) * 3, "single")
    assert err < 1e-6 and n < 35


def test_multilevel_requires_levels():  # :41
    with pytest.raises(AssertionError):
        poisson_setup((15 + 2, 3**4 + 2), "ml")


@pytest.mark.parametrize("N,coarse", [((18, 18, 6), (10, 10, 4)), ((18, 18, 4), (10, 10, 4)), ((18, 17, 6), (10, 17, 4))])
def test_coarsen_mask(N, coarse):  # :44-46 (mask observed through the coarse level size)
    s = OracleSim(tuple(n - 2 for n in N), (0.0, 0.0, 0.0))
    s.init_pois()
    assert s.level_field(1, "x").shape == tuple(reversed(coarse))


def test_multilevel_coarse_diag_and_update():  # :53-59
    err, s, _ = poisson_setup((10, 10), "ml")
    assert s.nlevels == 3
    ref = np.array([[0, 0, 0, 0], [0, -2, -2, 0], [0, -2, -2, 0], [0, 0, 0, 0]], np.float32)
    assert np.array_equal(s.level_field(2, "D"), ref)
    assert err < 1e-5
    L = s.level_field(0, "L")  # L[c,y,x];  reference: L[5:6,:,1] .= 0
    L[0, :, 4:6] = 0
    s.update()
    assert np.array_equal(s.level_field(2, "D"), -np.abs(ref) / 2)


def test_multilevel_convergence():  # :61-67
    err, s, n = poisson_setup((2**6 + 2, 2**6 + 2), "ml")
    assert err < 1e-6 and n <= 3
    err, s, n = poisson_setup((2**4 + 2,) * 3, "ml")
    assert err < 1e-6 and n <= 3


def test_semicoarsening_channel_2d():  # :70-74
    H = 2**4
    R = H // 4
    s = OracleSim((8 * H, H), (1.0, 0.0), nu=R / 100)
    s.measure_sphere((4 * H, H // 2), R)
    for _ in range(4):
        s.mom_step()
    assert len(s.iters) == 8 and np.all(s.iters <= 10)
    assert np.isfinite(s.field("u")).all()


def test_semicoarsening_duct_3d():  # :75-79
    H = 2**3
    R = H // 4
    s = OracleSim((8 * H, H, H), (1.0, 0.0, 0.0), nu=R / 100)
    s.measure_sphere((4 * H, H // 2, H // 2), R)
    for _ in range(4):
        s.mom_step()
    assert len(s.iters) == 8 and np.all(s.iters <= 12)


# ------------------------------------------------------------------- test/test_flow.jl
def test_limiters(oracle_lib):  # :2-9
    L = oracle_lib
    assert L.wlo_vanleer(1, 0, 1) == 0 and L.wlo_vanleer(1, 2, 1) == 2
    assert L.wlo_vanleer(1, 2, 3) == 2.5 and L.wlo_vanleer(3, 2, 1) == 1.5
    assert L.wlo_cds(1, 0, 1) == 0.5 and L.wlo_cds(1, 2, -1) == 0.5


def test_boundary_fluxes(oracle_lib):  # :11-41
    L = oracle_lib
    f = np.array([0.0, 0.5, 2.0], np.float32)
    q = 0  # quick
    assert L.wlo_phiuL(_fp(f), 3, 2, 1.0, q) == L.wlo_phi(_fp(f), 3, 2)
    assert L.wlo_phiuL(_fp(f), 3, 2, -1.0, q) == -L.wlo_quick(2.0, 0.5, 0.0)
    assert L.wlo_phiuR(_fp(f), 3, 3, 1.0, q) == L.wlo_quick(0.0, 0.5, 2.0)
    assert L.wlo_phiuR(_fp(f), 3, 3, -1.0, q) == -L.wlo_phi(_fp(f), 3, 3)
    assert L.wlo_phiu(_fp(f), 3, 3, 1.0, q) == L.wlo_phiuP(_fp(f), 3, 1, 3, 1.0, q)
    assert L.wlo_phiu(_fp(f), 3, 2, -1.0, q) == L.wlo_phiuP(_fp(f), 3, 0 + 1, 2, -1.0, q)
    g = np.array([1.0, 1.25, 1.5, 1.75, 2.0], np.float32)
    assert L.wlo_phiuP(_fp(g), 5, 1, 3, 1.0, q) == L.wlo_quick(g[0], g[1], g[2])
    assert L.wlo_phiuP(_fp(g), 5, 3, 3, 1.0, q) == L.wlo_quick(g[2], g[1], g[2])  # Ip = CIj(1,I,length(f)-2)


def test_impulsive_uniform_flow():  # :76-84
    U = (2 / 3, -1 / 3)
    s = OracleSim((2**4, 2**4), U)
    s.mom_step()
    u = s.field("u")
    assert L2in(u[0] - np.float32(U[0])) < 2e-5
    assert L2in(u[1] - np.float32(U[1])) < 1e-5


def tgv2d(i, x, y, t, k, nu):
    e = math.exp(-2 * k * k * nu * t)
    return (-np.sin(k * x) * np.cos(k * y) if i == 0 else np.cos(k * x) * np.sin(k * y)) * e


def tgv2d_field(N, k, nu, t):
    """apply!((i,x)->TGV(i,x,t,κ,ν),u): faces at loc(i,I) = I-1.5-δ_i/2  (src/core.jl:177)"""
    u = np.zeros((2, N, N), np.float32)
    idx = np.arange(1, N + 1, dtype=np.float32)
    for i in range(2):
        X = (idx - np.float32(1.5) - (np.float32(0.5) if i == 0 else 0))[None, :]
        Y = (idx - np.float32(1.5) - (np.float32(0.5) if i == 1 else 0))[:, None]
        u[i] = tgv2d(i, X, Y, t, np.float32(k), nu).astype(np.float32)
    return u


def test_tgv_2d_periodic():  # :100-109 with helper.jl:4-17 (T=Float32, Re=1e8)
    L = 64
    k = np.float32(2 * np.pi / L)
    nu = np.float32(1 / (k * 1e8))
    s = OracleSim((L, L), (0.0, 0.0), nu=float(nu), perdir=(1, 2), u0=tgv2d_field(L + 2, k, nu, 0.0))
    s.sim_step_until(np.pi / 100, U=1.0, Lscale=L)
    ue = tgv2d_field(L + 2, k, float(nu), s.time())
    u = s.field("u")
    assert L2in(u[0] - ue[0]) < 1e-4 and L2in(u[1] - ue[1]) < 1e-4


def test_sim_step_stop_time():  # test/test_simulation.jl:15-19
    radius = 8
    s = OracleSim((4 * radius,) * 2, (1.0, 0.0), nu=radius / 250)
    s.measure_sphere((2 * radius, 2 * radius), radius)
    s.sim_step_until(0.1, U=1.0, Lscale=radius)
    dt = s.dt
    assert s.time() / radius >= 0.1 > float(np.sum(dt[:-2])) / radius
    assert np.all(s.iters < 5)


# -------------------------------------------------------------------- test/test_core.jl
def test_BC_and_exitBC(oracle_lib):  # :19-48
    L = oracle_lib
    rng = np.random.default_rng(0)
    Ng, U = (6, 6), np.array([1.0, 0.5, 0.0], np.float32)
    noper = _ip([0, 0, 0])
    u = rng.random((2, 6, 6)).astype(np.float32)  # u[c,y,x]
    L.wlo_BC(2, _ip(Ng), _fp(u), _fp(U), 0, noper)
    assert (u[0, :, 0] == U[0]).all() and (u[0, :, 1] == U[0]).all() and (u[0, :, -1] == U[0]).all()
    assert (u[0, 0, 2:-1] == u[0, 1, 2:-1]).all() and (u[0, -1, 2:-1] == u[0, -2, 2:-1]).all()
    assert (u[1, 0, :] == U[1]).all() and (u[1, 1, :] == U[1]).all() and (u[1, -1, :] == U[1]).all()
    assert (u[1, 2:-1, 0] == u[1, 2:-1, 1]).all() and (u[1, 2:-1, -1] == u[1, 2:-1, -2]).all()
    u[0, :, -1] = 3
    L.wlo_BC(2, _ip(Ng), _fp(u), _fp(U), 1, noper)
    assert (u[0, :, -1] == 3).all()
    L.wlo_exitBC(2, _ip(Ng), _fp(u), _fp(u), 0.0)
    assert (u[0, 1:-1, -1] == U[0]).all()
    L.wlo_BC(2, _ip(Ng), _fp(u), _fp(U), 1, _ip([0, 1, 0]))  # periodic in y, save exit
    assert (u[0, 0:2, :] == u[0, -2:, :]).all()
    sg = rng.random((6, 6)).astype(np.float32)
    L.wlo_perBC(2, _ip(Ng), _fp(sg), _ip([1, 1, 0]))
    assert (sg[1:-1, 0] == sg[1:-1, -2]).all() and (sg[0, 1:-1] == sg[-2, 1:-1]).all()
    u = rng.random((2, 6, 6)).astype(np.float32)
    L.wlo_BC(2, _ip(Ng), _fp(u), _fp(U), 1, _ip([1, 0, 0]))  # x-periodic: saveexit has no effect
    assert (u[:, :, 0:2] == u[:, :, -2:]).all()
    assert (u[1, 0, :] == U[1]).all() and (u[1, 1, :] == U[1]).all() and (u[1, -1, :] == U[1]).all()


# ------------------------------------------------------------------ test/test_bodies.jl
def test_kernel_moments(oracle_lib):  # :2-5
    L = oracle_lib
    assert L.wlo_mu0(3.0, 6.0) == L.wlo_mu0(0.5, 1.0)
    assert L.wlo_mu0(0.0, 1.0) == 0.5
    assert L.wlo_mu0(float(np.float32(np.finfo(np.float32).eps) - np.float32(1)), 1.0) == 0.0
    assert abs(L.wlo_mu1(0.0, 2.0) - 2 * (1 / 4 - 1 / np.pi**2)) < 1e-7


# ---------------------------------------------------------------- src/util.jl:46-76 (sgs! with the docstring's smagorinsky νₜ)
def _sgs_numpy(f, u, sigma, Cs, Delta):
    """An array-level transliteration of sgs! (src/util.jl:66-76), S(I,u) and ∂(i,j,I,u) (src/Metrics.jl:42-44,140), independent of
    the C++ restatement.  f, u: reference index order (x,y,(z),c); ranges below are the reference's 1-based ranges."""
    F = np.float32
    D = u.shape[-1]
    N = u.shape[:-1]

    def sl(rng):
        return tuple(slice(lo - 1, hi) for lo, hi in rng)

    def shift(rng, d, k):
        r = list(rng)
        r[d] = (r[d][0] + k, r[d][1] + k)
        return r

    inside = [(2, n - 1) for n in N]

    def dudx(i, j, rng):
        ui = u[..., i]
        if i == j:
            return ui[sl(shift(rng, i, 1))] - ui[sl(rng)]
        P, M = shift(rng, j, 1), shift(rng, j, -1)
        return (ui[sl(P)] + ui[sl(shift(P, i, 1))] - ui[sl(M)] - ui[sl(shift(M, i, 1))]) / F(4)

    S = np.zeros(N + (D, D), F)
    for j in range(D):
        for i in range(D):
            S[sl(inside) + (i, j)] = (dudx(i, j, inside) + dudx(j, i, inside)) / F(2)
    s = np.zeros(N, F)
    for j in range(D):
        for i in range(D):
            s = s + S[..., i, j] * S[..., i, j]
    c = F(Cs) * F(Delta)
    nut = (c * c) * np.sqrt(s)
    for i in range(D):
        for j in range(D):
            R = [(3, N[d] - 1) if d == j else (2, N[d]) for d in range(D)]
            Rm = shift(R, j, -1)
            sigma[sl(R)] = -nut[sl(R)] * (u[..., i][sl(R)] - u[..., i][sl(Rm)])
            f[..., i][sl(R)] += sigma[sl(R)]
            f[..., i][sl(Rm)] -= sigma[sl(R)]


@pytest.mark.parametrize("dims", [(12, 10), (10, 8, 6)])
def test_sgs_restatement_equals_array_transliteration(dims):
    D = len(dims)
    rng = np.random.default_rng(5)
    s = OracleSim(dims, (1.0,) + (0.0,) * (D - 1), nu=0.01)
    u, f, sg = s.field("u"), s.field("f"), s.field("sigma")
    u[...] = rng.standard_normal(u.shape).astype(np.float32)
    f[...] = rng.standard_normal(f.shape).astype(np.float32)
    sg[...] = rng.standard_normal(sg.shape).astype(np.float32)
    ax = tuple(range(D, 0, -1)) + (0,)  # C order (c,(z),y,x) → reference order (x,y,(z),c)
    fr, ur, sr = np.transpose(f.copy(), ax), np.transpose(u.copy(), ax), np.transpose(sg.copy(), tuple(range(D - 1, -1, -1)))
    _sgs_numpy(fr, ur, sr, 0.2, 1.5)
    s.set_sgs(0.2, 1.5)
    s.sgs()
    assert np.array_equal(np.transpose(f, ax), fr)
    assert np.array_equal(np.transpose(sg, tuple(range(D - 1, -1, -1))), sr)
    # a uniform stream has no strain: the model adds exactly nothing
    u[...] = 0.75
    before = f.copy()
    s.sgs()
    assert np.array_equal(f, before)


# ---------------------------------------------------------------- test/test_flow.jl:161-173 ("Circle in accelerating flow")
def test_accelerating_circle_added_mass():
    """A circle of radius 32 in a 1024² box whose boundary velocity is uBC(t) = (t, 0): after one step the pressure force is the added
    mass of a cylinder, pressure_force/(πL²) ≈ (−1, 0) ± 0.04, the potential flow peaks at ≈ 2U over the body and the solver needs
    ≤ 2 V-cycles — through the enumerated boundary velocity (U1), accelerate!'s dU/dt term, BDIM, the projection and the force
    reduction (src/Flow.jl:64-73,156-232; src/Metrics.jl:116-127)."""
    radius, H = 32, 16
    n = radius * 2 * H
    s = OracleSim((n, n), (0.0, 0.0))
    s.measure_sphere((H * radius, H * radius), float(radius))
    s.init_pois()
    s.set_forcing(U1=(1.0, 0.0))
    s.mom_step()
    prim = [dict(kind=0, op=0, center=(H * radius, H * radius), R=float(radius), r=0.0, vel=(0.0, 0.0))]
    pf = s.body_forces(prim)[0] / (math.pi * radius ** 2)
    assert abs(pf[0] + 1.0) < 0.04 and abs(pf[1]) < 0.04, pf
    u = s.field("u")
    assert float(u.max()) / float(u[0][1, 1]) > 1.91
    for _ in range(3):
        s.mom_step()
    assert all(int(k) <= 2 for k in s.iters), list(s.iters)


# ---------------------------------------------------------------- test/test_simulation.jl:20-25
@pytest.mark.parametrize("exitBC", [True, False])
def test_remeasure_body_moving_with_the_stream(exitBC):
    """"remeasure works perfectly when V = U = 1": a circle translating with the free stream (AutoBody(circle, x − [t,0]), measured at
    t = ΣΔt before the step) leaves the flow undisturbed: u[:, radius, 1] ≈ 1 (isapprox of Float32: rtol √eps) after sim_step!(sim)."""
    radius = 8
    s = OracleSim((4 * radius, 4 * radius), (1.0, 0.0), nu=radius / 250, exitBC=exitBC)
    prims = [dict(kind=0, op=0, center=(2.0 * radius, 2.0 * radius), R=float(radius), r=0.0, vel=(1.0, 0.0))]
    s.measure_prims(prims, 1.0, 0.0)
    s.init_pois()
    s.measure_prims(prims, 1.0, s.time_next())  # sim_step!(sim): remeasure=true by default (src/WaterLily.jl:136-139,146-149)
    s.update()
    s.mom_step()
    row = s.field("u")[0][radius - 1, :]
    assert np.all(np.abs(row - 1.0) <= math.sqrt(np.finfo(np.float32).eps) * np.maximum(np.abs(row), 1.0)), row
