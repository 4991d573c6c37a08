"""GPU parity tests proper: the sm_100a path, called through the C ABI, against the CPU oracle.

The library is built with -fmad=false and the oracle with -ffp-contract=off, so the two execute the same IEEE
Float32 operations in the same order: operator-level results must agree to the last bit (a few ulp where a
double-precision reduction order differs).  Whole-simulation tests assert the tolerance BASELINE.json states —
relative L2 ≤ 1e-5 on u and p after 100 steps, Poisson iterations ±1 — and in practice measure 0.
"""
import numpy as np
import pytest

from util import F, make_pair, max_ulp, rel_l2, smooth_field, tgv3d_u0

pytestmark = pytest.mark.gpu

CASES = {
    "3d_box": dict(dims=(16, 12, 8), uBC=(1.0, 0.0, 0.0), nu=0.05),
    "3d_per": dict(dims=(16, 16, 16), uBC=(0.0, 0.0, 0.0), nu=0.01, perdir=(1, 2, 3)),
    "3d_mixed_exit": dict(dims=(24, 16, 8), uBC=(1.0, 0.0, 0.0), nu=0.02, perdir=(3,), exitBC=True),
    "2d_box": dict(dims=(32, 16), uBC=(1.0, 0.5), nu=0.03),
    "2d_per_y": dict(dims=(16, 32), uBC=(1.0, 0.0), nu=0.03, perdir=(2,)),
}


def upload_same_u(o, s, seed=1):
    D = o.D
    u = smooth_field(o.N, D, seed)
    o.field("u")[...] = u
    o.L.wlo_init_bc(o.h)
    s.flow.upload("u", u)
    from wl_b200 import lib as wlib
    wlib.check(s.flow.L, s.flow.L.wl_apply_bc(s.flow.h))


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("lam", ["quick", "cds", "vanLeer"])
def test_bc_and_conv_diff_bit_exact(name, lam):
    o, s = make_pair(**CASES[name], lam=lam)
    upload_same_u(o, s)
    assert np.array_equal(o.field("u"), s.flow.u), "BC!/exitBC! ghost fill differs"
    assert np.array_equal(o.field("u0"), s.flow.u0)
    o.L.wlo_conv_diff(o.h, 0)
    s.flow.L.wl_conv_diff(s.flow.h, 0)
    fo, fg = o.field("f"), s.flow.f
    assert max_ulp(fg, fo) == 0.0, f"conv_diff! differs by {max_ulp(fg, fo)} ulp"
    # stale-Φ bookkeeping on ghost cells of σ
    so, sg = o.field("sigma"), s.flow.σ
    ghost = np.ones(so.shape, bool)
    ghost[tuple(slice(1, -1) for _ in so.shape)] = False
    assert np.array_equal(so[ghost], sg[ghost])


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("lam", ["quick", "cds", "vanLeer"])
def test_mom_step_two_steps(name, lam):
    """Whole steps through the production kernels (3-D: fm_conv<λ> + k_bdim2, 2-D: k_conv_bdim1<λ>) for every limiter."""
    sphere = None
    cfg = dict(CASES[name], lam=lam)
    if name in ("3d_box", "3d_mixed_exit"):
        sphere = (tuple(d / 2 for d in cfg["dims"]), 3.0)
    if name == "2d_box":
        sphere = ((10.0, 8.0), 4.0)
    o, s = make_pair(**cfg, sphere=sphere)
    upload_same_u(o, s, seed=2)
    for _ in range(2):
        o.mom_step()
        s.flow.L.wl_mom_step(s.flow.h)
    assert list(o.iters) == list(s.pois.n)
    assert rel_l2(s.flow.u, o.field("u")) < 2e-6
    assert rel_l2(s.flow.p, o.field("p")) < 2e-5
    assert np.allclose(o.dt, s.flow.Δt, rtol=1e-6)


def test_body_measure_matches_oracle():
    o, s = make_pair((32, 24, 16), (1.0, 0.0, 0.0), nu=0.01, sphere=((12.0, 11.0, 8.0), 5.0), exitBC=True)
    for name in ("mu0", "mu1", "V", "sigma"):
        a, b = o.field(name), getattr(s.flow, {"mu0": "μ0", "mu1": "μ1", "V": "V", "sigma": "σ"}[name])
        assert max_ulp(b, a) <= 2.0, name
    for lvl in range(s.pois.nlevels):
        for arr in ("L", "D", "iD"):
            assert max_ulp(s.pois.level(lvl, arr), o.level_field(lvl, arr)) <= 4.0, (lvl, arr)


@pytest.mark.parametrize("N", [(10, 10), (66, 66), (18, 18, 18), (34, 18, 10)])
def test_poisson_operators(N):
    """mult!, residual!, Jacobi!, GaussSeidelRB!, Vcycle! level by level (src/Poisson.jl, src/MultiLevelPoisson.jl)."""
    dims = tuple(n - 2 for n in N)
    D = len(N)
    o, s = make_pair(dims, (0.0,) * D)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(tuple(reversed(N))).astype(F)
    z = rng.standard_normal(tuple(reversed(N))).astype(F)
    inner = tuple(slice(1, -1) for _ in N)
    z[inner] -= z[inner].mean()
    # mult!
    s.flow.upload("p", x)
    s.pois.mult()
    xo = x.copy()
    from oracle.oracle import _fp
    o.L.wlo_pois_mult(o.h, _fp(xo))
    assert max_ulp(s.flow.σ, o.field("sigma")) == 0.0
    # residual! + L2
    o.field("p")[...] = x
    o.field("sigma")[...] = z
    s.flow.upload("p", x)
    s.flow.upload("sigma", z)
    o.L.wlo_pois_residual(o.h)
    r2g = s.pois.residual()
    assert max_ulp(s.pois.level(0, "r"), o.level_field(0, "r")) <= 1.0
    assert abs(r2g - o.L.wlo_pois_L2(o.h)) <= 2e-6 * r2g
    # smoothers on the finest level
    for kind, kid, w in (("jacobi", 1, 1.0), ("gs", 0, 0.9)):  # Vcycle! always calls Jacobi!(fine) with ω=1
        o.L.wlo_pois_smooth(o.h, 0, kid, w)
        s.pois.smooth(0, kind, w)
        assert max_ulp(s.pois.level(0, "r"), o.level_field(0, "r")) <= 1.0, kind
        assert max_ulp(s.pois.level(0, "x"), o.level_field(0, "x")) <= 1.0, kind
    # one V-cycle
    o.L.wlo_pois_vcycle(o.h, 0.95)
    s.pois.vcycle(0.95)
    for lvl in range(s.pois.nlevels):
        assert max_ulp(s.pois.level(lvl, "r"), o.level_field(lvl, "r")) <= 2.0, lvl
        assert max_ulp(s.pois.level(lvl, "x"), o.level_field(lvl, "x")) <= 2.0, lvl


def _poisson_setup_gpu(N, kind):
    """Poisson_setup of test/test_poisson.jl:1-13 through the C ABI."""
    import wl_b200 as wl
    dims = tuple(n - 2 for n in N)
    D = len(N)
    s = wl.Simulation(dims, (0.0,) * D, 1.0, pois=kind)
    soln = np.zeros(tuple(reversed(N)), F)
    soln[...] = np.arange(1, N[0] + 1, dtype=F)
    first = (1,) * D
    soln -= soln[first]
    s.flow.upload("p", soln)
    s.pois.mult()
    s.flow.upload("p", np.zeros_like(soln))
    n = s.pois.solver()
    x = s.flow.p
    x -= x[first]
    inner = tuple(slice(1, -1) for _ in N)
    err = float(((x - soln)[inner].astype(np.float64) ** 2).sum() / (soln[inner].astype(np.float64) ** 2).sum())
    return err, s, n


def test_reference_poisson_known_answers_on_gpu():  # test/test_poisson.jl:15-25,53-67
    err, s, n = _poisson_setup_gpu((5, 5), "single")
    ref = np.array([[0, 0, 0, 0, 0], [0, -2, -3, -2, 0], [0, -3, -4, -3, 0], [0, -2, -3, -2, 0], [0, 0, 0, 0, 0]], F)
    assert np.array_equal(s.pois.level(0, "D"), ref) and err < 1e-5
    err, s, n = _poisson_setup_gpu((66, 66), "single")
    assert err < 1e-6 and n < 310
    err, s, n = _poisson_setup_gpu((18, 18, 18), "single")
    assert err < 1e-6 and n < 35
    err, s, n = _poisson_setup_gpu((10, 10), "multilevel")
    assert np.array_equal(s.pois.level(2, "D"), np.array([[0, 0, 0, 0], [0, -2, -2, 0], [0, -2, -2, 0], [0, 0, 0, 0]], F))
    assert err < 1e-5
    L = s.pois.level(0, "L")
    L[0, :, 4:6] = 0
    s.flow.upload("mu0", L)
    s.pois.update()
    assert np.array_equal(s.pois.level(2, "D"), np.array([[0, 0, 0, 0], [0, -1, -1, 0], [0, -1, -1, 0], [0, 0, 0, 0]], F))
    err, s, n = _poisson_setup_gpu((66, 66), "multilevel")
    assert err < 1e-6 and n <= 3
    err, s, n = _poisson_setup_gpu((18, 18, 18), "multilevel")
    assert err < 1e-6 and n <= 3
    import wl_b200 as wl
    with pytest.raises(wl.WLError):
        _poisson_setup_gpu((17, 83), "multilevel")


def test_reference_flow_known_answers_on_gpu():  # test/test_flow.jl:76-84, :100-109; test/test_poisson.jl:70-79
    import wl_b200 as wl
    from test_oracle_golden import L2in, tgv2d_field
    U = (2 / 3, -1 / 3)
    s = wl.Simulation((16, 16), U, 1.0)
    wl.sim_step(s)
    u = s.flow.u
    assert L2in(u[0] - F(U[0])) < 2e-5 and L2in(u[1] - F(U[1])) < 1e-5
    L = 64
    k = F(2 * np.pi / L)
    nu = F(1 / (k * 1e8))
    u0 = tgv2d_field(L + 2, k, nu, 0.0)
    s = wl.Simulation((L, L), (0.0, 0.0), L, U=1.0, ν=float(nu), perdir=(1, 2), u0=lambda i, x: u0[i])
    wl.sim_step(s, np.pi / 100)
    ue = tgv2d_field(L + 2, k, float(nu), s.flow.time())
    u = s.flow.u
    assert L2in(u[0] - ue[0]) < 1e-4 and L2in(u[1] - ue[1]) < 1e-4
    H = 16
    R = H // 4
    s = wl.Simulation((8 * H, H), (1.0, 0.0), R, ν=R / 100, body=wl.Sphere((4 * H, H // 2), R))
    for _ in range(4):
        wl.sim_step(s)
    assert len(s.pois.n) == 8 and np.all(s.pois.n <= 10)
    H = 8
    R = H // 4
    s = wl.Simulation((8 * H, H, H), (1.0, 0.0, 0.0), R, ν=R / 100, body=wl.Sphere((4 * H, H // 2, H // 2), R))
    for _ in range(4):
        wl.sim_step(s)
    assert len(s.pois.n) == 8 and np.all(s.pois.n <= 12)


def _hundred_steps(o, s, nsteps=100):
    import wl_b200 as wl
    for _ in range(nsteps):
        o.mom_step()
    wl.lib.check(s.flow.L, s.flow.L.wl_sim_step_n(s.flow.h, nsteps))
    eu = rel_l2(s.flow.u, o.field("u"))
    ep = rel_l2(s.flow.p, o.field("p"))
    dn = np.abs(np.asarray(o.iters, int) - np.asarray(s.pois.n, int))
    return eu, ep, dn


def test_tgv3d_100_steps_parity():
    """BASELINE.json north_star gate: rel. L2 ≤ 1e-5 on u and p after 100 steps, Poisson iteration count ±1."""
    n = 32
    u0 = tgv3d_u0((n + 2,) * 3, n)
    nu = float(F(1 / (2 * np.pi / n * 1600)))
    o, s = make_pair((n,) * 3, (0.0, 0.0, 0.0), nu=nu, perdir=(1, 2, 3), u0=u0)
    eu, ep, dn = _hundred_steps(o, s)
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)
    assert dn.max() <= 1
    assert np.allclose(o.dt, s.flow.Δt, rtol=1e-5)


def test_sphere_100_steps_parity():
    o, s = make_pair((64, 32, 32), (1.0, 0.0, 0.0), nu=8 / 100, sphere=((15.0, 15.0, 15.0), 4.0), exitBC=True)
    eu, ep, dn = _hundred_steps(o, s)
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)
    assert dn.max() <= 1


def test_circle_2d_100_steps_parity():
    """BASELINE.json configs[0]: README circle, dims (96,64), Re=100 (README.md:41-57)."""
    o, s = make_pair((96, 64), (1.0, 0.0), nu=16 / 100, sphere=((31.0, 31.0), 8.0))
    eu, ep, dn = _hundred_steps(o, s)
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)
    assert dn.max() <= 1


PCG_CASES = {
    # solver!(::Poisson) = pcg! until Σr² < tol (src/Poisson.jl:204-214), with a body / fully periodic (perdot, src/Poisson.jl:156-157)
    "single_2d_body": dict(dims=(32, 32), uBC=(1.0, 0.0), nu=0.05, sphere=((12.0, 16.0), 4.0), pois="single"),
    "single_2d_periodic": dict(dims=(32, 32), uBC=(0.0, 0.0), nu=0.01, perdir=(1, 2), pois="single"),
    "single_3d_periodic": dict(dims=(16, 16, 16), uBC=(0.0, 0.0, 0.0), nu=0.01, perdir=(1, 2, 3), pois="single"),
    # smooth!(p) = pcg!(p) inside the V-cycle (src/MultiLevelPoisson.jl:106 rebound; test/test_poisson.jl:62-67 runs both smoothers)
    "ml_pcg_2d_body": dict(dims=(64, 32), uBC=(1.0, 0.0), nu=0.05, sphere=((20.0, 16.0), 5.0), smoother="pcg"),
    "ml_pcg_3d_box": dict(dims=(32, 16, 16), uBC=(1.0, 0.0, 0.0), nu=0.05, sphere=((10.0, 8.0, 8.0), 3.0), smoother="pcg"),
    "ml_pcg_3d_periodic": dict(dims=(32, 32, 32), uBC=(0.0, 0.0, 0.0), nu=0.01, perdir=(1, 2, 3), smoother="pcg"),
}


@pytest.mark.parametrize("name", list(PCG_CASES))
def test_pcg_solver_and_smoother(name):
    """pcg! as the single-level solver and as the V-cycle smoother, walls / body / periodic (perdot), held to the north-star
    tolerance (1e-5; the dots are double-accumulated on both sides, so the iterates agree to rounding)."""
    cfg = PCG_CASES[name]
    o, s = make_pair(**cfg)
    if cfg.get("perdir"):
        upload_same_u(o, s, seed=4)
    from wl_b200 import lib as wlib
    for _ in range(3):
        o.mom_step()
        wlib.check(s.flow.L, s.flow.L.wl_mom_step(s.flow.h))
    assert np.abs(np.asarray(o.iters, int) - np.asarray(s.pois.n, int)).max() <= 1, (list(o.iters), list(s.pois.n))
    eu, ep = rel_l2(s.flow.u, o.field("u")), rel_l2(s.flow.p, o.field("p"))
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)


def test_div6_is_ieee_division_for_all_floats():
    """The flux kernels evaluate quick's (5c+2d-u)/6 with two FMAs instead of a division; the device checks all 2^32 inputs."""
    import ctypes as C
    import wl_b200 as wl
    L = wl.load_library()
    n = C.c_uint64(1)
    wl.lib.check(L, L.wl_selftest_div6(C.byref(n)))
    assert n.value == 0


def _tgv_pair(dims, oracle_too=True):
    n = dims[0]
    u0 = tgv3d_u0(tuple(d + 2 for d in dims), n)
    # break the symmetry a little so that every direction carries a different signal
    u0[2] = (0.1 * np.roll(u0[0], 3, axis=0)).astype(F)
    nu = float(F(1 / (2 * np.pi / n * 1600)))
    return make_pair(dims, (0.0, 0.0, 0.0), nu=nu, perdir=(1, 2, 3), u0=u0)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,nz", [((64, 64, 64), 0), ((128, 96, 64), 2)])
def test_fused_upstroke_matches_oracle(dims, nz):
    """f_vsmooth (prolongation + GaussSeidelRB! + increments in one pass, wl_vsmooth.cuh) and fm_conv4 on grids large enough to
    take those paths, with several tiles per direction and (nz=2) several z chunks: bit-identical to the oracle."""
    o, s = _tgv_pair(dims)
    if nz:
        from wl_b200 import lib as wlib
        wlib.check(s.flow.L, s.flow.L.wl_set_tuning(s.flow.h, b"vs_nz", nz))
    for _ in range(4):
        o.mom_step()
        s.flow.L.wl_mom_step(s.flow.h)
    assert np.array_equal(s.flow.u, o.field("u"))
    assert np.array_equal(s.flow.p, o.field("p"))
    assert list(np.asarray(o.iters)) == list(np.asarray(s.pois.n))
    assert np.array_equal(np.asarray(o.dt, F), np.asarray(s.flow.Δt, F))


@pytest.mark.gpu
def test_fused_kernels_equal_unfused_kernels():
    """128³ (two levels take f_vsmooth): the fused uniform-mode kernels against the separate march kernels, bit for bit."""
    import wl_b200 as wl
    outs = []
    for flags in (0, wl.lib.FLAGS["no_vsmooth"] | wl.lib.FLAGS["no_conv4"] | wl.lib.FLAGS["no_fused_uni"]):
        n = 128
        u0 = tgv3d_u0((n + 2,) * 3, n)
        u0[2] = (0.1 * np.roll(u0[0], 3, axis=0)).astype(F)
        nu = float(F(1 / (2 * np.pi / n * 1600)))
        s = wl.Simulation((n,) * 3, (0.0, 0.0, 0.0), float(n), ν=nu, perdir=(1, 2, 3), u0=lambda i, x: u0[i], flags=flags)
        wl.lib.check(s.flow.L, s.flow.L.wl_sim_step_n(s.flow.h, 3))
        outs.append((s.flow.u.copy(), s.flow.p.copy(), list(s.pois.n), np.asarray(s.flow.Δt).copy()))
    a, b = outs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[2] == b[2] and np.array_equal(a[3], b[3])


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [True, False])
def test_tiny_velocities_take_the_exact_division(periodic):
    """A quiescent far field holds velocities like 1e-34 (first steps of a wake; tails of a localised vortex): outside the
    proven range of the flux kernels' division-free x/6.  Both flux kernels (fm_conv4 in uniform mode, fm_conv with walls) must
    fall back to the IEEE division there and stay bit-identical to the oracle — not raise an error."""
    dims = (64, 64, 64) if periodic else (64, 32, 32)
    N = tuple(d + 2 for d in dims)
    rng = np.random.default_rng(5)
    u0 = tgv3d_u0(N, dims[0])
    u0[0] *= F(0.5)
    u0[1] = (1e-33 * rng.standard_normal(u0[1].shape)).astype(F)   # below 2^-100 ≈ 7.9e-31
    u0[2] = (1e-40 * rng.standard_normal(u0[2].shape)).astype(F)   # denormal
    if periodic:
        o, s = make_pair(dims, (0.0, 0.0, 0.0), nu=0.01, perdir=(1, 2, 3), u0=u0)
    else:
        o, s = make_pair(dims, (0.5, 0.0, 0.0), nu=0.01, u0=u0)
    from wl_b200 import lib as wlib
    for _ in range(2):
        o.mom_step()
        wlib.check(s.flow.L, s.flow.L.wl_mom_step(s.flow.h))
    s.flow.sync()
    assert np.array_equal(s.flow.u, o.field("u"))
    assert np.array_equal(s.flow.p, o.field("p"))


@pytest.mark.gpu
def test_upload_component_equals_upload():
    """wl_upload_component (apply! per component) and the staged host transfers: component-wise upload = whole-field upload,
    and a download returns exactly what was uploaded (reference layout, ghost cells included)."""
    o, s = make_pair((24, 16, 8), (1.0, 0.0, 0.0), nu=0.02)
    u = smooth_field(o.N, 3, seed=3)
    s.flow.upload("u", u)
    a = s.flow.download("u")
    for i in range(3):
        s.flow.upload_component("u", i, u[i] * F(2))
    b = s.flow.download("u")
    assert np.array_equal(a, u) and np.array_equal(b, u * F(2))
    with pytest.raises(Exception):
        s.flow.upload_component("u", 3, u[0])
    with pytest.raises(Exception):
        s.flow.upload_component("p", 1, u[0])


# ---- round 2: the north-star gate on every BASELINE.json config family, at sizes that take the production kernels ----------
def _gate(o, s, nsteps=100, exact=True):
    eu, ep, dn = _hundred_steps(o, s, nsteps)
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)
    assert dn.max() <= 1
    if exact:  # in practice the two paths execute the same IEEE operations
        assert np.array_equal(s.flow.u, o.field("u")) and np.array_equal(s.flow.p, o.field("p"))
        assert list(np.asarray(o.iters)) == list(np.asarray(s.pois.n))
        assert np.array_equal(np.asarray(o.dt, F), np.asarray(s.flow.Δt, F))


def test_tgv128_100_steps_gate():
    """BASELINE.json configs[1]: 3-D TGV 128³, periodic, 100 steps — fm_conv4 with several z chunks, f_vsmooth on two levels,
    f_jacobi_uni2, f_divres_uni, f_correct_cfl, the persistent coarse-level kernel, over a developing flow (ω back-off,
    mean-removal and n_V ≠ 2 branches as they occur)."""
    n = 128
    u0 = tgv3d_u0((n + 2,) * 3, n)
    nu = float(F(1 / (2 * np.pi / n * 1600)))
    o, s = make_pair((n,) * 3, (0.0, 0.0, 0.0), nu=nu, perdir=(1, 2, 3), u0=u0)
    _gate(o, s)
    import wl_b200 as wl
    flag = wl.lib.C.c_int()
    s.flow.L.wl_is_const_coeff(s.flow.h, wl.lib.C.byref(flag))
    assert flag.value == 1


def test_sphere_128_100_steps_gate():
    """configs[2] family (sphere wake, exit BC) at 128×64×64: general-mode march kernels with semi-uniform blocks."""
    o, s = make_pair((128, 64, 64), (1.0, 0.0, 0.0), nu=8 / 100, sphere=((31.0, 31.0, 31.0), 8.0), exitBC=True)
    _gate(o, s)


def test_torus_100_steps_gate():
    """configs[3] family (donut: torus SDF, axis along x, R/r = 4) at 128×64×64 with the exit BC."""
    o, s = make_pair((128, 64, 64), (1.0, 0.0, 0.0), nu=16 / 1000, torus=((32.0, 32.0, 32.0), 16.0, 4.0), exitBC=True)
    for name in ("mu0", "mu1", "V"):
        a, b = o.field(name), getattr(s.flow, {"mu0": "μ0", "mu1": "μ1", "V": "V"}[name])
        assert max_ulp(b, a) <= 2.0, name
    _gate(o, s, exact=False)


@pytest.mark.parametrize("lam", ["quick", "cds", "vanLeer"])
@pytest.mark.parametrize("case", ["per64", "wall64", "sphere64"])
def test_limiters_on_the_hot_flux_kernels(case, lam):
    """cds / vanLeer / quick through fm_conv4<λ> (periodic 64³, uniform mode) and fm_conv<λ> (walls, exit plane, body) on grids
    with several blocks per direction: bit-identical to the oracle after 3 steps."""
    if case == "per64":
        u0 = tgv3d_u0((66,) * 3, 64)
        u0[2] = (0.1 * np.roll(u0[0], 3, axis=0)).astype(F)
        o, s = make_pair((64, 64, 64), (0.0, 0.0, 0.0), nu=0.01, perdir=(1, 2, 3), u0=u0, lam=lam)
    elif case == "wall64":
        o, s = make_pair((64, 32, 32), (1.0, 0.0, 0.0), nu=0.02, perdir=(3,), exitBC=True, lam=lam)
        upload_same_u(o, s, seed=6)
    else:
        o, s = make_pair((64, 32, 32), (1.0, 0.0, 0.0), nu=0.05, sphere=((15.0, 15.0, 15.0), 4.0), lam=lam)
        upload_same_u(o, s, seed=7)
    from wl_b200 import lib as wlib
    for _ in range(3):
        o.mom_step()
        wlib.check(s.flow.L, s.flow.L.wl_mom_step(s.flow.h))
    assert np.array_equal(s.flow.u, o.field("u"))
    assert np.array_equal(s.flow.p, o.field("p"))
    assert list(np.asarray(o.iters)) == list(np.asarray(s.pois.n))


@pytest.mark.parametrize("periodic", [False, True])
def test_body_upload_without_update_rebuilds_the_hierarchy(periodic):
    """A host binding that uploads μ₀/μ₁/V after construction and never calls wl_update (the reference measures the body BEFORE it
    builds the Poisson, src/WaterLily.jl:104-105, and calls no update!): the library must rebuild D/iD/coarse L — and, on a fully
    periodic domain, leave constant-coefficient mode — before it steps."""
    import wl_b200 as wl
    from wl_b200 import lib as wlib
    kw = dict(perdir=(1, 2, 3), uBC=(0.0, 0.0, 0.0)) if periodic else dict(uBC=(1.0, 0.0, 0.0))
    sphere = ((15.0, 15.0, 15.0), 4.0)
    dims = (32, 32, 32)
    o, s = make_pair(dims, nu=0.05, sphere=sphere, measure=False, **kw)
    upload_same_u(o, s, seed=8)
    mu0, mu1, V, sigma = wl.measure_body(s.flow.N, wl.Sphere(*sphere), 1.0)
    s.flow.upload("mu0", mu0)
    s.flow.upload("mu1", mu1)
    s.flow.upload("V", V)
    wlib.check(s.flow.L, s.flow.L.wl_measure_bc(s.flow.h))  # BC!(μ₀), BC!(V): the tail of measure! — but NO wl_update
    for _ in range(2):
        o.mom_step()
        wlib.check(s.flow.L, s.flow.L.wl_mom_step(s.flow.h))
    flag = wlib.C.c_int(7)
    s.flow.L.wl_is_const_coeff(s.flow.h, wlib.C.byref(flag))
    assert flag.value == 0
    assert list(np.asarray(o.iters)) == list(np.asarray(s.pois.n))
    assert rel_l2(s.flow.u, o.field("u")) <= 1e-6 and rel_l2(s.flow.p, o.field("p")) <= 1e-5


def test_linf_and_solver_log_on_gpu():
    """L∞(p) and the solver log (src/Poisson.jl:190, src/MultiLevelPoisson.jl:111,116): rows (iter, r∞, r₂, ω) against the oracle's
    (iter, r₂, ω) rows, and the last r∞ against the oracle's L∞ of its final residual."""
    import wl_b200 as wl
    o, s = make_pair((32, 16, 16), (1.0, 0.0, 0.0), nu=0.04, sphere=((8.0, 7.0, 7.0), 3.0))
    wl.lib.check(s.flow.L, s.flow.L.wl_set_logging(s.flow.h, 1))
    o.mom_step()
    wl.lib.check(s.flow.L, s.flow.L.wl_mom_step(s.flow.h))
    lg, lo = s.pois.log, o.log
    assert lg.shape[0] == lo.shape[0] >= 4
    assert np.array_equal(lg[:, 0], lo[:, 0])
    assert np.allclose(lg[:, 2], lo[:, 1], rtol=1e-5) and np.allclose(lg[:, 3], lo[:, 2], rtol=1e-6)
    assert np.isclose(lg[-1, 1], o.L.wlo_pois_Linf(o.h), rtol=1e-6)
    wl.lib.check(s.flow.L, s.flow.L.wl_set_logging(s.flow.h, 0))
    wl.lib.check(s.flow.L, s.flow.L.wl_mom_step(s.flow.h))
    assert s.pois.log.shape[0] == lg.shape[0]  # nothing is appended while logging is off


def _two_rank_worker(rank, world, port, case, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import wl_b200 as wl
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(wl.dist_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    idb = bytes(idt.cpu().numpy().tobytes())
    n, steps = 64, 5
    if case == "tgv":
        dims = (n, n, n)
        u0g = tgv3d_u0((n + 2,) * 3, n)
        kw = dict(ν=float(F(1 / (2 * np.pi / n * 1600))), perdir=(1, 2, 3))
        uBC, body = (0.0, 0.0, 0.0), None
    else:
        dims = (2 * n, n, n)
        u0g = None
        kw = dict(ν=n / 8 / 100.0, exitBC=True)
        uBC, body = (1.0, 0.0, 0.0), wl.Sphere((n / 2 - 1,) * 3, n / 8)
    nzl = dims[2] // world
    u0f = None
    if u0g is not None:
        sl = u0g[:, rank * nzl: rank * nzl + nzl + 2]
        u0f = lambda i, x: sl[i]  # noqa: E731
    sim = wl.Simulation(dims, uBC, float(n), u0=u0f, body=body, device=rank, dist=(rank, world, idb), **kw)
    wl.lib.check(sim.flow.L, sim.flow.L.wl_sim_step_n(sim.flow.h, steps))
    u, p, dt, its = sim.flow.u, sim.flow.p, sim.flow.Δt, sim.pois.n
    ok = None
    if rank == 0:
        u0f1 = (lambda i, x: u0g[i]) if u0g is not None else None
        ref = wl.Simulation(dims, uBC, float(n), u0=u0f1, body=body, device=rank, **kw)
        wl.lib.check(ref.flow.L, ref.flow.L.wl_sim_step_n(ref.flow.h, steps))
        ur, pr = ref.flow.u[:, 0: nzl + 2], ref.flow.p[0: nzl + 2]
        ok = (bool(np.array_equal(u[:, 1:-1], ur[:, 1:-1])), bool(np.array_equal(p[1:-1], pr[1:-1])),
              bool(np.array_equal(dt, ref.flow.Δt)), list(its) == list(ref.pois.n))
    dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["tgv", "sphere"])
def test_two_rank_slab_run_is_bit_identical_to_one_gpu(case):
    """§8e parity: a 2-rank z-slab run (P2P halos, all-reduces, replicated coarse levels) equals the single-GPU run bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 150
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for pr in procs:
        pr.join(60)
    assert res[0] == (True, True, True, True), res


# ---- §8f-1: measure! on the device (parametrised bodies, set operations, moving bodies with remeasure=true) -----------------
BODY_CASES = {
    "sphere3d": (dict(dims=(32, 24, 16), uBC=(1.0, 0.0, 0.0), nu=0.02), lambda wl: wl.Sphere((12.0, 11.0, 8.0), 5.0), 0.0),
    "circle2d": (dict(dims=(48, 32), uBC=(1.0, 0.0), nu=0.02), lambda wl: wl.Sphere((15.5, 16.0), 6.0), 0.0),
    "torus": (dict(dims=(32, 32, 32), uBC=(1.0, 0.0, 0.0), nu=0.02), lambda wl: wl.Torus((14.0, 16.0, 15.5), 8.0, 2.5), 0.0),
    "moving_sphere": (dict(dims=(32, 24, 16), uBC=(0.0, 0.0, 0.0), nu=0.02), lambda wl: wl.Sphere((10.0, 11.0, 8.0), 4.0, velocity=(0.5, 0.1, 0.0)), 3.0),
    "two_spheres_minus_hole": (dict(dims=(48, 24, 24), uBC=(1.0, 0.0, 0.0), nu=0.02),
                               lambda wl: (wl.Sphere((14.0, 12.0, 12.0), 5.0) | wl.Sphere((20.0, 12.0, 12.0), 4.0)) - wl.Sphere((17.0, 12.0, 12.0), 2.0), 0.0),
    "lens": (dict(dims=(32, 24, 24), uBC=(1.0, 0.0, 0.0), nu=0.02),
             lambda wl: wl.Sphere((13.0, 12.0, 12.0), 6.0) & wl.Sphere((18.0, 12.0, 12.0), 6.0), 0.0),
}


@pytest.mark.parametrize("name", list(BODY_CASES))
def test_device_measure_matches_oracle(name):
    """wl_set_body + wl_measure (k_measure) against the oracle's measure! for every primitive, the set operations and a moving
    body at t > 0: σ, μ₀, μ₁, V to 2 ulp (sinpi/cospi libraries differ in the last bit of the double), the hierarchy to 4 ulp."""
    import oracle
    import wl_b200 as wl
    cfg, mk, t = BODY_CASES[name]
    body = mk(wl)
    o = oracle.OracleSim(cfg["dims"], cfg["uBC"], nu=cfg["nu"])
    o.measure_prims(body.prims(), 1.0, t)
    o.init_pois()
    s = wl.Simulation(cfg["dims"], cfg["uBC"], 1.0, ν=cfg["nu"], body=body)
    assert s.device_body
    wl.measure(s, t=t)
    inner = tuple(slice(1, -1) for _ in cfg["dims"])
    for nm, attr in (("mu0", "μ0"), ("mu1", "μ1"), ("V", "V")):
        assert max_ulp(getattr(s.flow, attr), o.field(nm)) <= 2.0, nm
    assert max_ulp(s.flow.σ[inner], o.field("sigma")[inner]) <= 2.0
    for lvl in range(s.pois.nlevels):
        for arr in ("L", "D", "iD"):
            assert max_ulp(s.pois.level(lvl, arr), o.level_field(lvl, arr)) <= 4.0, (lvl, arr)
    # and the host (NumPy) measurement of the same body, for the static single primitives
    if t == 0.0 and len(body.prims()) == 1:
        h = wl.Simulation(cfg["dims"], cfg["uBC"], 1.0, ν=cfg["nu"], body=body, host_measure=True)
        assert not h.device_body
        assert max_ulp(h.flow.μ0, s.flow.μ0) <= 2.0 and max_ulp(h.flow.μ1, s.flow.μ1) <= 2.0


def test_moving_sphere_remeasure_parity():
    """sim_step!(sim; remeasure=true) (src/WaterLily.jl:136-149) with a sphere translating through quiescent fluid: every step
    measure!(sim, t=sum(Δt)) + update!(pois) + mom_step! inside the library, against the oracle doing the same."""
    import oracle
    import wl_b200 as wl
    dims, nu = (48, 32, 32), 0.05
    body = wl.Sphere((14.0, 15.5, 16.0), 5.0, velocity=(0.4, 0.0, 0.0))
    o = oracle.OracleSim(dims, (0.0, 0.0, 0.0), nu=nu)
    o.measure_prims(body.prims(), 1.0, 0.0)
    o.init_pois()
    s = wl.Simulation(dims, (0.0, 0.0, 0.0), 10.0, U=0.4, ν=nu, body=body)
    for _ in range(12):
        o.measure_prims(body.prims(), 1.0, o.time_next())
        o.update()
        o.mom_step()
        wl.sim_step(s, remeasure=True)
    assert np.abs(np.asarray(o.iters, int) - np.asarray(s.pois.n, int)).max() <= 1, (list(o.iters), list(s.pois.n))
    eu, ep = rel_l2(s.flow.u, o.field("u")), rel_l2(s.flow.p, o.field("p"))
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)
    assert np.allclose(o.dt, s.flow.Δt, rtol=1e-6)
    # the batched loop does the same
    s2 = wl.Simulation(dims, (0.0, 0.0, 0.0), 10.0, U=0.4, ν=nu, body=body)
    wl.lib.check(s2.flow.L, s2.flow.L.wl_set_remeasure(s2.flow.h, 1))
    wl.lib.check(s2.flow.L, s2.flow.L.wl_sim_step_n(s2.flow.h, 12))
    assert np.array_equal(s2.flow.u, s.flow.u) and np.array_equal(s2.flow.p, s.flow.p)


# ---- §8f-2/3/4: forces and moments, MeanFlow + checkpoint, enumerated forcings -------------------------------------------------
def test_forces_and_moments_match_oracle_and_known_answers():
    """pressure_force / viscous_force / pressure_moment / viscous_moment as one fused device reduction (src/Metrics.jl:111-190):
    against the oracle after a few steps of a sphere wake, and the reference's own known answers (test/test_metrics.jl:36-66):
    hydrostatic pressure p = y gives force/(πR²) ≈ (0,1) and no moment; a fluid at rest exerts no viscous force."""
    import oracle
    import wl_b200 as wl
    body = wl.Sphere((15.0, 15.5, 16.0), 5.0)
    o, s = make_pair((48, 32, 32), (1.0, 0.0, 0.0), nu=0.05, sphere=((15.0, 15.5, 16.0), 5.0))
    for _ in range(5):
        o.mom_step()
        wl.sim_step(s)
    x0 = (15.0, 15.5, 16.0)
    fo = o.body_forces(body.prims(), x0)
    fg = np.stack([wl.pressure_force(s), wl.viscous_force(s), wl.pressure_moment(x0, s), wl.viscous_moment(x0, s)])
    scale = np.abs(fo[0]).max()
    assert np.allclose(fg, fo, rtol=1e-5, atol=1e-6 * scale), (fg, fo)
    assert np.allclose(wl.total_force(s), fo[0] + fo[1], rtol=1e-5, atol=1e-6 * scale)
    assert abs(fg[0][0]) > 1.0 and abs(fg[0][1]) < 0.05 * abs(fg[0][0])  # a streamwise force on a sphere in a +x stream, (almost) no lift
    # known answers, 2-D (test/test_metrics.jl:36-40, 56, 65)
    N = 34  # ghost-padded size (the reference test uses bare N×N arrays; here the arrays belong to a 32² simulation)
    R = 8
    c = wl.Sphere((N / 2, N / 2), R)
    s2 = wl.Simulation((N - 2, N - 2), (0.0, 0.0), 1.0, body=c)
    y = (np.arange(1, N + 1, dtype=F) - F(1.5))[:, None] * np.ones((1, N), F)
    s2.flow.upload("p", y)
    f = wl.pressure_force(s2) / (np.pi * R ** 2)
    assert np.abs(f - np.array([0.0, 1.0])).sum() < 2e-3, f
    assert abs(wl.pressure_moment((N / 2, N / 2), s2)[0]) < 1e-3 * np.pi * R ** 2
    assert np.all(wl.viscous_force(s2) == 0.0) and np.all(wl.viscous_moment((N / 2, N / 2), s2) == 0.0)


def test_meanflow_and_checkpoint():
    """MeanFlow on the device (src/Metrics.jl:205-261) against a NumPy restatement of update!, and save!/load! of a flow
    (ext/WaterLilyJLD2Ext.jl:11-50): a restarted run continues the saved one bit for bit."""
    import os
    import tempfile
    import wl_b200 as wl
    s = wl.Simulation((32, 24), (1.0, 0.0), 8.0, ν=0.02, body=wl.Sphere((10.0, 12.0), 3.0))
    mf = wl.MeanFlow(s.flow, uu_stats=True)
    assert np.all(mf.P == 0) and np.all(mf.U == 0) and np.all(mf.UU == 0) and list(mf.t) == [0.0]
    P = np.zeros_like(s.flow.p)
    U = np.zeros_like(s.flow.u)
    UU = np.zeros((2, 2) + P.shape, F)
    t = [F(0.0)]
    for k in range(6):
        for _ in range(2):
            wl.sim_step(s)
        mf.update()
        dt = F(F(s.flow.time()) - t[-1])
        eps = F(dt / F(F(dt + F(t[-1] - t[0])) + np.finfo(F).eps))
        if len(t) == 1:
            eps = F(1.0)
        p, u = s.flow.p, s.flow.u
        P = (eps * p + (F(1) - eps) * P).astype(F)
        U = (eps * u + (F(1) - eps) * U).astype(F)
        for i in range(2):
            for j in range(2):
                UU[j, i] = (eps * (u[i] * u[j]) + (F(1) - eps) * UU[j, i]).astype(F)
        t.append(F(t[-1] + dt))
        if k == 0:  # the first update takes the instantaneous field
            assert np.array_equal(mf.U, u) and np.array_equal(mf.P, p)
    assert max_ulp(mf.P, P) <= 1.0 and max_ulp(mf.U, U) <= 1.0 and max_ulp(mf.UU, UU) <= 1.0
    assert np.allclose(mf.t, np.array(t, F), rtol=1e-6)
    tau = mf.uu()
    assert np.allclose(tau[1, 0], mf.UU[1, 0] - mf.U[0] * mf.U[1], atol=float(np.sqrt(np.finfo(F).eps)))
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "chk.npz")
        wl.save(fn, s)
        wl.save(os.path.join(d, "mf.npz"), mf)
        for _ in range(3):
            wl.sim_step(s)
        r = wl.Simulation((32, 24), (1.0, 0.0), 8.0, ν=0.02, body=wl.Sphere((10.0, 12.0), 3.0))
        wl.load(r, fn)
        for _ in range(3):
            wl.sim_step(r)
        assert np.array_equal(r.flow.u, s.flow.u) and np.array_equal(r.flow.p, s.flow.p)
        assert np.array_equal(np.asarray(r.flow.Δt), np.asarray(s.flow.Δt))
        mf2 = wl.MeanFlow(r.flow, uu_stats=True)
        wl.load(mf2, os.path.join(d, "mf.npz"))
        assert np.array_equal(mf2.U, mf.U) and np.array_equal(mf2.UU, mf.UU) and np.array_equal(mf2.t, mf.t)
    mf.copy_to()
    assert np.array_equal(s.flow.u, mf.U) and np.array_equal(s.flow.p, mf.P)
    mf.reset()
    assert np.all(mf.U == 0) and list(mf.t) == [0.0]


FORCING_CASES = {
    # the reference's constant-jerk test (test/test_flow.jl:111-121, helper.jl:26-33): periodic in x, g = (t·jerk, 0)
    "jerk_2d": dict(dims=(8, 8), uBC=(float(np.sqrt(8.0)), 0.0), nu=0.001, dt0=0.001, perdir=(1,), g1=(4.0, 0.0)),
    "gravity_body_3d": dict(dims=(32, 16, 16), uBC=(1.0, 0.0, 0.0), nu=0.05, sphere=((10.0, 8.0, 8.0), 3.0), g0=(0.0, -0.2, 0.0), g1=(0.05, 0.0, 0.0)),
    "accelerating_inflow_3d": dict(dims=(32, 16, 16), uBC=(1.0, 0.0, 0.0), nu=0.05, sphere=((10.0, 8.0, 8.0), 3.0), U1=(0.1, 0.0, 0.0), U2=(0.02, 0.0, 0.0)),
    "periodic_uniform_mode": dict(dims=(64, 64, 64), uBC=(0.0, 0.0, 0.0), nu=0.01, perdir=(1, 2, 3), g0=(0.0, 0.0, -0.1), g1=(0.02, 0.0, 0.0), tgv=True),
}


@pytest.mark.parametrize("name", list(FORCING_CASES))
def test_enumerated_forcings_match_oracle(name):
    """accelerate! with g(t) = g0 + g1·t and a time-dependent uniform uBC(t) = U0 + U1·t + ½U2·t² (src/Flow.jl:64-73,
    src/core.jl:201-219) through k_conv_bdim1 (2-D), fm_conv + k_f_lowghost (walls, body) and fm_conv4 (uniform mode)."""
    import oracle
    import wl_b200 as wl
    c = dict(FORCING_CASES[name])
    g0, g1, U1, U2 = (c.pop(k, None) for k in ("g0", "g1", "U1", "U2"))
    tgv = c.pop("tgv", False)
    sphere = c.pop("sphere", None)
    dims, uBC = c.pop("dims"), c.pop("uBC")
    u0 = None
    if tgv:
        u0 = tgv3d_u0(tuple(d + 2 for d in dims), dims[0])
    o = oracle.OracleSim(dims, uBC, nu=c["nu"], dt0=c.get("dt0", 0.25), perdir=c.get("perdir", ()), u0=u0)
    if sphere:
        o.measure_sphere(*sphere)
    o.init_pois()
    o.set_forcing(g0 or (0, 0, 0), g1 or (0, 0, 0), U1 or (0, 0, 0), U2 or (0, 0, 0))
    ubc = wl.TimeBC(uBC, U1, U2) if (U1 or U2) else uBC
    s = wl.Simulation(dims, ubc, float(dims[0]), U=1.0, ν=c["nu"], Δt=c.get("dt0", 0.25), perdir=c.get("perdir", ()),
                      g=wl.Forcing(g0, g1) if (g0 or g1) else None, body=wl.Sphere(*sphere) if sphere else None,
                      u0=(lambda i, x: u0[i]) if tgv else None)
    nsteps = 30 if name == "jerk_2d" else 6
    for _ in range(nsteps):
        o.mom_step()
        wl.sim_step(s)
    assert np.abs(np.asarray(o.iters, int) - np.asarray(s.pois.n, int)).max() <= 1
    eu, ep = rel_l2(s.flow.u, o.field("u")), rel_l2(s.flow.p, o.field("p"))
    assert eu <= 1e-5 and ep <= 1e-5, (eu, ep)
    assert np.allclose(o.dt, s.flow.Δt, rtol=1e-6)
    if name == "jerk_2d":  # exact: u_x = U0 + ½·jerk·t² (test/test_flow.jl:116-121)
        from test_oracle_golden import L2in
        u = s.flow.u
        uf = F(np.sqrt(8.0) + 0.5 * 4.0 * s.flow.time() ** 2)
        assert L2in(u[0] - uf) < 1e-4 * max(1.0, float(uf)) ** 2 and L2in(u[1]) < 1e-4


SGS_CASES = {
    # walls + body + exit plane (general kernels), fully periodic TGV (leaves uniform mode for the udf), 2-D with a periodic direction
    "sphere_exit_3d": dict(dims=(32, 16, 16), uBC=(1.0, 0.0, 0.0), nu=0.01, sphere=((10.0, 8.0, 8.0), 3.0), exitBC=True, Cs=0.2, Delta=1.0),
    "tgv_periodic_3d": dict(dims=(64, 64, 64), uBC=(0.0, 0.0, 0.0), nu=1e-4, perdir=(1, 2, 3), tgv=True, Cs=0.17, Delta=1.0),
    "channel_2d": dict(dims=(48, 32), uBC=(1.0, 0.0), nu=0.002, perdir=(1,), rand=True, Cs=0.3, Delta=1.5),
}


@pytest.mark.parametrize("name", list(SGS_CASES))
def test_sgs_udf_matches_oracle(name):
    """sim_step!(sim; udf=sgs!, νₜ=smagorinsky, S, Cs, Δ) (src/util.jl:46-76) as the library's built-in (wl_set_sgs): bit for bit
    against the oracle's restatement over 6 steps, and a different flow from the one without the model."""
    import oracle
    import wl_b200 as wl
    c = dict(SGS_CASES[name])
    dims, uBC = c["dims"], c["uBC"]
    u0 = None
    if c.get("tgv"):
        u0 = tgv3d_u0(tuple(d + 2 for d in dims), dims[0])
    if c.get("rand"):
        u0 = smooth_field(tuple(d + 2 for d in dims), len(dims), seed=3, amp=0.3)
        u0[0] += 1.0
    sphere = c.get("sphere")
    o = oracle.OracleSim(dims, uBC, nu=c["nu"], perdir=c.get("perdir", ()), exitBC=c.get("exitBC", False), u0=u0)
    if sphere:
        o.measure_sphere(*sphere)
    o.init_pois()
    o.set_sgs(c["Cs"], c["Delta"])
    u0f = (lambda i, x: u0[i]) if u0 is not None else None
    mk = lambda: wl.Simulation(dims, uBC, float(dims[0]), U=1.0, ν=c["nu"], perdir=c.get("perdir", ()), exitBC=c.get("exitBC", False),
                               body=wl.Sphere(*sphere) if sphere else None, u0=u0f)
    s, plain = mk(), mk()
    for _ in range(6):
        o.mom_step()
        wl.sim_step(s, udf=wl.sgs, νₜ=wl.smagorinsky, Cs=c["Cs"], Δ=c["Delta"])
        wl.sim_step(plain)
    assert list(np.asarray(o.iters, int)) == list(np.asarray(s.pois.n, int))
    assert np.array_equal(s.flow.u, o.field("u")) and np.array_equal(s.flow.p, o.field("p"))
    assert np.array_equal(np.asarray(o.dt, F), np.asarray(s.flow.Δt, F))
    assert rel_l2(plain.flow.u, s.flow.u) > 1e-5  # the model does something
    # switching the udf off returns to the plain kernels (uniform mode again on the periodic case)
    wl.sim_step(s)
    o.set_sgs(0.0, 0.0)
    o.mom_step()
    assert np.array_equal(s.flow.u, o.field("u"))
    with pytest.raises(wl.WLError):
        wl.sim_step(s, udf=lambda flow, t: None)


def test_large_host_transfers_round_trip():
    """Fields of ≥ 64 MB per component take the pinned-ring path between the device and ordinary (pageable) host memory
    (copy_in / copy_out, wl_b200.cu): a 256³ velocity and pressure field must survive upload → download bit for bit, from pageable and
    from pinned host memory, and a second handle (memory from the chunk pool of the first) must start from clean state."""
    import torch
    import wl_b200 as wl
    dims = (256, 256, 256)
    rng = np.random.default_rng(11)
    shape = (3,) + tuple(d + 2 for d in reversed(dims))
    a = rng.standard_normal(shape, dtype=np.float32)
    fl = wl.Flow(dims, (0.0, 0.0, 0.0), perdir=(1, 2, 3))
    fl.upload("u0", a)                                   # pageable source → ring
    assert np.array_equal(fl.u0, a)                      # pageable destination → ring
    pinned = torch.from_numpy(a[0].copy()).pin_memory().numpy()
    fl.upload("sigma", pinned)                           # pinned source → direct DMA
    assert np.array_equal(fl.σ, a[0])
    for i in range(3):
        fl.upload_component("u0", i, a[2 - i])
    assert np.array_equal(fl.u0, a[::-1])
    fl.close()
    fl2 = wl.Flow(dims, (0.0, 0.0, 0.0), perdir=(1, 2, 3))  # chunks come back from the pool: everything is zeroed again
    assert not fl2.σ.any() and not fl2.f.any()
    fl2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["sphere_walls", "torus_periodic_y", "tgv_general_coeff"])
def test_tiny_level_kernels_equal_the_cooperative_kernel(case):
    """The one-block shared-memory kernels for the innermost multigrid levels (k_tiny_gen with bodies / walls / semi-coarsened
    levels, k_tiny_uni in uniform mode) against the same levels run inside k_small_levels (WL_FLAG_NO_TINY), and the host reading
    the residual norms from the mapped mirror against copy + synchronise (WL_FLAG_NO_FAST_READ), programmatic dependent launch
    against ordinary stream order (WL_FLAG_NO_PDL): bit for bit over 8 steps."""
    import wl_b200 as wl
    outs = []
    for flags in (0, wl.lib.FLAGS["no_tiny"] | wl.lib.FLAGS["no_fast_read"] | wl.lib.FLAGS["no_pdl"]):
        if case == "sphere_walls":
            s = wl.Simulation((128, 64, 64), (1.0, 0.0, 0.0), 16.0, ν=0.01, body=wl.Sphere((31.0, 31.0, 31.0), 8.0), exitBC=True, flags=flags)
        elif case == "torus_periodic_y":
            s = wl.Simulation((96, 64, 32), (1.0, 0.0, 0.0), 16.0, ν=0.02, perdir=(2,), body=wl.Torus((30.0, 32.0, 16.0), 9.0, 3.0), flags=flags)
        else:
            n = 64
            u0 = tgv3d_u0((n + 2,) * 3, n)
            s = wl.Simulation((n,) * 3, (0.0, 0.0, 0.0), float(n), ν=0.001, perdir=(1, 2, 3), u0=lambda i, x: u0[i],
                              flags=flags | wl.lib.FLAGS["general_coeff"])
        wl.lib.check(s.flow.L, s.flow.L.wl_sim_step_n(s.flow.h, 8))
        outs.append((s.flow.u.copy(), s.flow.p.copy(), list(s.pois.n), np.asarray(s.flow.Δt).copy()))
        s.close()
    a, b = outs
    assert a[2] == b[2] and np.array_equal(a[3], b[3])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_accelerating_circle_added_mass_on_gpu():
    """test/test_flow.jl:161-173 through the C ABI: uBC(t) = (t, 0) as TimeBC, a circle of radius 32 in a 1024² box measured on the
    device; pressure_force/(πL²) ≈ (−1, 0) ± 0.04 after one step, peak velocity ≈ 2U, n ≤ 2 — and the oracle's bits."""
    import math
    import oracle
    import wl_b200 as wl
    radius, H = 32, 16
    n = radius * 2 * H
    c = (float(H * radius), float(H * radius))
    s = wl.Simulation((n, n), wl.TimeBC((0.0, 0.0), U1=(1.0, 0.0)), float(radius), U=1.0, body=wl.Sphere(c, float(radius)))
    o = oracle.OracleSim((n, n), (0.0, 0.0))
    o.measure_sphere(c, float(radius))
    o.init_pois()
    o.set_forcing(U1=(1.0, 0.0))
    wl.sim_step(s)
    o.mom_step()
    pf = np.asarray(wl.pressure_force(s), np.float64) / (math.pi * radius ** 2)
    assert abs(pf[0] + 1.0) < 0.04 and abs(pf[1]) < 0.04, pf
    u = s.flow.u
    assert float(u.max()) / float(u[0][1, 1]) > 1.91
    assert rel_l2(u, o.field("u")) <= 1e-5 and rel_l2(s.flow.p, o.field("p")) <= 1e-5
    for _ in range(3):
        wl.sim_step(s)
        o.mom_step()
    assert all(int(k) <= 2 for k in s.pois.n), list(s.pois.n)
    assert list(np.asarray(s.pois.n, int)) == list(np.asarray(o.iters, int))


@pytest.mark.parametrize("exitBC", [True, False])
def test_remeasure_body_moving_with_the_stream_on_gpu(exitBC):
    """test/test_simulation.jl:20-25 through the C ABI: sim_step!(sim) with remeasure=true and a circle that translates with the free
    stream leaves u[:, radius, 1] ≈ 1; the fields equal the oracle's."""
    import math
    import oracle
    import wl_b200 as wl
    radius = 8
    body = wl.Sphere((2.0 * radius, 2.0 * radius), float(radius), velocity=(1.0, 0.0))
    s = wl.Simulation((4 * radius, 4 * radius), (1.0, 0.0), float(radius), ν=radius / 250, body=body, exitBC=exitBC)
    o = oracle.OracleSim((4 * radius, 4 * radius), (1.0, 0.0), nu=radius / 250, exitBC=exitBC)
    o.measure_prims(body.prims(), 1.0, 0.0)
    o.init_pois()
    o.measure_prims(body.prims(), 1.0, o.time_next())
    o.update()
    o.mom_step()
    wl.sim_step(s, remeasure=True)
    row = s.flow.u[0][radius - 1, :]
    assert np.all(np.abs(row - 1.0) <= math.sqrt(np.finfo(np.float32).eps) * np.maximum(np.abs(row), 1.0)), row
    assert rel_l2(s.flow.u, o.field("u")) <= 1e-5 and rel_l2(s.flow.p, o.field("p")) <= 1e-5
    assert list(np.asarray(s.pois.n, int)) == list(np.asarray(o.iters, int))
