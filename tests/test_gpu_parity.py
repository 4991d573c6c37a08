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
def test_mom_step_two_steps(name):
    sphere = None
    cfg = CASES[name]
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


def test_pcg_smoother_and_single_level_step():
    o, s = make_pair((32, 32), (1.0, 0.0), nu=0.05, sphere=((12.0, 16.0), 4.0), pois="single")
    for _ in range(3):
        o.mom_step()
        s.flow.L.wl_mom_step(s.flow.h)
    assert np.abs(np.asarray(o.iters, int) - np.asarray(s.pois.n, int)).max() <= 1
    assert rel_l2(s.flow.u, o.field("u")) < 1e-4


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
def test_fused_upstroke_matches_oracle(dims, nz, monkeypatch):
    """f_vsmooth (prolongation + GaussSeidelRB! + increments in one pass, wl_vsmooth.cuh) and fm_conv4 on grids large enough to
    take those paths, with several tiles per direction and (nz=2) several z chunks: bit-identical to the oracle."""
    if nz:
        monkeypatch.setenv("WL_VS_NZ", str(nz))
    o, s = _tgv_pair(dims)
    for _ in range(4):
        o.mom_step()
        s.flow.L.wl_mom_step(s.flow.h)
    assert np.array_equal(s.flow.u, o.field("u"))
    assert np.array_equal(s.flow.p, o.field("p"))
    assert list(np.asarray(o.iters)) == list(np.asarray(s.pois.n))
    assert np.array_equal(np.asarray(o.dt, F), np.asarray(s.flow.Δt, F))


@pytest.mark.gpu
def test_fused_kernels_equal_unfused_kernels(monkeypatch):
    """128³ (two levels take f_vsmooth): the fused uniform-mode kernels against the separate march kernels, bit for bit."""
    import wl_b200 as wl
    outs = []
    for flags in ({"WL_VSMOOTH": "1", "WL_CONV4": "1"}, {"WL_VSMOOTH": "0", "WL_CONV4": "0"}):
        for k, v in flags.items():
            monkeypatch.setenv(k, v)
        n = 128
        u0 = tgv3d_u0((n + 2,) * 3, n)
        u0[2] = (0.1 * np.roll(u0[0], 3, axis=0)).astype(F)
        nu = float(F(1 / (2 * np.pi / n * 1600)))
        s = wl.Simulation((n,) * 3, (0.0, 0.0, 0.0), float(n), ν=nu, perdir=(1, 2, 3), u0=lambda i, x: u0[i])
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
