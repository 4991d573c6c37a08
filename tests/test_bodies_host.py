"""Host-side body measurement (set-up only: src/Body.jl:28-60, src/AutoBody.jl:29-37) for the bodies the BASELINE configurations use.
The sphere is checked against the oracle in test_gpu_parity.py; here: the torus of configs[3] (WaterLily-Examples ThreeD_Donut recipe),
whose gradient is analytic in body.py where the reference differentiates the SDF with ForwardDiff."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waterlily_loader import wl  # noqa: E402

F = np.float32


def test_torus_sdf_and_gradient():
    body = wl.Torus((16.0, 16.0, 16.0), 8.0, 2.0)
    rng = np.random.default_rng(0)
    pts = (rng.uniform(4, 28, size=(3, 2000))).astype(F)
    x = [pts[0], pts[1], pts[2]]
    d = body._sdf(x)
    g = body._g(x)
    # points on the tube's centre circle are at distance -r, points on the axis at sqrt(R²+x²) - r
    on_circle = [np.array([16.0], F), np.array([16.0 + 8.0], F), np.array([16.0], F)]
    assert abs(float(body._sdf(on_circle)[0]) + 2.0) < 1e-6
    on_axis = [np.array([19.0], F), np.array([16.0], F), np.array([16.0], F)]
    assert abs(float(body._sdf(on_axis)[0]) - (np.sqrt(64.0 + 9.0) - 2.0)) < 1e-5
    # |∇sdf| = 1 and the analytic gradient matches central differences of the SDF (what ForwardDiff returns up to rounding)
    nrm = np.sqrt(sum(np.asarray(gi, np.float64) ** 2 for gi in g))
    ok = np.isfinite(nrm)
    assert np.allclose(nrm[ok], 1.0, atol=1e-5)
    h = 1e-2
    for k in range(3):
        xp = [xi.astype(np.float64) for xi in x]
        xm = [xi.astype(np.float64) for xi in x]
        xp[k] = xp[k] + h
        xm[k] = xm[k] - h
        fd = (body._sdf([a.astype(F) for a in xp]).astype(np.float64) - body._sdf([a.astype(F) for a in xm]).astype(np.float64)) / (2 * h)
        far = np.abs(d) > 0.2  # away from the kink of the distance field on the centre circle
        assert np.allclose(np.asarray(g[k], np.float64)[far & ok], fd[far & ok], atol=5e-3)


def test_torus_measure_body_fields():
    N = (34, 34, 34)
    body = wl.Torus((16.0, 16.0, 16.0), 8.0, 3.0)
    mu0, mu1, V, sig = wl.measure_body(N, body, 1.0)
    assert mu0.shape == (3, 34, 34, 34) and mu1.shape == (9, 34, 34, 34) and V.shape == (3, 34, 34, 34)
    assert mu0.min() >= 0.0 and mu0.max() <= 1.0
    assert np.all(V == 0)  # static body
    # deep inside the tube (a cell centre on the centre circle: x=16 ↔ index 17.5, so take the nearest cells) μ₀ = 0; far away μ₀ = 1, μ₁ = 0
    k, j, i = 17, 25, 17  # C order (z, y, x): y − 16 ≈ 8 = R
    assert float(mu0[:, k, j, i].max()) == 0.0
    assert np.all(mu0[:, 2, 2, 2] == 1.0) and np.all(mu1[:, 2, 2, 2] == 0.0)
    # the smoothed band is thin: between 2 and 12 % of the cells of this box are neither 0 nor 1
    band = np.mean((mu0[0] > 0) & (mu0[0] < 1))
    assert 0.005 < band < 0.2


def test_oracle_prims_equal_the_single_body_paths_and_set_operation_identities():
    """The oracle's parametrised-body measure! (translation map + set operations, src/Body.jl:88-107) against its own sphere /
    torus paths, and identities the reference's tuple min/max imply: a ∪ a = a, a ∩ a = a, (a ∪ b) with b far away = a near a,
    a moving sphere at time t = the static sphere at the displaced centre with V = velocity in the band."""
    import oracle
    dims = (24, 16, 16)

    def fields(o):
        return {k: o.field(k).copy() for k in ("mu0", "mu1", "V", "sigma")}

    sph = dict(kind=0, op=0, center=[9.0, 7.5, 8.0], R=3.5, r=0.0, vel=[0.0, 0.0, 0.0])
    a = oracle.OracleSim(dims, (1.0, 0.0, 0.0))
    a.measure_sphere((9.0, 7.5, 8.0), 3.5)
    fa = fields(a)
    b = oracle.OracleSim(dims, (1.0, 0.0, 0.0))
    b.measure_prims([sph])
    fb = fields(b)
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    tor = dict(kind=1, op=0, center=[10.0, 8.0, 8.0], R=4.0, r=1.5, vel=[0.0, 0.0, 0.0])
    a.measure_torus((10.0, 8.0, 8.0), 4.0, 1.5)
    b.measure_prims([tor])
    for k in fa:
        assert np.array_equal(a.field(k), b.field(k)), k
    # a ∪ a, a ∩ a: the set operations go through measure(…)[1] for σ (normalised distance), so compare μ₀, μ₁, V only
    for op in (0, 1):
        b.measure_prims([sph, dict(sph, op=op)])
        for k in ("mu0", "mu1", "V"):
            assert np.array_equal(fb[k], b.field(k)), (op, k)
    far = dict(kind=0, op=0, center=[60.0, 7.5, 8.0], R=2.0, r=0.0, vel=[0.0, 0.0, 0.0])
    b.measure_prims([sph, far])
    assert np.array_equal(fb["mu0"], b.field("mu0")) and np.array_equal(fb["mu1"], b.field("mu1"))
    # a − b with b covering a: nothing left (μ₀ = 1 everywhere inside the domain)
    big = dict(kind=0, op=2, center=[9.0, 7.5, 8.0], R=6.0, r=0.0, vel=[0.0, 0.0, 0.0])
    b.measure_prims([sph, big])
    nobody = oracle.OracleSim(dims, (1.0, 0.0, 0.0))  # μ₀ = 1 with BC!(μ₀,0) applied (src/Flow.jl:145)
    assert np.array_equal(b.field("mu0"), nobody.field("mu0")) and np.all(b.field("mu1") == 0.0)
    # translation map: centre 9 + 0.5·2 = 10
    mov = dict(sph, vel=[0.5, 0.0, 0.0])
    b.measure_prims([mov], t=2.0)
    a.measure_sphere((10.0, 7.5, 8.0), 3.5)
    assert np.array_equal(a.field("mu0"), b.field("mu0")) and np.array_equal(a.field("mu1"), b.field("mu1"))
    V = b.field("V")
    band = b.field("sigma") ** 2 < 9.0
    # V = velocity on the faces measured inside the band (a face at the band's edge can lie beyond fastd²: 0 there), 0 elsewhere
    band[0], band[-1], band[:, 0], band[:, -1], band[:, :, 0], band[:, :, -1] = (False,) * 6  # σ is only defined inside
    vb = V[0][band]
    assert set(np.unique(vb)) <= {np.float32(0.0), np.float32(0.5)} and (vb == np.float32(0.5)).mean() > 0.9
    assert np.all(V[1] == 0) and np.all(V[0][1:-1, 1:-1, 1:-1][~band[1:-1, 1:-1, 1:-1]] == 0)
