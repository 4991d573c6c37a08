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
