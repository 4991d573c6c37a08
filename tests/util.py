"""Shared helpers for the parity tests: build the same case on the oracle and on the B200 library."""
import numpy as np

F = np.float32


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / (den if den > 0 else 1.0))


def max_ulp(a, b):
    """largest difference in units of the last place of the larger magnitude (Float32)"""
    a = np.asarray(a, F)
    b = np.asarray(b, F)
    sp = np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(F))
    with np.errstate(invalid="ignore"):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64)) / sp
    return float(np.nanmax(d)) if d.size else 0.0


def smooth_field(N, D, seed=0, amp=1.0):
    """A deterministic, smooth-plus-noise vector field on a ghost-padded grid, C order (D, ..., N2, N1)."""
    rng = np.random.default_rng(seed)
    shape = tuple(reversed(N))
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    u = np.zeros((D,) + shape, F)
    for i in range(D):
        v = np.zeros(shape)
        for k, gk in enumerate(grids):
            v += np.sin(2 * np.pi * (k + 1 + i) * gk / max(shape[k] - 2, 1) + 0.3 * i)
        u[i] = (amp * (v / D + 0.2 * rng.standard_normal(shape))).astype(F)
    return u


def tgv3d_u0(N, L):
    """3-D Taylor-Green vortex IC at the face locations loc(i,I) (src/core.jl:177)."""
    k = F(2 * np.pi / L)
    shape = tuple(reversed(N))
    u = np.zeros((3,) + shape, F)
    idx = [np.arange(1, n + 1, dtype=F) for n in N]

    def coord(d, i):
        return idx[d] - F(1.5) - (F(0.5) if d == i else F(0))

    for i in range(2):
        x = coord(0, i)[None, None, :]
        y = coord(1, i)[None, :, None]
        z = coord(2, i)[:, None, None]
        if i == 0:
            u[0] = (-np.sin(k * x) * np.cos(k * y) * np.cos(k * z)).astype(F)
        else:
            u[1] = (np.cos(k * x) * np.sin(k * y) * np.cos(k * z)).astype(F)
    return u


def make_pair(dims, uBC, nu=0.0, dt0=0.25, perdir=(), exitBC=False, lam="quick", u0=None, sphere=None, torus=None, pois="ml",
              smoother="gs", flags=0, measure=True, **kw):
    """Returns (oracle OracleSim, B200 Simulation) for the same configuration.  `sphere`=(center, radius),
    `torus`=(center, R, r) (axis along x).  `flags`: WL_FLAG_* bits or a list of names from wl_b200.lib.FLAGS.
    `measure=False` builds the B200 side without its body (the caller uploads μ₀/μ₁/V itself)."""
    import oracle
    import wl_b200 as wl

    if not isinstance(flags, int):
        flags = sum(wl.lib.FLAGS[f] for f in flags)
    o = oracle.OracleSim(dims, uBC, nu=nu, dt0=dt0, perdir=perdir, exitBC=exitBC, lam=lam, u0=u0, pois=pois)
    if sphere is not None:
        o.measure_sphere(*sphere)
    if torus is not None:
        o.measure_torus(*torus)
    o.init_pois()
    if smoother != "gs":
        o.set_solver(smoother=smoother)
    body = None
    if measure:
        body = wl.Sphere(*sphere) if sphere is not None else wl.Torus(*torus) if torus is not None else None
    u0f = None
    if u0 is not None:
        u0arr = np.asarray(u0, F)

        def u0f(i, x):
            return u0arr[i]
    s = wl.Simulation(dims, uBC, 1.0, ν=nu, Δt=dt0, perdir=perdir, exitBC=exitBC, λ=lam, u0=u0f, body=body,
                      pois="multilevel" if pois == "ml" else "single", smoother=smoother, flags=flags, **kw)
    return o, s
