"""Immersed-body set-up on the host (numpy, Float32): `measure!` for static bodies.

Body measurement is set-up only on this path (BASELINE.json north_star): the host evaluates the signed
distance function, fills μ₀, μ₁, V, σ exactly as `measure!(flow,body;ϵ)` does (src/Body.jl:28-51,
src/AutoBody.jl:29-37) and uploads them through wl_upload; wl_update then rebuilds the Poisson hierarchy.
"""
import numpy as np

F = np.float32


class NoBody:
    """struct NoBody (src/Body.jl:81-83)"""


class AutoBody:
    """AutoBody(sdf) (src/AutoBody.jl:10-14) for a static map.  `sdf(x)` takes a tuple of D coordinate arrays and
    returns the signed distance; `grad(x)` returns the D gradient components (the reference gets them from
    ForwardDiff; here the caller supplies them, or central differences in Float64 are used)."""

    def __init__(self, sdf, grad=None):
        self.sdf = sdf
        self._grad = grad

    def grad(self, x):
        if self._grad is not None:
            return self._grad(x)
        h = 1e-4
        x64 = [np.asarray(c, np.float64) for c in x]
        out = []
        for d in range(len(x)):
            xp = list(x64)
            xm = list(x64)
            xp[d] = x64[d] + h
            xm[d] = x64[d] - h
            out.append(((np.asarray(self.sdf(xp), np.float64) - np.asarray(self.sdf(xm), np.float64)) / (2 * h)).astype(F))
        return out


class _Parametrised:
    """Bodies the library measures ON THE DEVICE (wl_set_body / wl_measure): primitives with a rigid translation
    map (x,t) -> x .- velocity.*t, combined with the reference's lazy set operations (src/Body.jl:88-103):
    a | b or a + b (∪), a & b (∩), a - b (a ∩ −b)."""

    def prims(self):
        raise NotImplementedError

    def __or__(self, other):
        return BodySetOp(self, other, 0)

    __add__ = __or__

    def __and__(self, other):
        return BodySetOp(self, other, 1)

    def __sub__(self, other):
        return BodySetOp(self, other, 2)


class BodySetOp(_Parametrised):
    """SetBody(op, a, b) restricted to left-leaning trees: b must be a primitive."""

    def __init__(self, a, b, op):
        pb = b.prims()
        if len(pb) != 1:
            raise ValueError("set operations take a primitive on the right-hand side: write ((a ∪ b) ∪ c), not a ∪ (b ∪ c)")
        self._p = a.prims() + [dict(pb[0], op=op)]

    def prims(self):
        return [dict(p) for p in self._p]


class Sphere(AutoBody, _Parametrised):
    """AutoBody((x,t)->√sum(abs2, x .- center) - radius, (x,t)->x .- velocity.*t)  (README.md:118-119; circle in 2-D)"""

    def __init__(self, center, radius, velocity=None):
        self.c = [F(v) for v in center]
        self.R = F(radius)
        self.vel = [F(v) for v in (velocity if velocity is not None else (0,) * len(center))]
        super().__init__(self._sdf, self._g)

    def prims(self):
        return [dict(kind=0, op=0, center=[float(v) for v in self.c], R=float(self.R), r=0.0, vel=[float(v) for v in self.vel])]

    def _sdf(self, x):
        s = F(0)
        for d, xd in enumerate(x):
            s = s + (xd - self.c[d]) * (xd - self.c[d])
        return np.sqrt(s) - self.R

    def _g(self, x):
        s = F(0)
        for d, xd in enumerate(x):
            s = s + (xd - self.c[d]) * (xd - self.c[d])
        m = np.sqrt(s)
        with np.errstate(invalid="ignore", divide="ignore"):
            return [(xd - self.c[d]) / m for d, xd in enumerate(x)]


class Torus(AutoBody, _Parametrised):
    """Torus with its axis along x (WaterLily-Examples ThreeD_Donut recipe): major radius R, minor radius r."""

    def __init__(self, center, R, r, velocity=None):
        self.c = [F(v) for v in center]
        self.R = F(R)
        self.r = F(r)
        self.vel = [F(v) for v in (velocity if velocity is not None else (0, 0, 0))]
        super().__init__(self._sdf, self._g)

    def prims(self):
        return [dict(kind=1, op=0, center=[float(v) for v in self.c], R=float(self.R), r=float(self.r), vel=[float(v) for v in self.vel])]

    def _parts(self, x):
        xx, y, z = x[0] - self.c[0], x[1] - self.c[1], x[2] - self.c[2]
        rho = np.sqrt(y * y + z * z)
        q = rho - self.R
        m = np.sqrt(q * q + xx * xx)
        return xx, y, z, rho, q, m

    def _sdf(self, x):
        return self._parts(x)[5] - self.r

    def _g(self, x):
        xx, y, z, rho, q, m = self._parts(x)
        with np.errstate(invalid="ignore", divide="ignore"):
            return [xx / m, (q / m) * (y / rho), (q / m) * (z / rho)]


def prim_array(body):
    """ctypes array of wl_body_prim for a parametrised body"""
    from .lib import BodyPrim
    ps = body.prims()
    arr = (BodyPrim * len(ps))()
    for q, p in enumerate(ps):
        arr[q].kind, arr[q].op, arr[q].R, arr[q].r = p["kind"], p["op"], p["R"], p["r"]
        for d in range(3):
            arr[q].center[d] = p["center"][d] if d < len(p["center"]) else 0.0
            arr[q].vel[d] = p["vel"][d] if d < len(p["vel"]) else 0.0
    return arr


def _sinpi64(x):
    """sinpi in double with the argument reduced to [-1/2, 1/2] first (exact at multiples of 1/2 like Julia's sinpi); |x| ≤ 1.5"""
    a = np.abs(x)
    r = np.where(a <= 0.5, np.sin(np.pi * a), np.sin(np.pi * (1.0 - a)))
    return np.where(x < 0, -r, r)


def _sinpi(x):
    return _sinpi64(np.asarray(x, F).astype(np.float64)).astype(F)


def _cospi(x):
    return _sinpi64(0.5 - np.abs(np.asarray(x, F).astype(np.float64))).astype(F)


def _kern0(d):  # src/Body.jl:55
    return (F(1) + d + _sinpi(d) / F(np.pi)) / F(2)


def _kern1(d):  # src/Body.jl:56
    return (F(1) - d * d) / F(4) - (d * _sinpi(d) + (F(1) + _cospi(d)) / F(np.pi)) / (F(2) * F(np.pi))


def mu0_kernel(d, eps):
    """μ₀(d,ϵ) = d/ϵ < -1+√eps(d) ? 0 : kern₀(min(d/ϵ,1))  (src/Body.jl:59); eps(d) is the ulp at d."""
    d = np.asarray(d, F)
    e = F(eps)
    s = d / e
    ulp = np.spacing(np.abs(d)).astype(F)
    return np.where(s < F(-1) + np.sqrt(ulp), F(0), _kern0(np.minimum(s, F(1)))).astype(F)


def mu1_kernel(d, eps):
    """μ₁(d,ϵ) = ϵ·kern₁(clamp(d/ϵ,-1,1))  (src/Body.jl:60)"""
    d = np.asarray(d, F)
    e = F(eps)
    return (e * _kern1(np.clip(d / e, F(-1), F(1)))).astype(F)


def _loc(N, i, zoff=0):
    """Coordinate arrays of loc(i,I) for every cell (src/core.jl:177); i=-1 gives cell centres.  C-order shapes.
    `zoff` = global index of the local plane 0 (z-slab decomposition)."""
    D = len(N)
    shape = tuple(reversed(N))
    out = []
    for d in range(D):
        idx = np.arange(1, N[d] + 1, dtype=F) + F(zoff if d == 2 else 0) - F(1.5) - (F(0.5) if d == i else F(0))
        sh = [1] * D
        sh[D - 1 - d] = N[d]
        out.append(np.broadcast_to(idx.reshape(sh), shape))
    return out


def measure_body(N, body, eps=1.0, zoff=0):
    """measure!(flow,body;ϵ) for a static AutoBody (src/Body.jl:28-51 + src/AutoBody.jl:29-37) before the two BC!
    calls (those run on the device).  N = ghost-padded sizes.  Returns (mu0[D,...], mu1[D*D,...], V[D,...], sigma[...])
    in C order (component slowest, x fastest); μ₁[I,i,j] is component i + D*j."""
    D = len(N)
    shape = tuple(reversed(N))
    eps = F(eps)
    V = np.zeros((D,) + shape, F)
    mu0 = np.ones((D,) + shape, F)
    mu1 = np.zeros((D * D,) + shape, F)
    sigma = np.zeros(shape, F)
    d2 = F((F(2) + eps) ** 2)
    inner = tuple(slice(1, -1) for _ in range(D))
    xc = [c[inner] for c in _loc(N, -1, zoff)]
    sig_in = np.asarray(body.sdf(xc), F)
    sigma[inner] = sig_in
    band = sig_in * sig_in < d2
    inside_far = (~band) & (sig_in < 0)
    for i in range(D):
        xf = [c[inner][band] for c in _loc(N, i, zoff)]
        d = np.asarray(body.sdf(xf), F)
        near = ~(d * d > d2)  # measure(): skip n when d² > fastd²
        g = [np.asarray(c, F) for c in body.grad(xf)]
        bad = np.zeros(d.shape, bool)
        for c in g:
            bad |= np.isnan(c)
        use = near & ~bad
        m = np.sqrt(sum((c * c for c in g), F(0))).astype(F)
        with np.errstate(invalid="ignore", divide="ignore"):
            dn = np.where(use, d / m, d).astype(F)
            n = [np.where(use, c / m, F(0)).astype(F) for c in g]
        di = np.where(np.abs(dn) <= F(0.5), dn, np.copysign(dn, sig_in[band])).astype(F)
        tmp = mu0[i][inner]
        tmp[band] = mu0_kernel(di, eps)
        tmp[inside_far] = F(0)
        mu0[i][inner] = tmp
        k1 = mu1_kernel(di, eps)
        for j in range(D):
            tmp = mu1[i + D * j][inner]
            tmp[band] = k1 * n[j]
            mu1[i + D * j][inner] = tmp
    return mu0, mu1, V, sigma
