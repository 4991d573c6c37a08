"""The flux kernels evaluate quick(u,c,d) = median((5c+2d-u)/6, c, median(10c-9u, c, d)) (src/Flow.jl:6, :27-36) in a clamp form,
mirrored by an exact multiplication with s = sign(d-c) (wl_conv4.cuh: flux_p).  This is a statement about IEEE Float32 values, so it can
be checked on the CPU: both forms, operation by operation in Float32, must give the same value for every input."""
import numpy as np

F = np.float32


def median3(a, b, c):
    return np.maximum(np.minimum(a, b), np.minimum(np.maximum(a, b), c))


def quick_reference(u, c, d):
    a = (F(5) * c + F(2) * d - u) / F(6)
    b = F(10) * c - F(9) * u
    return median3(a, c, median3(b, c, d))


def quick_clamp(u, c, d, uf):
    """flux_p's form: s = sign(t*uf) with t = u0c - um1; here (u,c,d) are already the upwind-selected values, so d - c = ±t with the
    sign of uf: s = sign(d - c)."""
    t = np.where(uf > 0, d - c, c - d)  # u0c - um1
    s = np.where(np.signbit(t * uf), F(-1), F(1)).astype(F)
    cs, ds, us = c * s, d * s, u * s
    a = (F(5) * cs + F(2) * ds - us) / F(6)
    b = F(10) * cs - F(9) * us
    return s * np.minimum(np.maximum(np.minimum(a, b), cs), ds)


def _cases(n, rng):
    u = rng.standard_normal(n).astype(F)
    c = rng.standard_normal(n).astype(F)
    d = rng.standard_normal(n).astype(F)
    # ties and degenerate triples
    k = n // 10
    c[:k] = d[:k]
    u[k:2 * k] = c[k:2 * k]
    u[2 * k:3 * k] = d[2 * k:3 * k]
    c[3 * k:4 * k] = 0
    d[4 * k:5 * k] = 0
    # wide dynamic range
    scale = (10.0 ** rng.integers(-30, 30, n)).astype(F)
    return u * scale, c * scale, d * scale


def test_quick_clamp_form_equals_median_form():
    rng = np.random.default_rng(11)
    u, c, d = _cases(2_000_000, rng)
    uf = rng.standard_normal(u.size).astype(F)  # any non-zero face velocity: its sign decides which of c, d is upwind
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        ref = quick_reference(u, c, d)
        got = quick_clamp(u, c, d, uf)
    ok = np.isfinite(ref)
    assert ok.mean() > 0.99
    assert np.array_equal(ref[ok], got[ok])  # value-identical (±0 compare equal)


def test_median3_is_the_middle_value():
    rng = np.random.default_rng(12)
    x = rng.standard_normal((3, 100000)).astype(F)
    x[1, :1000] = x[0, :1000]
    assert np.array_equal(median3(x[0], x[1], x[2]), np.sort(x, axis=0)[1])


def test_division_by_six_through_double_is_the_float_division():
    """div6_slow (wl_kernels.cuh): (float)((double)x / 6.0) == x / 6.f for every Float32, subnormal results and ties included.
    (The GPU test checks all 2^32 inputs on the device; here a 10 M sample of bit patterns plus every subnormal with 3 | m or not.)"""
    rng = np.random.default_rng(13)
    bits = rng.integers(0, 2 ** 32, 10_000_000, dtype=np.uint64).astype(np.uint32)
    sub = np.arange(0, 1 << 23, 7, dtype=np.uint32)            # subnormals and their negatives
    small = (np.arange(0, 1 << 22, 5, dtype=np.uint32) + np.uint32(1 << 23))  # the smallest normals: quotients are subnormal
    x = np.concatenate([bits, sub, sub | np.uint32(1 << 31), small]).view(F)
    with np.errstate(all="ignore"):
        a = x / F(6)
        b = (x.astype(np.float64) / 6.0).astype(F)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert bool(same.all())
