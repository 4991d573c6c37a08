"""`Flow`, `MultiLevelPoisson`/`Poisson` and `mom_step!` of the reference (src/Flow.jl:114-167,
src/MultiLevelPoisson.jl:61-77, src/Poisson.jl:22-39) as thin views over one wl_handle.

All state lives on the B200 in the library's internal layout; the attributes below download (or upload)
arrays in the reference layout, transposed to C order: reference `u[x,y,z,c]` is `flow.u[c,z,y,x]` here.
"""
import ctypes as C

import numpy as np

from . import lib as _lib

F = np.float32

# convective schemes λ (src/Flow.jl:4-6) are selected by name; the kernels specialise on them at compile time
quick, cds, vanLeer = "quick", "cds", "vanLeer"
_LAM = {"quick": 0, "cds": 1, "vanLeer": 2}
_FIELDS = {"u": 0, "u0": 1, "f": 2, "p": 3, "sigma": 4, "V": 5, "mu0": 6, "mu1": 7}
_LVL = {"L": 0, "D": 1, "iD": 2, "x": 3, "eps": 4, "r": 5, "z": 6}


class Forcing:
    """g(i,x,t) = g0[i] + g1[i]·t: the enumerated, spatially uniform stand-in for the reference's body-force closure `g`
    (accelerate!, src/Flow.jl:64-73).  Evaluated on the device inside the flux kernels."""

    def __init__(self, g0=None, g1=None):
        self.g0, self.g1 = g0, g1


class TimeBC:
    """uBC(i,x,t) = U0[i] + U1[i]·t + ½·U2[i]·t²: the enumerated stand-in for a function-valued boundary velocity (BC! with a
    function uBC, src/core.jl:201-219; its time derivative enters accelerate!, src/Flow.jl:72)."""

    def __init__(self, U0, U1=None, U2=None):
        self.U0, self.U1, self.U2 = tuple(float(v) for v in U0), U1, U2

    def __iter__(self):
        return iter(self.U0)

    def __len__(self):
        return len(self.U0)


def sgs(*a, **k):
    """sgs!(flow,u,t;νₜ,S,Cs,Δ) (src/util.jl:46-76) — a marker: `sim_step(sim, udf=sgs, νₜ=smagorinsky, Cs=…, Δ=…)` selects the
    library's built-in device implementation (wl_set_sgs); it cannot be called on the host."""
    raise _lib.WLError("sgs is evaluated on the device: pass it as sim_step(sim, udf=sgs, νₜ=smagorinsky, Cs=…, Δ=…)")


def smagorinsky(*a, **k):
    """smagorinsky(I;S,Cs,Δ) = (Cs·Δ)²·√(S[I,:,:]⋅S[I,:,:]) (src/util.jl:57-62) — a marker for the `νₜ` keyword of `sim_step`."""
    raise _lib.WLError("smagorinsky is evaluated on the device: pass it as the νₜ keyword of sim_step together with udf=sgs")


def loc_grid(N, i, zoff=0):
    """loc(i,I) for all cells of a ghost-padded grid (src/core.jl:177): list of D broadcastable coordinate arrays."""
    from .body import _loc
    return _loc(N, i, zoff)


def dist_unique_id(fmad=False):
    """128-byte NCCL id for wl_create_dist; rank 0 calls this and broadcasts the bytes to the other ranks."""
    L = _lib.load_library(fmad)
    buf = C.create_string_buffer(128)
    _lib.check(L, L.wl_dist_unique_id(buf))
    return buf.raw


class Flow:
    """Flow(N,uBC;Δt,ν,u0,perdir,exitBC,λ,T=Float32) (src/Flow.jl:133-147).  `uBC` must be a tuple (function-valued
    boundary conditions are host closures and are not supported through the C ABI); `u0` may be a tuple or a
    function u0(i, x) evaluated on the host at the face locations (apply!, src/Flow.jl:81-82; i is 0-based here)."""

    def __init__(self, N, uBC, Δt=0.25, ν=0.0, g=None, u0=None, perdir=(), exitBC=False, λ=quick, T=np.float32,
                 pois="multilevel", smoother="gs", tol=1e-4, itmx=0, device=0, fmad=False, flags=0, dist=None):
        if callable(uBC):
            raise _lib.WLError("function-valued uBC is a host closure: not supported by the B200 C ABI (SURVEY.md §8b); "
                               "use TimeBC(U0, U1, U2) for U(t) = U0 + U1·t + ½·U2·t²")
        if g is not None and not isinstance(g, Forcing):
            raise _lib.WLError("acceleration g(i,x,t) is a host closure: not supported by the B200 C ABI; "
                               "use Forcing(g0, g1) for g(t) = g0 + g1·t")
        tbc = uBC if isinstance(uBC, TimeBC) else None
        if tbc is not None:
            uBC = tbc.U0
        if np.dtype(T) != np.float32:
            raise _lib.WLError("only T=Float32 is supported")
        self.L = _lib.load_library(fmad)
        D = len(N)
        self.D = D
        self.dims = tuple(int(n) for n in N)  # GLOBAL interior size
        # z-slab decomposition: dist = (rank, world, nccl_id_bytes); arrays seen through this object are the rank's own slab
        self.rank, self.world = (dist[0], dist[1]) if dist else (0, 1)
        self.N = tuple(n + 2 for n in self.dims)
        self.zoff = 0
        if self.world > 1:
            nzl = self.dims[2] // self.world
            self.N = self.N[:2] + (nzl + 2,)
            self.zoff = self.rank * nzl
        self.uBC = tuple(float(v) for v in uBC)
        self.ν = float(ν)
        self.exitBC = bool(exitBC)
        self.perdir = tuple(perdir)
        self.λ = λ
        cfg = _lib.Config()
        cfg.D = D
        for d in range(3):
            cfg.n[d] = self.dims[d] if d < D else 1
            cfg.uBC[d] = self.uBC[d] if d < D else 0.0
            cfg.perdir[d] = 1 if (d + 1) in self.perdir else 0
        cfg.exitBC = int(exitBC)
        cfg.lam = _LAM[λ]
        cfg.nu = ν
        cfg.dt0 = Δt
        cfg.pois_kind = 0 if pois == "multilevel" else 1
        cfg.smoother = 0 if smoother == "gs" else 1
        cfg.tol = tol
        cfg.itmx = itmx
        cfg.device = device
        cfg.flags = flags
        self.h = C.c_void_p()
        if self.world > 1:
            idb = C.create_string_buffer(bytes(dist[2]), 128)
            _lib.check(self.L, self.L.wl_create_dist(C.byref(cfg), self.rank, self.world, idb, C.byref(self.h)))
        else:
            _lib.check(self.L, self.L.wl_create(C.byref(cfg), C.byref(self.h)))
        if g is not None or tbc is not None:
            def v3(a):
                out = (C.c_float * 3)()
                for d, x in enumerate(a or ()):
                    out[d] = float(x)
                return out
            _lib.check(self.L, self.L.wl_set_forcing(self.h, v3(g.g0 if g else None), v3(g.g1 if g else None),
                                                     v3(tbc.U1 if tbc else None), v3(tbc.U2 if tbc else None)))
        if u0 is not None:
            # apply!(u0, a.u) (src/Flow.jl:140): component by component, so that an array the caller already holds goes to the
            # device without being assembled into one more host copy
            shp = tuple(reversed(self.N))
            for i in range(D):
                a = u0(i, loc_grid(self.N, i, self.zoff)) if callable(u0) else u0[i]
                a = np.asarray(a, F)
                if a.shape != shp:
                    a = np.broadcast_to(a, shp)
                self.upload_component("u", i, a)
            _lib.check(self.L, self.L.wl_apply_bc(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.wl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- array access (reference layout, C order) -------------------------------------------
    def _shape(self, name):
        nc = {"u": self.D, "u0": self.D, "f": self.D, "V": self.D, "mu0": self.D, "mu1": self.D * self.D}.get(name, 1)
        sp = tuple(reversed(self.N))
        return sp if nc == 1 else (nc,) + sp

    def download(self, name):
        out = np.empty(self._shape(name), F)
        _lib.check(self.L, self.L.wl_download(self.h, _FIELDS[name], out.ctypes.data_as(C.c_void_p), 0))
        return out

    def upload_component(self, name, comp, arr):
        a = np.ascontiguousarray(arr, F)
        sp = tuple(reversed(self.N))
        if a.shape != sp:
            raise ValueError(f"{name}[{comp}]: expected shape {sp}, got {a.shape}")
        _lib.check(self.L, self.L.wl_upload_component(self.h, _FIELDS[name], int(comp), a.ctypes.data_as(C.c_void_p), 0))

    def upload(self, name, arr):
        a = np.ascontiguousarray(arr, F)
        if a.shape != self._shape(name):
            raise ValueError(f"{name}: expected shape {self._shape(name)}, got {a.shape}")
        _lib.check(self.L, self.L.wl_upload(self.h, _FIELDS[name], a.ctypes.data_as(C.c_void_p), 0))

    def upload_device(self, name, ptr):
        _lib.check(self.L, self.L.wl_upload(self.h, _FIELDS[name], C.c_void_p(ptr), 1))

    def download_device(self, name, ptr):
        _lib.check(self.L, self.L.wl_download(self.h, _FIELDS[name], C.c_void_p(ptr), 1))

    u = property(lambda s: s.download("u"))
    u0 = property(lambda s: s.download("u0"))
    f = property(lambda s: s.download("f"))
    p = property(lambda s: s.download("p"))
    σ = property(lambda s: s.download("sigma"))
    V = property(lambda s: s.download("V"))
    μ0 = property(lambda s: s.download("mu0"))
    μ1 = property(lambda s: s.download("mu1"))

    @property
    def Δt(self):
        n = C.c_int(0)
        _lib.check(self.L, self.L.wl_get_dt(self.h, None, C.byref(n)))
        out = np.zeros(n.value, F)
        _lib.check(self.L, self.L.wl_get_dt(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out

    def time(self):
        """time(a) = sum(Δt[1:end-1])  (src/Flow.jl:174)"""
        t = C.c_double()
        _lib.check(self.L, self.L.wl_time(self.h, C.byref(t)))
        return t.value

    def CFL(self):
        v = C.c_float()
        _lib.check(self.L, self.L.wl_cfl(self.h, C.byref(v)))
        return v.value

    def sync(self):
        _lib.check(self.L, self.L.wl_sync(self.h))

    def set_profiling(self, on=True):
        _lib.check(self.L, self.L.wl_set_profiling(self.h, int(on)))

    def timings(self):
        """{kernel: (launches, total_ms, total_cells)} recorded with CUDA events on the library's stream since the last call;
        total_cells = ghost-padded cells of the levels those launches ran on, summed over the launches."""
        n = C.c_int(0)
        _lib.check(self.L, self.L.wl_get_timings(self.h, None, C.byref(n)))
        buf = C.create_string_buffer(n.value + 16)
        n = C.c_int(n.value + 16)
        _lib.check(self.L, self.L.wl_get_timings(self.h, buf, C.byref(n)))
        out = {}
        for line in buf.value.decode().splitlines():
            k, c, ms, cells = line.split()
            c0, m0, n0 = out.get(k, (0, 0.0, 0.0))
            out[k] = (c0 + int(c), m0 + float(ms), n0 + float(cells))
        return out

    @property
    def launches(self):
        n = C.c_int64()
        _lib.check(self.L, self.L.wl_launch_count(self.h, C.byref(n)))
        return n.value


class MultiLevelPoisson:
    """MultiLevelPoisson(flow.p,flow.μ₀,flow.σ;perdir) (src/MultiLevelPoisson.jl:61-77): a view on the hierarchy the
    flow's handle owns (x≡p, L≡μ₀, z≡σ alias the flow arrays exactly as in the reference)."""

    kind = "multilevel"

    def __init__(self, flow):
        self.flow = flow
        self.L = flow.L
        self.h = flow.h

    @property
    def n(self):
        n = C.c_int(0)
        _lib.check(self.L, self.L.wl_get_iters(self.h, None, C.byref(n)))
        out = np.zeros(n.value, np.int16)
        _lib.check(self.L, self.L.wl_get_iters(self.h, out.ctypes.data_as(C.POINTER(C.c_int16)), C.byref(n)))
        return out

    @property
    def nlevels(self):
        n = C.c_int()
        _lib.check(self.L, self.L.wl_num_levels(self.h, C.byref(n)))
        return n.value

    def level_dims(self, level):
        N = (C.c_int32 * 3)()
        _lib.check(self.L, self.L.wl_level_dims(self.h, level, N))
        return tuple(N[d] for d in range(self.flow.D))

    def level(self, level, name):
        Ns = self.level_dims(level)
        nc = self.flow.D if name == "L" else 1
        sp = tuple(reversed(Ns))
        out = np.empty(sp if nc == 1 else (nc,) + sp, F)
        _lib.check(self.L, self.L.wl_download_level(self.h, level, _LVL[name], out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def set_level(self, level, name, arr):
        a = np.ascontiguousarray(arr, F)
        _lib.check(self.L, self.L.wl_upload_level(self.h, level, _LVL[name], a.ctypes.data_as(C.POINTER(C.c_float))))

    def update(self):
        """update!(pois) (src/MultiLevelPoisson.jl:79-86)"""
        _lib.check(self.L, self.L.wl_update(self.h))

    def mult(self):
        """mult!(pois,x) with x≡flow.p: flow.σ = A·p (src/Poisson.jl:63-69)"""
        _lib.check(self.L, self.L.wl_pois_mult(self.h))

    def solver(self):
        """solver!(pois) (src/MultiLevelPoisson.jl:108-127 / src/Poisson.jl:204-214): solves A·p = σ in place"""
        n = C.c_int()
        _lib.check(self.L, self.L.wl_pois_solve(self.h, C.byref(n)))
        return n.value

    def residual(self):
        r2 = C.c_float()
        _lib.check(self.L, self.L.wl_pois_residual(self.h, C.byref(r2)))
        return r2.value

    def smooth(self, level=0, kind="gs", ω=1.0):
        _lib.check(self.L, self.L.wl_pois_smooth(self.h, level, {"gs": 0, "jacobi": 1, "pcg": 2}[kind], ω))

    def vcycle(self, ω=1.0):
        _lib.check(self.L, self.L.wl_pois_vcycle(self.h, ω))

    @property
    def log(self):
        n = C.c_int(0)
        _lib.check(self.L, self.L.wl_get_solver_log(self.h, None, C.byref(n)))
        out = np.zeros((n.value, 4), F)
        if n.value:
            _lib.check(self.L, self.L.wl_get_solver_log(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out


class Poisson(MultiLevelPoisson):
    """Poisson(x,L,z;perdir) (src/Poisson.jl:22-39): the single-level system solved with pcg!."""
    kind = "single"


def mom_step(flow, pois=None):
    """mom_step!(a,b) (src/Flow.jl:156-167)"""
    _lib.check(flow.L, flow.L.wl_mom_step(flow.h))
