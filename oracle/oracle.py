"""ctypes driver for oracle/libwl_oracle.so (the CPU restatement of WaterLily.jl's hot path).

TEST INFRASTRUCTURE ONLY — never imported by the product package.  Each method cites the
reference function it drives (paths relative to the WaterLily.jl tree).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwl_oracle.so")
_lib = None

FIELD_IDS = {"u": 0, "u0": 1, "f": 2, "p": 3, "sigma": 4, "V": 5, "mu0": 6, "mu1": 7}
LEVEL_IDS = {"L": 0, "D": 1, "iD": 2, "x": 3, "eps": 4, "r": 5, "z": 6}
LAMBDA = {"quick": 0, "cds": 1, "vanLeer": 2}


class Config(C.Structure):
    _fields_ = [("D", C.c_int), ("n", C.c_int * 3), ("uBC", C.c_float * 3), ("perdir", C.c_int * 3),
                ("exitBC", C.c_int), ("lam", C.c_int), ("nu", C.c_float), ("dt0", C.c_float)]


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "wl_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libwl_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.wlo_create.restype = C.c_void_p
        L.wlo_create.argtypes = [C.POINTER(Config)]
        L.wlo_field.restype = fp
        L.wlo_field.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.wlo_level_field.restype = fp
        L.wlo_level_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        for name in ("wlo_destroy", "wlo_init_bc", "wlo_update", "wlo_mom_step", "wlo_bdim", "wlo_bc_u", "wlo_pois_residual"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.wlo_init_pois.argtypes = [C.c_void_p, C.c_int]
        L.wlo_set_solver.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.wlo_project.argtypes = [C.c_void_p, C.c_float]
        L.wlo_conv_diff.argtypes = [C.c_void_p, C.c_int]
        L.wlo_cfl.argtypes = [C.c_void_p]
        L.wlo_cfl.restype = C.c_float
        L.wlo_measure_sphere.argtypes = [C.c_void_p, fp, C.c_float, C.c_float]
        L.wlo_measure_torus.argtypes = [C.c_void_p, fp, C.c_float, C.c_float, C.c_float]
        L.wlo_measure_prims.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.wlo_measure_prims.restype = None
        L.wlo_body_forces.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, fp, C.POINTER(C.c_double)]
        L.wlo_body_forces.restype = None
        L.wlo_set_forcing.argtypes = [C.c_void_p, fp, fp, fp, fp]
        L.wlo_set_forcing.restype = None
        L.wlo_set_sgs.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.wlo_set_sgs.restype = None
        L.wlo_sgs.argtypes = [C.c_void_p, C.c_int]
        L.wlo_sgs.restype = None
        for name in ("wlo_dt_len", "wlo_iters_len", "wlo_log_len", "wlo_num_levels", "wlo_pois_solve"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.wlo_get_dt.argtypes = [C.c_void_p, fp]
        L.wlo_push_dt.argtypes = [C.c_void_p, C.c_float]
        L.wlo_get_iters.argtypes = [C.c_void_p, C.POINTER(C.c_int16)]
        L.wlo_get_log.argtypes = [C.c_void_p, fp]
        L.wlo_pois_mult.argtypes = [C.c_void_p, fp]
        L.wlo_pois_L2.argtypes = [C.c_void_p]
        L.wlo_pois_L2.restype = C.c_float
        L.wlo_pois_Linf.argtypes = [C.c_void_p]
        L.wlo_pois_Linf.restype = C.c_float
        L.wlo_pois_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.wlo_pois_vcycle.argtypes = [C.c_void_p, C.c_float]
        for name in ("wlo_quick", "wlo_vanleer", "wlo_cds"):
            getattr(L, name).argtypes = [C.c_float] * 3
            getattr(L, name).restype = C.c_float
        for name in ("wlo_mu0", "wlo_mu1"):
            getattr(L, name).argtypes = [C.c_float] * 2
            getattr(L, name).restype = C.c_float
        for name in ("wlo_phiuL", "wlo_phiuR", "wlo_phiu"):
            getattr(L, name).argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_int]
            getattr(L, name).restype = C.c_float
        L.wlo_phiuP.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.wlo_phiuP.restype = C.c_float
        L.wlo_phi.argtypes = [fp, C.c_int, C.c_int]
        L.wlo_phi.restype = C.c_float
        ip = C.POINTER(C.c_int)
        L.wlo_BC.argtypes = [C.c_int, ip, fp, fp, C.c_int, ip]
        L.wlo_exitBC.argtypes = [C.c_int, ip, fp, fp, C.c_float]
        L.wlo_perBC.argtypes = [C.c_int, ip, fp, ip]
        L.wlo_num_threads.restype = C.c_int
        L.wlo_set_num_threads.argtypes = [C.c_int]
        L.wlo_set_num_threads.restype = None
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(seq):
    return (C.c_int * len(seq))(*seq)


def _per(D, perdir):
    """perdir is a tuple of 1-based periodic directions like the reference's `perdir=(1,2)`."""
    return [1 if (d + 1) in perdir else 0 for d in range(3)]


class OracleSim:
    """Oracle-side `Simulation` (src/WaterLily.jl:86-106) for a constant uBC tuple.

    Arrays are numpy views straight into the oracle's memory, in the reference layout
    transposed to C order: shape (D, N3, N2, N1) for 3-D vectors (component slowest, x fastest).
    """

    def __init__(self, dims, uBC, nu=0.0, dt0=0.25, perdir=(), exitBC=False, lam="quick", u0=None, pois="ml"):
        self.L = lib()
        D = len(dims)
        self.D = D
        self.dims = tuple(dims)
        self.N = tuple(d + 2 for d in dims)
        cfg = Config()
        cfg.D = D
        for d in range(3):
            cfg.n[d] = dims[d] if d < D else 1
            cfg.uBC[d] = uBC[d] if d < D else 0.0
        per = _per(D, perdir)
        for d in range(3):
            cfg.perdir[d] = per[d]
        cfg.exitBC = int(exitBC)
        cfg.lam = LAMBDA[lam]
        cfg.nu = nu
        cfg.dt0 = dt0
        self.h = C.c_void_p(self.L.wlo_create(C.byref(cfg)))
        if u0 is not None:  # apply!(u0,u) on all cells incl. ghosts, then BC!  (src/Flow.jl:140-142)
            self.field("u")[...] = u0
            self.L.wlo_init_bc(self.h)
        self.pois_kind = pois
        self.nlevels = None

    def __del__(self):
        try:
            self.L.wlo_destroy(self.h)
        except Exception:
            pass

    # --- array access -----------------------------------------------------------------
    def _shape(self, ncomp):
        sp = tuple(reversed(self.N))
        return sp if ncomp == 1 else (ncomp,) + sp

    def field(self, name):
        n = C.c_uint64()
        p = self.L.wlo_field(self.h, FIELD_IDS[name], C.byref(n))
        a = np.ctypeslib.as_array(p, shape=(n.value,))
        cells = int(np.prod(self.N))
        return a.reshape(self._shape(n.value // cells))

    def level_field(self, level, name):
        N = (C.c_int * 3)()
        p = self.L.wlo_level_field(self.h, level, LEVEL_IDS[name], N)
        Ns = tuple(N[d] for d in range(self.D))
        cells = int(np.prod(Ns))
        ncomp = self.D if name == "L" else 1
        a = np.ctypeslib.as_array(p, shape=(cells * ncomp,))
        sp = tuple(reversed(Ns))
        return a.reshape(sp if ncomp == 1 else (ncomp,) + sp)

    # --- setup ------------------------------------------------------------------------
    def measure_sphere(self, center, radius, eps=1.0):
        c = np.zeros(3, np.float32)
        c[: self.D] = center
        self.L.wlo_measure_sphere(self.h, _fp(c), radius, eps)

    def measure_torus(self, center, R, r, eps=1.0):
        c = np.asarray(center, np.float32)
        self.L.wlo_measure_torus(self.h, _fp(c), R, r, eps)

    def measure_prims(self, prims, eps=1.0, t=0.0):
        """measure!(flow, body; t, ϵ) for primitives with a translation map and set operations (src/Body.jl:28-51,88-107).
        `prims`: list of dicts kind/op/center/R/r/vel (kind 0 sphere, 1 torus; op 0 ∪, 1 ∩, 2 −)."""
        class P(C.Structure):
            _fields_ = [("kind", C.c_int), ("op", C.c_int), ("c", C.c_float * 3), ("R", C.c_float), ("r", C.c_float), ("vel", C.c_float * 3)]
        arr = (P * len(prims))()
        for q, p in enumerate(prims):
            arr[q].kind, arr[q].op, arr[q].R, arr[q].r = p["kind"], p["op"], p["R"], p["r"]
            for d in range(3):
                arr[q].c[d] = p["center"][d] if d < len(p["center"]) else 0.0
                arr[q].vel[d] = p["vel"][d] if d < len(p["vel"]) else 0.0
        self.L.wlo_measure_prims(self.h, arr, len(prims), eps, t)

    def _prim_array(self, prims):
        class P(C.Structure):
            _fields_ = [("kind", C.c_int), ("op", C.c_int), ("c", C.c_float * 3), ("R", C.c_float), ("r", C.c_float), ("vel", C.c_float * 3)]
        arr = (P * len(prims))()
        for q, p in enumerate(prims):
            arr[q].kind, arr[q].op, arr[q].R, arr[q].r = p["kind"], p["op"], p["R"], p["r"]
            for d in range(3):
                arr[q].c[d] = p["center"][d] if d < len(p["center"]) else 0.0
                arr[q].vel[d] = p["vel"][d] if d < len(p["vel"]) else 0.0
        return arr

    def body_forces(self, prims, x0=None, t=None):
        """rows: pressure_force, viscous_force, pressure_moment(x₀), viscous_moment(x₀) (src/Metrics.jl:121-190) at t = time(flow)"""
        arr = self._prim_array(prims)
        x0a = np.zeros(3, np.float32)
        if x0 is not None:
            x0a[: len(x0)] = x0
        out = (C.c_double * 12)()
        self.L.wlo_body_forces(self.h, arr, len(prims), self.time() if t is None else t, _fp(x0a), out)
        return np.array(out[:], np.float64).reshape(4, 3)[:, : self.D]

    def set_forcing(self, g0=(0, 0, 0), g1=(0, 0, 0), U1=(0, 0, 0), U2=(0, 0, 0)):
        """g_i(t) = g0_i + g1_i·t and U_i(t) = uBC_i + U1_i·t + ½·U2_i·t² in place of the closures g(i,x,t), uBC(i,x,t)"""
        def v(a):
            out = np.zeros(3, np.float32)
            out[: len(a)] = a
            return out
        self._forcing = [v(g0), v(g1), v(U1), v(U2)]
        self.L.wlo_set_forcing(self.h, *[_fp(a) for a in self._forcing])

    def set_sgs(self, Cs, Delta):
        """udf = sgs! with νₜ = smagorinsky (src/util.jl:46-76); Cs·Δ = 0 switches it off"""
        self.L.wlo_set_sgs(self.h, float(Cs), float(Delta))

    def sgs(self, from_u0=False):
        """sgs!(flow,u,t) alone: flow.f += sub-grid fluxes of flow.u (or flow.u⁰)"""
        self.L.wlo_sgs(self.h, int(from_u0))

    def time_next(self):
        """sum(Δt): the default t of measure!(sim) (src/WaterLily.jl:146)"""
        return float(np.float32(np.sum(self.dt.astype(np.float64))))

    def init_pois(self):
        """pois_ctor(flow) (src/WaterLily.jl:97,105)"""
        self.nlevels = self.L.wlo_init_pois(self.h, 0 if self.pois_kind == "ml" else 1)
        if self.nlevels < 0:
            raise AssertionError("MultiLevelPoisson requires size=a2ⁿ, where n>2")
        return self.nlevels

    def set_solver(self, tol=1e-4, itmx=32, smoother="gs"):
        self.L.wlo_set_solver(self.h, tol, itmx, 0 if smoother == "gs" else 1)

    def update(self):
        self.L.wlo_update(self.h)

    # --- stepping ---------------------------------------------------------------------
    def mom_step(self):
        if self.nlevels is None:
            self.init_pois()
        self.L.wlo_mom_step(self.h)

    @property
    def dt(self):
        n = self.L.wlo_dt_len(self.h)
        out = np.zeros(n, np.float32)
        self.L.wlo_get_dt(self.h, _fp(out))
        return out

    @property
    def iters(self):
        n = self.L.wlo_iters_len(self.h)
        out = np.zeros(n, np.int16)
        self.L.wlo_get_iters(self.h, out.ctypes.data_as(C.POINTER(C.c_int16)))
        return out

    @property
    def log(self):
        n = self.L.wlo_log_len(self.h)
        out = np.zeros((n, 3), np.float32)
        if n:
            self.L.wlo_get_log(self.h, _fp(out))
        return out

    def time(self):
        """time(a) = sum(Δt[1:end-1]) (src/Flow.jl:174).  Julia's pairwise/@simd Float32 sum has no pinned order; the oracle
        returns the Float32 nearest to the exact sum (accumulated in double), which every summation order agrees with to a few ulp."""
        return float(np.float32(np.sum(self.dt[:-1].astype(np.float64))))

    def sim_step_until(self, t_end, U=1.0, Lscale=1.0, max_steps=10**9):
        """sim_step!(sim,t_end;remeasure=false) (src/WaterLily.jl:128-135)"""
        k = 0
        while self.time() * U / Lscale < t_end and k < max_steps:
            self.mom_step()
            k += 1
        return k
