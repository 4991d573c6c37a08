"""Forces, moments and running averages (src/Metrics.jl:111-257) and checkpoints (ext/WaterLilyJLD2Ext.jl:11-50) over the
B200 library: every reduction and every average runs on the device; only the results (12 doubles, or arrays on request) cross."""
import ctypes as C

import numpy as np

from . import lib as _lib

F = np.float32


def _forces(sim, x0=None):
    fl = sim.flow
    if not getattr(sim, "device_body", False):
        raise _lib.WLError("forces and moments need a parametrised body (Sphere, Torus, set operations): nds(body,x,t) of an AutoBody "
                           "closure lives on the host")
    out = (C.c_double * 12)()
    x0a = None
    if x0 is not None:
        x0a = (C.c_float * 3)(*[float(v) for v in x0] + [0.0] * (3 - len(x0)))
    _lib.check(fl.L, fl.L.wl_body_forces(fl.h, x0a, out))
    return np.array(out[:], np.float64).reshape(4, 3)[:, :fl.D]


def pressure_force(sim):
    """pressure_force(sim) (src/Metrics.jl:121-129)"""
    return _forces(sim)[0]


def viscous_force(sim):
    """viscous_force(sim) (src/Metrics.jl:147-154)"""
    return _forces(sim)[1]


def total_force(sim):
    """total_force(sim) = pressure_force(sim) .+ viscous_force(sim) (src/Metrics.jl:161)"""
    f = _forces(sim)
    return f[0] + f[1]


def pressure_moment(x0, sim):
    """pressure_moment(x₀,sim) (src/Metrics.jl:169-177)"""
    return _forces(sim, x0)[2]


def viscous_moment(x0, sim):
    """viscous_moment(x₀,sim) (src/Metrics.jl:184-190)"""
    return _forces(sim, x0)[3]


def total_moment(x0, sim):
    """total_moment(x₀,sim) (src/Metrics.jl:197)"""
    f = _forces(sim, x0)
    return f[2] + f[3]


class MeanFlow:
    """MeanFlow(flow; t_init=time(flow), uu_stats=false) (src/Metrics.jl:205-229): P, U, UU live on the device."""

    def __init__(self, flow, uu_stats=False):
        self.flow = flow
        self.uu_stats = bool(uu_stats)
        _lib.check(flow.L, flow.L.wl_meanflow_init(flow.h, int(uu_stats)))

    def _get(self, which, nc):
        fl = self.flow
        sp = tuple(reversed(fl.N))
        out = np.empty(sp if nc == 1 else (nc,) + sp, F)
        _lib.check(fl.L, fl.L.wl_meanflow_download(fl.h, which, out.ctypes.data_as(C.c_void_p), 0))
        return out

    def _set(self, which, arr):
        fl = self.flow
        a = np.ascontiguousarray(arr, F)
        _lib.check(fl.L, fl.L.wl_meanflow_upload(fl.h, which, a.ctypes.data_as(C.c_void_p), 0))

    P = property(lambda s: s._get(0, 1))
    U = property(lambda s: s._get(1, s.flow.D))

    @property
    def UU(self):
        """UU[j, i, ...] = ⟨u_i u_j⟩ (component i + D·j slowest, like flow.μ₁)"""
        if not self.uu_stats:
            return None
        D = self.flow.D
        return self._get(2, D * D).reshape((D, D) + tuple(reversed(self.flow.N)))

    @property
    def t(self):
        fl = self.flow
        n = C.c_int(0)
        _lib.check(fl.L, fl.L.wl_meanflow_get_times(fl.h, None, C.byref(n)))
        out = np.zeros(n.value, F)
        _lib.check(fl.L, fl.L.wl_meanflow_get_times(fl.h, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out

    def time(self):
        """time(meanflow) = t[end] - t[1] (src/Metrics.jl:227)"""
        t = self.t
        return float(F(t[-1] - t[0]))

    def update(self, flow=None):
        """update!(meanflow, flow) (src/Metrics.jl:236-247)"""
        _lib.check(self.flow.L, self.flow.L.wl_meanflow_update(self.flow.h))

    def reset(self, t_init=0.0):
        """reset!(meanflow; t_init) (src/Metrics.jl:229-234)"""
        _lib.check(self.flow.L, self.flow.L.wl_meanflow_reset(self.flow.h, float(t_init)))

    def uu(self):
        """uu(meanflow): τ[I,i,j] = UU[I,i,j] − U[I,i]·U[I,j] (src/Metrics.jl:249-256)"""
        U, UU = self.U, self.UU
        D = self.flow.D
        tau = np.empty_like(UU)
        for i in range(D):
            for j in range(D):
                tau[j, i] = UU[j, i] - U[i] * U[j]
        return tau

    def copy_to(self, flow=None):
        """copy!(flow, meanflow) (src/Metrics.jl:258-261)"""
        _lib.check(self.flow.L, self.flow.L.wl_meanflow_copy_to_flow(self.flow.h))


def update(meanflow, flow=None):
    meanflow.update(flow)


def save(fname, obj):
    """save!(fname, flow) / save!(fname, meanflow) (ext/WaterLilyJLD2Ext.jl:11-31): u, p, Δt (flow) or P, U, UU, t (meanflow) in the
    reference layout (Float32, ghost-padded) — a NumPy .npz here, since JLD2 is a Julia package."""
    from .flow import Flow
    from .simulation import Simulation
    if isinstance(obj, Simulation):
        obj = obj.flow
    if isinstance(obj, Flow):
        np.savez(fname, u=obj.u, p=obj.p, dt=np.asarray(obj.Δt, F))
    elif isinstance(obj, MeanFlow):
        d = dict(P=obj.P, U=obj.U, t=obj.t)
        if obj.uu_stats:
            d["UU"] = obj.UU
        np.savez(fname, **d)
    else:
        raise TypeError("save(fname, flow | sim | meanflow)")


def load(obj, fname):
    """load!(flow; fname) / load!(meanflow; fname) (ext/WaterLilyJLD2Ext.jl:33-50): restores u, p, Δt (the next step continues the
    saved run bit for bit) or P, U, UU, t."""
    from .flow import Flow
    from .simulation import Simulation
    if not str(fname).endswith(".npz"):
        fname = str(fname) + ".npz"
    z = np.load(fname)
    if isinstance(obj, Simulation):
        obj = obj.flow
    if isinstance(obj, Flow):
        obj.upload("u", z["u"])
        obj.upload("p", z["p"])
        dt = np.ascontiguousarray(z["dt"], F)
        _lib.check(obj.L, obj.L.wl_set_dt(obj.h, dt.ctypes.data_as(C.POINTER(C.c_float)), len(dt)))
    elif isinstance(obj, MeanFlow):
        obj._set(0, z["P"])
        obj._set(1, z["U"])
        if obj.uu_stats and "UU" in z:
            obj._set(2, z["UU"].reshape((-1,) + z["P"].shape))
        t = np.ascontiguousarray(z["t"], F)
        _lib.check(obj.flow.L, obj.flow.L.wl_meanflow_set_times(obj.flow.h, t.ctypes.data_as(C.POINTER(C.c_float)), len(t)))
    else:
        raise TypeError("load(flow | sim | meanflow, fname)")
