"""`Simulation`, `sim_step!`, `sim_time`, `measure!` (src/WaterLily.jl:86-149) over the B200 library."""
import ctypes as C

import numpy as np

from . import lib as _lib
from .body import NoBody, _Parametrised, measure_body, prim_array
from .flow import Flow, MultiLevelPoisson, Poisson, quick, sgs, smagorinsky

F = np.float32


class Simulation:
    """Simulation(dims,uBC,L;U,Δt,ν,ϵ,perdir,u0,exitBC,λ,body,T,flow_ctor,pois_ctor) (src/WaterLily.jl:93-106).

    `mem` is implied (the B200); `flow_ctor(dims,uBC,**kw)` and `pois_ctor(flow)` are the reference's plug-in
    factories and default to the B200 `Flow` and `MultiLevelPoisson`."""

    def __init__(self, dims, uBC, L, U=None, Δt=0.25, ν=0.0, g=None, ϵ=1, perdir=(), u0=None, exitBC=False, λ=quick,
                 body=None, T=np.float32, flow_ctor=None, pois_ctor=None, host_measure=False, **kw):
        from .flow import TimeBC
        if (callable(uBC) or isinstance(uBC, TimeBC)) and U is None:
            raise AssertionError("`U` (velocity scale) must be specified if boundary conditions `uBC` is a `Function`")
        if U is None:
            U = float(np.sqrt(sum(float(v) ** 2 for v in uBC)))
        self.U, self.L, self.ϵ = U, L, ϵ
        self.body = body if body is not None else NoBody()
        pois = kw.pop("pois", "multilevel")
        if flow_ctor is None:
            def flow_ctor(dims, uBC, **k):
                return Flow(dims, uBC, **k)
        self.flow = flow_ctor(dims, uBC, u0=u0, Δt=Δt, ν=ν, g=g, λ=λ, T=T, perdir=perdir, exitBC=exitBC, pois=pois, **kw)
        if pois_ctor is None:
            def pois_ctor(flow):
                return MultiLevelPoisson(flow) if pois == "multilevel" else Poisson(flow)
        self.pois = pois_ctor(self.flow)
        # parametrised bodies (Sphere, Torus, set operations of them) are measured on the device; `host_measure=True` keeps the
        # NumPy path (what a host binding does for an arbitrary AutoBody closure)
        self.device_body = isinstance(self.body, _Parametrised) and not host_measure
        if self.device_body:
            arr = prim_array(self.body)
            _lib.check(self.flow.L, self.flow.L.wl_set_body(self.flow.h, arr, len(arr), float(ϵ)))
        measure(self, t=0.0)

    def close(self):
        self.flow.close()


def measure(sim, t=None):
    """measure!(sim, t=sum(Δt)) (src/WaterLily.jl:146-149): measure!(flow,body;t,ϵ) + update!(pois).  Parametrised bodies are
    measured by the library on the device at time t; any other AutoBody is measured here in NumPy (static: t is ignored)."""
    if isinstance(sim.body, NoBody):
        return
    fl = sim.flow
    if getattr(sim, "device_body", False):
        if t is None:
            tt = C.c_double()
            _lib.check(fl.L, fl.L.wl_time_next(fl.h, C.byref(tt)))
            t = tt.value
        _lib.check(fl.L, fl.L.wl_measure(fl.h, float(t)))
        sim.pois.update()
        return
    mu0, mu1, V, sigma = measure_body(fl.N, sim.body, sim.ϵ, fl.zoff)
    fl.upload("mu0", mu0)
    fl.upload("mu1", mu1)
    fl.upload("V", V)
    fl.upload("sigma", sigma)
    _lib.check(fl.L, fl.L.wl_measure_bc(fl.h))
    sim.pois.update()


def sim_time(sim):
    """sim_time(sim) = time(sim)*U/L (src/WaterLily.jl:117)"""
    return sim.flow.time() * sim.U / sim.L


def sim_step(sim, t_end=None, remeasure=False, max_steps=2**62, verbose=False, udf=None, νₜ=None, S=None, Cs=None, Δ=None, **kwargs):
    """sim_step!(sim,t_end;remeasure,max_steps,verbose,udf,kwargs...) and sim_step!(sim;remeasure,udf,kwargs...) (src/WaterLily.jl:128-139).
    `remeasure=True` re-measures a parametrised (device-measured) body at t = sum(Δt) before every step, inside the library;
    a body given as a host closure cannot be re-measured per step through the C ABI.
    `udf`: the one user-defined function the library has built in is the reference's own LES model,
    `sim_step(sim; udf=sgs, νₜ=smagorinsky, Cs, Δ)` (src/util.jl:46-76; `S` is scratch the library owns, accepted and ignored);
    any other udf is a host closure and is rejected."""
    fl = sim.flow
    if udf is sgs:
        if νₜ is not smagorinsky or Cs is None or Δ is None:
            raise _lib.WLError("udf=sgs needs νₜ=smagorinsky (the built-in eddy viscosity), Cs and Δ")
        _lib.check(fl.L, fl.L.wl_set_sgs(fl.h, float(Cs), float(Δ)))
    elif udf is not None or kwargs:
        raise _lib.WLError("udf is a host closure: not supported by the B200 C ABI (built in: udf=sgs with νₜ=smagorinsky)")
    else:
        _lib.check(fl.L, fl.L.wl_set_sgs(fl.h, 0.0, 0.0))
    if remeasure and not isinstance(sim.body, NoBody):
        if not getattr(sim, "device_body", False):
            raise _lib.WLError("remeasure=true needs a parametrised body (Sphere, Torus, set operations): an AutoBody closure "
                               "lives on the host; pass remeasure=False or measure and upload it yourself")
    _lib.check(fl.L, fl.L.wl_set_remeasure(fl.h, int(bool(remeasure) and getattr(sim, "device_body", False))))
    if t_end is None:
        _lib.check(fl.L, fl.L.wl_mom_step(fl.h))
        return 1
    if verbose:
        k = 0
        while sim_time(sim) < t_end and k < max_steps:
            _lib.check(fl.L, fl.L.wl_mom_step(fl.h))
            sim_info(sim)
            k += 1
        return k
    k = C.c_int64()
    _lib.check(fl.L, fl.L.wl_sim_step_until(fl.h, float(t_end), float(sim.U), float(sim.L), int(max_steps), C.byref(k)))
    return k.value


def sim_info(sim):
    """sim_info(sim) (src/WaterLily.jl:155)"""
    print(f"tU/L={round(sim_time(sim), 4)}, Δt={round(float(sim.flow.Δt[-1]), 3)}")
