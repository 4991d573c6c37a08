"""Runs the sphere-wake configuration for a few steps and reports Δt, iterations and velocity extrema (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import wl_b200 as wl
m = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 13
R = m / 8; c = m / 2 - 1
sim = wl.Simulation((2 * m, m, m), (1., 0., 0.), 2 * R, ν=2 * R / 3700, exitBC=True, body=wl.Sphere((c, c, c), R))
for k in range(steps):
    try:
        wl.lib.check(sim.flow.L, sim.flow.L.wl_mom_step(sim.flow.h))
        sim.flow.sync()
    except Exception as e:
        print("step", k, "failed:", e)
        break
    u = sim.flow.u
    a = np.abs(u[np.isfinite(u)])
    nz = a[a > 0]
    print("step", k, "dt", sim.flow.Δt[-1], "iters", list(sim.pois.n[-2:]), "max|u|", a.max(), "min nonzero |u|", nz.min(), "nonfinite", int((~np.isfinite(u)).sum()))
