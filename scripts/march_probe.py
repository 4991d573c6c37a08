import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import wl_b200 as wl
from bench import make_case, build_sim, tgv_u0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
case = make_case("tgv%d" % n)
sim = build_sim(case, tgv_u0(n))
fl = sim.flow
wl.lib.check(fl.L, fl.L.wl_sim_step_n(fl.h, 2))
fl.set_profiling(True)
wl.lib.check(fl.L, fl.L.wl_sim_step_n(fl.h, 3))
t = fl.timings()
cells = (n + 2) ** 3
# per-launch time of the finest-level launch ≈ total/launches for single-level kernels
for k in ("fm_conv", "f_div_residual", "f_correct", "f_cfl", "f_resid_fix"):
    c, ms = t[k]; print("%-16s %7.1f us/launch" % (k, ms / c * 1e3))
fl.set_profiling(False)
for kind in ("gs", "jacobi"):
    fl.set_profiling(True)
    for _ in range(5): sim.pois.smooth(0, kind, 0.9)
    tt = fl.timings(); fl.set_profiling(False)
    print(kind, {k: round(v[1] / v[0] * 1e3, 1) for k, v in tt.items() if k != "k_set_scalar"})
