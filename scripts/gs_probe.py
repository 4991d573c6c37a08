import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import wl_b200 as wl
from bench import make_case, build_sim, tgv_u0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
case = make_case("tgv%d" % n)
sim = build_sim(case, tgv_u0(n))
fl = sim.flow
wl.lib.check(fl.L, fl.L.wl_sim_step_n(fl.h, 2))
for lvl in range(min(4, sim.pois.nlevels)):
    for kind in ("gs", "jacobi"):
        fl.set_profiling(True)
        for _ in range(5):
            sim.pois.smooth(lvl, kind, 0.9)
        t = fl.timings()
        fl.set_profiling(False)
        print("level", lvl, sim.pois.level_dims(lvl), kind, {k: round(v[1] / v[0] * 1e3, 1) for k, v in t.items() if k != "k_set_scalar"}, "us/launch")
