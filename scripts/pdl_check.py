"""Programmatic dependent launch must not change a bit: the bench workloads stepped with and without it (WL_FLAG_NO_PDL)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import wl_b200 as wl

for name, steps in (("sphere", 8), ("tgv128", 40), ("tgv512", 6)):
    case = bench.make_case(name)
    u0 = bench.tgv_u0(case["dims"][0]) if case["u0"] else None
    outs = []
    for flags in (0, wl.lib.FLAGS["no_pdl"]):
        body = None
        if case["body"]:
            body = wl.Torus(*case["body"][1:]) if case["body"][0] == "torus" else wl.Sphere(*case["body"])
        u0f = (lambda i, x: u0[i]) if u0 is not None else None
        s = wl.Simulation(case["dims"], case["uBC"], case["L"], ν=case["nu"], perdir=case["perdir"], exitBC=case["exitBC"], body=body, u0=u0f, flags=flags)
        wl.lib.check(s.flow.L, s.flow.L.wl_sim_step_n(s.flow.h, steps))
        outs.append((s.flow.u, s.flow.p, list(s.pois.n), np.asarray(s.flow.Δt).copy()))
        s.close()
    a, b = outs
    print(name, "u equal", bool(np.array_equal(a[0], b[0])), "p equal", bool(np.array_equal(a[1], b[1])), "iters equal", a[2] == b[2], "dt equal", bool(np.array_equal(a[3], b[3])),
          "iters", a[2][:12])
