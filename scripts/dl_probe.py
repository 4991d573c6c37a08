"""Where does the end-to-end time of the sphere workload go?  (download of u, p through the pinned ring)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import wl_b200 as wl
case = bench.make_case("sphere")
for rep in range(2):
    t0 = time.perf_counter()
    sim = bench.build_sim(case, None, None, 0)
    sim.flow.sync(); t1 = time.perf_counter()
    for _ in range(3):
        wl.sim_step(sim)
    sim.flow.sync(); t2 = time.perf_counter()
    for _ in range(5):
        wl.sim_step(sim); dt = float(sim.flow.Δt[-1])
    t3 = time.perf_counter()
    u = sim.flow.u; t4 = time.perf_counter()
    p = sim.flow.p; t5 = time.perf_counter()
    u2 = sim.flow.u; t6 = time.perf_counter()
    sim.close(); t7 = time.perf_counter()
    print("build %.3f warm3 %.3f step5 %.3f u %.3f p %.3f u again %.3f close %.3f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))
