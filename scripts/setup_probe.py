"""Where does Simulation(...) set-up time go at 512³?  (e2e includes it)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import wl_b200 as wl
case = bench.make_case("tgv512")
u0 = bench.tgv_u0(512)
pinned = torch.from_numpy(u0).pin_memory().numpy()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    fl = wl.Flow(case["dims"], case["uBC"], ν=case["nu"], perdir=case["perdir"])
    fl.sync(); t1 = time.perf_counter()
    fl.upload("u", pinned); fl.sync(); t2 = time.perf_counter()
    arr = np.empty(pinned.shape, np.float32); arr[...] = pinned; t3 = time.perf_counter()
    fl.upload("u", arr); fl.sync(); t4 = time.perf_counter()
    u = fl.u; t5 = time.perf_counter()
    fl.close(); t6 = time.perf_counter()
    print("create %.3f  upload(pinned) %.3f  host copy %.3f  upload(pageable) %.3f  download u %.3f  close %.3f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5))
