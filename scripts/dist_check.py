"""torchrun -n P scripts/dist_check.py: P-rank z-slab run vs the same case on one GPU (rank 0 runs both)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
import wl_b200 as wl
from util import tgv3d_u0, rel_l2

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
case = sys.argv[3] if len(sys.argv) > 3 else "tgv"
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.frombuffer(bytearray(wl.dist_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(idt, 0)
idb = bytes(idt.cpu().numpy().tobytes())
if case == "tgv":
    dims = (n, n, n); u0g = tgv3d_u0((n + 2,) * 3, n); nu = float(np.float32(1 / (2 * np.pi / n * 1600)))
    kw = dict(ν=nu, perdir=(1, 2, 3)); uBC = (0., 0., 0.); body = None
else:
    dims = (2 * n, n, n); u0g = None; kw = dict(ν=n / 8 / 100., exitBC=True); uBC = (1., 0., 0.); body = wl.Sphere((n / 2 - 1,) * 3, n / 8)
nzl = dims[2] // world
u0f = None
if u0g is not None:
    sl = u0g[:, rank * nzl: rank * nzl + nzl + 2]
    u0f = lambda i, x: sl[i]
sim = wl.Simulation(dims, uBC, float(n), u0=u0f, body=body, device=local, dist=(rank, world, idb), **kw)
wl.lib.check(sim.flow.L, sim.flow.L.wl_sim_step_n(sim.flow.h, steps))
u = sim.flow.u; p = sim.flow.p
dt = sim.flow.Δt; its = sim.pois.n
if rank == 0:
    u0f1 = (lambda i, x: u0g[i]) if u0g is not None else None
    ref = wl.Simulation(dims, uBC, float(n), u0=u0f1, body=body, device=local, **kw)
    wl.lib.check(ref.flow.L, ref.flow.L.wl_sim_step_n(ref.flow.h, steps))
    ur = ref.flow.u[:, 0: nzl + 2]; pr = ref.flow.p[0: nzl + 2]
    print("rank0 slab vs single-GPU: u rel-L2 %.3e  p rel-L2 %.3e  max|du| %.3e" % (rel_l2(u, ur), rel_l2(p[1:-1], pr[1:-1]), np.abs(u - ur).max()))
    print("dt equal:", np.array_equal(dt, ref.flow.Δt), "iters", list(its), list(ref.pois.n), "uni", sim.flow.launches)
dist.barrier()
dist.destroy_process_group()
