"""Host-side logic of the z-slab multi-GPU path, on CPU with the gloo backend (world_size 2).

The device side (halo exchange, all-reduces, replicated coarse levels) is covered on real GPUs by
scripts/dist_check.py, which compares a P-rank run bit for bit with the single-GPU run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    from waterlily_loader import wl  # noqa: F401  (imports the package without touching CUDA)
    n = 16
    nzl = n // world
    # 1) every rank builds only its own slab of the initial condition; gathered they must equal the global field
    mine = torch.from_numpy(bench.tgv_u0(n, rank * nzl, nzl)[:, 1:-1].copy())
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    glob = bench.tgv_u0(n)
    ok_ic = np.array_equal(torch.cat(parts, dim=1).numpy(), glob[:, 1:-1])
    # ghost planes of a slab are the neighbouring slab's interior planes (what the halo exchange must deliver)
    sl = bench.tgv_u0(n, rank * nzl, nzl)
    ok_ghost = np.array_equal(sl[:, 0], glob[:, rank * nzl]) and np.array_equal(sl[:, -1], glob[:, rank * nzl + nzl + 1])
    # 2) the 128-byte communicator id travels from rank 0 to everyone (here a stand-in blob; the real one comes from NCCL)
    blob = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(blob, 0)
    ok_id = bool((blob == torch.arange(128, dtype=torch.uint8)).all())
    # 3) body measurement of a slab (global coordinates through zoff) equals the slice of the global measurement
    N = (18, 18, 18)
    body = wl.Sphere((8.0, 7.0, 9.0), 3.0)
    g0, g1, gV, gs = wl.measure_body(N, body, 1.0)
    Nl = (18, 18, nzl + 2)
    l0, l1, lV, ls = wl.measure_body(Nl, body, 1.0, zoff=rank * nzl)
    z = slice(rank * nzl + 1, rank * nzl + nzl + 1)
    ok_body = np.array_equal(l0[:, 1:-1], g0[:, z]) and np.array_equal(l1[:, 1:-1], g1[:, z]) and np.array_equal(ls[1:-1], gs[z])
    # 4) max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, ok_ic, ok_ghost, ok_id, ok_body, float(t.item())))
    dist.destroy_process_group()


def test_slab_decomposition_host_logic_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok_ic, ok_ghost, ok_id, ok_body, tmax in res:
        assert ok_ic and ok_ghost and ok_id and ok_body, (rank, ok_ic, ok_ghost, ok_id, ok_body)
        assert tmax == float(world)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads on a CPU-only box and exports every function include/wl_b200.h declares."""
    import re
    import wl_b200 as wl
    L = wl.load_library()
    hdr = open(os.path.join(ROOT, "include", "wl_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(wl_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert L.wl_device_count() >= 0


def test_no_cuda_device_fails_loudly():
    import wl_b200 as wl
    L = wl.load_library()
    if L.wl_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(wl.WLError, match="no CPU fallback"):
        wl.Simulation((16, 16), (1.0, 0.0), 16)


def test_host_closures_are_rejected():
    import wl_b200 as wl
    with pytest.raises(wl.WLError):
        wl.Flow((16, 16), lambda i, x, t: 1.0)
    with pytest.raises(wl.WLError):
        wl.Flow((16, 16), (1.0, 0.0), g=lambda i, x, t: 0.0)


def test_python_body_kernels_match_reference_values():
    """μ₀/μ₁ known answers of test/test_bodies.jl:2-5 for the product's own (NumPy) measure, which never touches the oracle."""
    import wl_b200 as wl
    assert wl.mu0_kernel(3.0, 6.0) == wl.mu0_kernel(0.5, 1.0)
    assert wl.mu0_kernel(0.0, 1.0) == 0.5
    assert wl.mu0_kernel(np.float32(np.finfo(np.float32).eps) - np.float32(1), 1.0) == 0.0
    assert abs(float(wl.mu1_kernel(0.0, 2.0)) - 2 * (1 / 4 - 1 / np.pi**2)) < 1e-7


def test_flag_table_matches_the_header():
    """waterlily.jl_b200/lib.py's FLAGS mirror the WL_FLAG_* bits of include/wl_b200.h (the parity tests select kernel variants by name)."""
    import re
    import wl_b200 as wl
    hdr = open(os.path.join(ROOT, "include", "wl_b200.h")).read()
    bits = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"WL_FLAG_([A-Z0-9_]+)\s*=\s*(\d+)", hdr)}
    assert bits and bits == wl.lib.FLAGS
    vals = sorted(bits.values())
    assert all(v & (v - 1) == 0 for v in vals) and len(set(vals)) == len(vals)  # distinct single bits
