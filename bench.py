#!/usr/bin/env python
"""bench.py — sim_step! throughput of the B200 mom_step! library, ns per cell per step (lower is better).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tgv512|tgv128|sphere|...] [--impl reference]

One "step" is one mom_step! (predictor + projection + corrector + projection + CFL) of the workload.  `value` is the
whole-job number with all state resident in HBM; `e2e` goes through the public host API (wl_b200.Simulation /
sim_step) with the initial condition coming from host memory and u, p going back to host memory inside the timed
region.  See DESIGN.md §measurement for the algorithmic-byte table behind `roofline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "sim_step_ns_per_cell_per_step"
UNIT = "ns/cell/step"

WORKLOADS = {
    # BASELINE.json configs; tgv512 is the configuration the north_star target (≤0.05 ns/cell/step) is quoted on
    "tgv512": dict(kind="tgv", n=512),
    "tgv256": dict(kind="tgv", n=256),
    "tgv128": dict(kind="tgv", n=128),   # configs[1]
    "tgv64": dict(kind="tgv", n=64),
    "sphere": dict(kind="sphere", dims=(512, 256, 256)),  # configs[2]
    "sphere128": dict(kind="sphere", dims=(128, 64, 64)),
    "donut": dict(kind="donut", dims=(1024, 512, 512)),   # configs[3]: torus, axis along x, 8 GPUs
    "donut256": dict(kind="donut", dims=(256, 128, 128)),
    "tgv1024": dict(kind="tgv", n=1024),                  # configs[4]: strong-scaling sweep
}

# Algorithmic 4-byte words per launch and ghost-padded cell of the level the launch runs on, every field read once / written once
# (perfect halo reuse) — DESIGN.md §kernels.  (general variable-coefficient form, uniform-coefficient form)
ALG_WORDS = {
    "fm_conv": (10.5, 7.5),        # general: read u/u⁰ 3(+3 corrector) V 3, write f 3; uniform: read u⁰ 3 (+u 3 corrector), write u 3
    "fm_conv4g": (10.5, 10.5),     # general mode: read u/u⁰ 3(+3 corrector) V 3, write f 3
    "fm_conv4": (7.5, 7.5),        # uniform mode only: read u⁰ 3 (+u 3 corrector), write u 3
    "f_vsmooth": (4.125, 4.125),   # uniform mode only: read r x (+ coarse x 1/8), write r' x  — replaces f_increment<PROLONG>, f_gs_a, 3 f_gs_half, f_increment
    "k_conv_bdim1": (10.5, 10.5),
    "k_bdim2": (22.5, 22.5),       # read f3 V3 μ₀3 μ₁9 (+u3 corrector); write u3
    "k_f_lowghost": (0, 0),
    "f_div_residual": (12, 6),     # read u3 p [L3 D iD]; write x r [z]
    "k_div_residual": (12, 12),
    "f_divres_uni": (6, 6),        # uniform mode: read u3 p; write x r
    "f_jacobi_uni2": (4.125, 4.125),  # uniform mode, level 1: read r x; write r' x (+ coarse r 1/8)
    "f_resid_fix": (1, 1),
    "k_resid_fix": (1, 1),
    "f_jacobi": (9.125, 4.125),    # read r x [iD L3 D]; write r' x (+ coarse r 1/8)
    "k_jacobi": (9, 9),
    "k_restrict": (1.125, 1.125),
    "f_gs_a": (6, 2),              # read r [iD L3]; write ϵ
    "f_gs_half": (7, 3),           # read ϵ r [iD L3]; write ϵ
    "k_gs_init": (3, 3),
    "k_gs_sweep": (5, 5),
    "f_increment": (9, 5),         # read ϵ r x [L3 D]; write r x   (prolongation source: ϵ costs 1/8)
    "k_increment": (9, 9),
    "k_prolong_inc": (8.125, 8.125),
    "f_correct": (11, 8),          # read x u3 [L3]; write u3 p
    "f_correct_cfl": (8, 8),       # uniform mode: the same out of place; the corrector launch also forms CFL's flux_out (no extra bytes)
    "k_correct": (11, 11),
    "f_cfl": (4, 3),               # read u3 [write σ]
    "k_cfl": (4, 4),
}


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures: profiles/ncu_traffic.json
    maps workload → kernel → {bytes, file, src, src_sha16}.  The entry is stamped with the hash of the kernel's source file at capture
    time; if the source has changed since, the number is reported with "traffic_stale": true instead of silently going stale."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f)[workload][kernel]
        with open(os.path.join(ROOT, ent["src"]), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()[:16]
        return ent["bytes"], {"traffic_source": ent["file"], "traffic_stale": sha != ent["src_sha16"]}
    except Exception:
        return None, {}


# general mode, semi-uniform blocks: the face coefficients come from registers, D and iD are still read; BDIM-2 reads f only
SEMI_WORDS = {"f_jacobi": 6.125, "f_gs_a": 3, "f_gs_half": 4, "f_increment": 6, "f_div_residual": 9, "f_correct": 8, "k_bdim2": 7.5}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tgv_u0(n, z0=0, nzl=None):
    """TGV initial condition on the ghost-padded planes z0 … z0+nzl+1 (the whole domain by default)."""
    F = np.float32
    nzl = n if nzl is None else nzl
    k = F(2 * np.pi / n)
    u = np.zeros((3, nzl + 2, n + 2, n + 2), F)
    idx = np.arange(1, n + 3, dtype=F)
    idz = np.arange(1 + z0, z0 + nzl + 3, dtype=F)
    for i in range(2):
        x = (idx - F(1.5) - (F(0.5) if i == 0 else F(0)))[None, None, :]
        y = (idx - F(1.5) - (F(0.5) if i == 1 else F(0)))[None, :, None]
        z = (idz - F(1.5))[:, None, None]
        u[i] = (-np.sin(k * x) * np.cos(k * y) * np.cos(k * z) if i == 0 else np.cos(k * x) * np.sin(k * y) * np.cos(k * z)).astype(F)
    return u


def make_case(name):
    w = WORKLOADS[name]
    if w["kind"] == "tgv":
        n = w["n"]
        nu = float(np.float32(1 / (2 * np.pi / n * 1600)))
        return dict(dims=(n, n, n), uBC=(0.0, 0.0, 0.0), L=float(n), nu=nu, perdir=(1, 2, 3), exitBC=False, body=None, u0=("tgv", n))
    d = w["dims"]
    m = d[1]
    if w["kind"] == "donut":  # centre (m/2,m/2,m/2), major radius m/4, minor m/16, L=R, ν=R/1000
        R = m / 4
        c = m / 2
        return dict(dims=d, uBC=(1.0, 0.0, 0.0), L=R, nu=R / 1000, perdir=(), exitBC=True, body=("torus", (c, c, c), R, m / 16), u0=None)
    R = m / 8
    c = m / 2 - 1
    return dict(dims=d, uBC=(1.0, 0.0, 0.0), L=2 * R, nu=2 * R / 3700, perdir=(), exitBC=True, body=((c, c, c), R), u0=None)


def build_sim(case, u0_host=None, dist=None, device=0):
    import wl_b200 as wl
    body = None
    if case["body"]:
        body = wl.Torus(*case["body"][1:]) if case["body"][0] == "torus" else wl.Sphere(*case["body"])
    u0 = None
    if u0_host is not None:
        def u0(i, x):
            return u0_host[i]
    return wl.Simulation(case["dims"], case["uBC"], case["L"], ν=case["nu"], perdir=case["perdir"], exitBC=case["exitBC"], body=body, u0=u0,
                         dist=dist, device=device)


class ClockSampler:
    def __init__(self, gpu=0):
        self.samples, self.reasons = [], set()
        self.maxmhz = None
        self._stop = threading.Event()
        self.gpu = gpu
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.maxmhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=5)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.maxmhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def oracle_run(case, steps, warmup):
    """The CPU port (oracle/) on all host threads: returns seconds per step and the thread count."""
    import oracle
    from oracle import OracleSim
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every core this process may run on explicitly
    oracle.lib().wlo_set_num_threads(len(os.sched_getaffinity(0)))
    u0 = tgv_u0(case["u0"][1]) if case["u0"] else None
    o = OracleSim(case["dims"], case["uBC"], nu=case["nu"], perdir=case["perdir"], exitBC=case["exitBC"], u0=u0)
    if case["body"]:
        o.measure_sphere(*case["body"])
    o.init_pois()
    for _ in range(warmup):
        o.mom_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.mom_step()
    dt = (time.perf_counter() - t0) / steps
    return dt, oracle.lib().wlo_num_threads(), int(np.sum(o.iters[-2 * steps:]))


def cpu_sample_for(name):
    """Bounded CPU sample of the same workload family (per-cell metric): TGV → 128³, sphere → 128×64×64."""
    return "tgv128" if WORKLOADS[name]["kind"] == "tgv" else "sphere128"  # (the donut is timed on the sphere sample: same kernels)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = cpu_sample_for(args.workload)
    case = make_case(sample)
    cells = int(np.prod(case["dims"]))
    sec, threads, nv = oracle_run(case, args.steps, args.warmup)
    val = sec * 1e9 / cells
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample}: {args.steps} mom_step! of the oracle port (OpenMP, one parallel loop per reference @loop); "
                                       "Julia is absent so the reference itself cannot run"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="tgv512", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import wl_b200 as wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = make_case(args.workload)
    cells = int(np.prod(case["dims"]))
    nzl = case["dims"][2] // world

    def new_dist():
        """z-slab decomposition: rank 0 creates the NCCL id of the library's own communicator, torch broadcasts it (one id per handle)"""
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(wl.dist_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        return (rank, world, bytes(idt.cpu().numpy().tobytes()))

    distarg = new_dist()
    u0_host = tgv_u0(case["u0"][1], rank * nzl, nzl) if case["u0"] else None

    # ---- device-resident throughput --------------------------------------------------------------
    sim = build_sim(case, u0_host, distarg, local)
    fl = sim.flow
    check = wl.lib.check
    check(fl.L, fl.L.wl_sim_step_n(fl.h, args.warmup))
    fl.sync()
    import ctypes as C
    sp = C.c_void_p()
    check(fl.L, fl.L.wl_stream(fl.h, C.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_before = len(sim.pois.n)
    l_before = fl.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        e0.record(stream)
        check(fl.L, fl.L.wl_sim_step_n(fl.h, args.steps))
        e1.record(stream)
        fl.sync()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = fl.launches - l_before
    iters = sim.pois.n[n_before:]
    n_v = float(np.sum(iters)) / args.steps
    ns_cell = ms * 1e6 / args.steps / cells

    # ---- per-kernel CUDA-event pass over the same number of steps (roofline of the dominant kernel) ----
    fl.set_profiling(True)
    check(fl.L, fl.L.wl_sim_step_n(fl.h, args.steps))
    tim = fl.timings()
    fl.set_profiling(False)
    peak, peak_src = peaks()
    padded = int(np.prod([d + 2 for d in case["dims"]]))
    import ctypes as C2
    uni = C2.c_int(0)
    check(fl.L, fl.L.wl_is_const_coeff(fl.h, C2.byref(uni)))
    uni = bool(uni.value)
    tot_ms = sum(v[1] for v in tim.values())
    ksum = {}
    for k, (cnt, m, ncells) in sorted(tim.items(), key=lambda kv: -kv[1][1]):
        ent = {"launches": cnt, "ms": round(m, 3), "share": round(m / tot_ms, 4)}
        words = ALG_WORDS.get(k)
        if words and not uni and k in SEMI_WORDS:
            words = (SEMI_WORDS[k], words[1])  # general mode: all but a few per cent of the blocks are semi-uniform (no body nearby)
        if words and words[1 if uni else 0] > 0:
            w = words[1 if uni else 0]
            # the library reports, per kernel, the ghost-padded cells of the level every launch ran on (summed over the launches)
            per_launch = w * 4 * ncells / cnt
            ach = per_launch / (m / cnt * 1e-3) / 1e9
            ent.update(alg_bytes_per_launch=int(per_launch), achieved_gbs=round(ach, 1), frac=round(ach / peak, 4))
        ksum[k] = ent
    name = next(iter(ksum))  # dominant kernel = largest share of device time
    top = ksum[name]
    roof = {"bound": "hbm", "kernel": name, "achieved": top.get("achieved_gbs"), "peak": peak, "unit": "GB/s", "frac": top.get("frac"),
            "traffic": None, "peak_source": peak_src, "alg_bytes_per_launch": top.get("alg_bytes_per_launch"),
            "avg_launch_us": round(top["ms"] / top["launches"] * 1e3, 2),
            "note": "achieved = algorithmic bytes per launch / CUDA-event time on the library stream, averaged over the launches of the timed steps "
                    "(multigrid kernels: bytes and time summed over the levels they run on)"}
    if name in ("fm_conv", "fm_conv4"):
        roof["note"] += "; the flux kernel is FP32-issue-bound (9.4 QUICK face fluxes of ≈36 exact-IEEE operations per cell), not HBM-bound: see DESIGN.md"
    roof["traffic"], extra = ncu_traffic(args.workload, name)
    roof.update(extra)
    # whole-step roofline with the SURVEY §8d formulas
    if uni:
        b_alg, formula = 200 + 56 * n_v, "uniform-coefficient (no body, periodic): 200 + 56·n_V B/cell/step (SURVEY.md §8d)"
    else:
        b_alg = (468 if case["body"] else 264) + 120 * n_v
        formula = "general coefficients: %d + 120·n_V B/cell/step (SURVEY.md §8d)" % (468 if case["body"] else 264)
    step_ach = b_alg * padded / (ms * 1e-3 / args.steps) / 1e9
    roof["step"] = {"alg_bytes_per_cell": round(b_alg, 1), "n_V": round(n_v, 3), "achieved": round(step_ach, 1),
                    "frac": round(step_ach / (peak * world), 4), "peak_all_gpus": round(peak * world, 1), "formula": formula}

    line = {"metric": METRIC, "value": round(ns_cell, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "dims": list(case["dims"]), "periodic": list(case["perdir"]), "body": bool(case["body"]),
                       "poisson_iters_per_step": round(n_v, 3), "l2_flush": "state (%.1f GB) exceeds L2" % (padded * 32 * 4 / 1e9),
                       "kernels": "uniform-coefficient march kernels" if uni else "general variable-coefficient (semi-uniform blocks away from the body)",
                       "parallelism": "single GPU" if world == 1 else "z-slab x%d (halo planes, all-reduce and coarse-level all-gather over NVLink peer memory; NCCL for set-up; coarse levels replicated)" % world},
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "kernel_times": ksum}

    # ---- end to end through the host API: host u0 → Simulation → K × sim_step (+Δt readback) → u,p to host ----
    if not args.no_e2e:
        sim.close()
        del sim

        def e2e_leg(u0_arr):
            torch.cuda.synchronize()
            distarg2 = new_dist()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            sim2 = build_sim(case, u0_arr, distarg2, local)
            h2d = (u0_arr.nbytes if u0_arr is not None else 0)
            if case["body"]:
                h2d += 4 * padded * (3 + 9 + 3 + 1)
            t_setup = time.perf_counter() - t0
            for _ in range(args.warmup):
                wl.sim_step(sim2)
            sim2.flow.sync()
            t1 = time.perf_counter()
            last_dt = 0.0
            for _ in range(args.steps):
                wl.sim_step(sim2)
                last_dt = float(sim2.flow.Δt[-1])  # device→host read of the step's result
            u = sim2.flow.u
            p = sim2.flow.p
            t2 = time.perf_counter()
            e2e_s = (t2 - t1) + t_setup
            d2h = (u.nbytes + p.nbytes) * world + 4 * len(sim2.flow.Δt) * args.steps
            h2d *= world
            if world > 1:
                tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e_s = float(tt.item())
            sim2.close()
            return e2e_s, t_setup, h2d, d2h, last_dt

        pinned = torch.from_numpy(u0_host).pin_memory().numpy() if u0_host is not None else None
        e2e_s, t_setup, h2d, d2h, last_dt = e2e_leg(pinned)
        line["e2e"] = {"value": round(e2e_s * 1e9 / args.steps / cells, 5), "unit": UNIT, "h2d_bytes_per_step": int(h2d / args.steps),
                       "d2h_bytes_per_step": int(d2h / args.steps),
                       "what": "Simulation(host u0, pinned) set-up + K×sim_step with Δt read back each step + u,p downloaded to host, all timed "
                               "(max over ranks; every rank moves its own z slab; the process-wide NCCL communicator made for the first "
                               "handle is reused)",
                       "setup_s": round(t_setup, 3), "last_dt": last_dt}
        if u0_host is not None:  # the same from ordinary pageable memory (what a Julia Array is)
            e2e_p, t_setup_p, _, _, _ = e2e_leg(np.array(u0_host, copy=True))
            line["e2e"]["pageable_host_value"] = round(e2e_p * 1e9 / args.steps / cells, 5)
            line["e2e"]["pageable_setup_s"] = round(t_setup_p, 3)

    # ---- CPU baseline beside it (rank 0, bounded sample) ------------------------------------------------
    if not args.no_cpu and rank == 0:
        sample = cpu_sample_for(args.workload)
        ccase = make_case(sample)
        sec, threads, _ = oracle_run(ccase, 3, 1)
        line["cpu_baseline"] = {"value": round(sec * 1e9 / int(np.prod(ccase["dims"])), 3), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{sample}: 3 mom_step! of the oracle port after 1 warm-up (OpenMP over all host threads)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
