"""ctypes binding of include/wl_b200.h.  Fails loudly when the CUDA library is missing or unusable."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_lib = {}


class WLError(RuntimeError):
    pass


class Config(C.Structure):
    """struct wl_config (include/wl_b200.h)"""
    _fields_ = [("D", C.c_int32), ("n", C.c_int32 * 3), ("uBC", C.c_float * 3), ("perdir", C.c_int32 * 3),
                ("exitBC", C.c_int32), ("lam", C.c_int32), ("nu", C.c_float), ("dt0", C.c_float),
                ("pois_kind", C.c_int32), ("smoother", C.c_int32), ("tol", C.c_float), ("itmx", C.c_int32),
                ("device", C.c_int32), ("flags", C.c_int32)]


class BodyPrim(C.Structure):
    """struct wl_body_prim (include/wl_b200.h)"""
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("center", C.c_float * 3), ("R", C.c_float), ("r", C.c_float),
                ("vel", C.c_float * 3)]


# -fmad=false: the production library executes plain IEEE Float32 operations in the reference's order, which
# makes it bit-identical to the CPU oracle (compiled with -ffp-contract=off) over whole simulations — see
# DESIGN.md §numerics.  The HBM-bound kernels do not pay for it; the flux kernel (FP32-issue-bound) pays ≈20 % more
# instructions than an FMA-contracted build — the price of holding the 1e-5 gate after 100 steps.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-pthread", "-shared"]


def library_path(fmad=False):
    if os.environ.get("WL_B200_LIB"):  # A/B experiments against another build of the same ABI
        return os.environ["WL_B200_LIB"]
    return os.path.join(_CSRC, "libwl_b200_fmad.so" if fmad else "libwl_b200.so")


def build_library(fmad=False, force=False, verbose=False):
    """nvcc-compile csrc/wl_b200.cu for sm_100a in-tree.  `fmad=True` builds an FMA-contracted variant that exists
    only for performance experiments (bench.py --fmad); it is not bit-identical to the oracle."""
    exact = fmad
    out = library_path(fmad)
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "wl_b200.h"))
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not (fmad and f == "-fmad=false")]
    cmd = [nvcc] + flags + ["-o", out, os.path.join(_CSRC, "wl_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return out


def load_library(fmad=False):
    key = bool(fmad)
    if key in _lib:
        return _lib[key]
    path = library_path(fmad)
    if not os.path.exists(path):
        raise WLError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback for the mom_step! path)")
    L = C.CDLL(path)
    H = C.c_void_p
    fp = C.POINTER(C.c_float)
    L.wl_last_error.restype = C.c_char_p
    L.wl_device_count.restype = C.c_int
    sig = {
        "wl_create": [C.POINTER(Config), C.POINTER(H)],
        "wl_destroy": [H],
        "wl_release_pool": [],
        "wl_upload": [H, C.c_int, C.c_void_p, C.c_int],
        "wl_download": [H, C.c_int, C.c_void_p, C.c_int],
        "wl_upload_component": [H, C.c_int, C.c_int, C.c_void_p, C.c_int],
        "wl_apply_bc": [H],
        "wl_update": [H],
        "wl_measure_bc": [H],
        "wl_mom_step": [H],
        "wl_sim_step_n": [H, C.c_int],
        "wl_sim_step_until": [H, C.c_double, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_int64)],
        "wl_project": [H, C.c_float],
        "wl_conv_diff": [H, C.c_int],
        "wl_cfl": [H, fp],
        "wl_pois_mult": [H],
        "wl_pois_solve": [H, C.POINTER(C.c_int)],
        "wl_pois_residual": [H, fp],
        "wl_pois_smooth": [H, C.c_int, C.c_int, C.c_float],
        "wl_pois_vcycle": [H, C.c_float],
        "wl_num_levels": [H, C.POINTER(C.c_int)],
        "wl_level_dims": [H, C.c_int, C.POINTER(C.c_int32)],
        "wl_download_level": [H, C.c_int, C.c_int, fp],
        "wl_upload_level": [H, C.c_int, C.c_int, fp],
        "wl_get_dt": [H, fp, C.POINTER(C.c_int)],
        "wl_set_dt": [H, fp, C.c_int],
        "wl_get_iters": [H, C.POINTER(C.c_int16), C.POINTER(C.c_int)],
        "wl_get_solver_log": [H, fp, C.POINTER(C.c_int)],
        "wl_set_logging": [H, C.c_int],
        "wl_time": [H, C.POINTER(C.c_double)],
        "wl_sync": [H],
        "wl_set_profiling": [H, C.c_int],
        "wl_get_timings": [H, C.c_char_p, C.POINTER(C.c_int)],
        "wl_launch_count": [H, C.POINTER(C.c_int64)],
        "wl_is_const_coeff": [H, C.POINTER(C.c_int)],
        "wl_set_tuning": [H, C.c_char_p, C.c_int],
        "wl_set_body": [H, C.c_void_p, C.c_int, C.c_float],
        "wl_measure": [H, C.c_float],
        "wl_set_remeasure": [H, C.c_int],
        "wl_time_next": [H, C.POINTER(C.c_double)],
        "wl_body_forces": [H, C.POINTER(C.c_float), C.POINTER(C.c_double)],
        "wl_set_forcing": [H, fp, fp, fp, fp],
        "wl_set_sgs": [H, C.c_float, C.c_float],
        "wl_meanflow_init": [H, C.c_int],
        "wl_meanflow_update": [H],
        "wl_meanflow_reset": [H, C.c_float],
        "wl_meanflow_copy_to_flow": [H],
        "wl_meanflow_download": [H, C.c_int, C.c_void_p, C.c_int],
        "wl_meanflow_upload": [H, C.c_int, C.c_void_p, C.c_int],
        "wl_meanflow_get_times": [H, fp, C.POINTER(C.c_int)],
        "wl_meanflow_set_times": [H, fp, C.c_int],
        "wl_stream": [H, C.POINTER(C.c_void_p)],
    }
    sig["wl_selftest_div6"] = [C.POINTER(C.c_uint64)]
    sig["wl_dist_unique_id"] = [C.c_void_p]
    sig["wl_create_dist"] = [C.POINTER(Config), C.c_int, C.c_int, C.c_void_p, C.POINTER(H)]
    for name, args in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export what the header declares
        fn.argtypes = args
        fn.restype = C.c_int
    L._wl_symbols = list(sig) + ["wl_last_error", "wl_device_count"]
    _lib[key] = L
    return L


FLAGS = {"general_coeff": 1, "unfused_gs": 2, "no_persistent": 4, "nccl_halo": 8, "no_vsmooth": 16, "no_conv4": 32,
         "no_fused_uni": 64, "no_semi": 128, "no_tiny": 256, "no_fast_read": 512, "no_prefetch": 1024, "nccl_allreduce": 2048, "no_pdl": 4096}  # WL_FLAG_* (include/wl_b200.h)


def check(L, rc):
    if rc != 0:
        raise WLError(L.wl_last_error().decode())


def release_pool(fmad=False):
    """Returns the device memory pooled from closed simulations to the driver (wl_release_pool)."""
    load_library(fmad).wl_release_pool()
