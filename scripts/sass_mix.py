"""Instruction mix of the hot kernels from the built library's SASS (cuobjdump -sass): python scripts/sass_mix.py > profiles/r2_sass_mix.txt
Counts are static instructions per kernel (not executed counts): they show WHICH machine instructions the kernels are made of —
LDGSTS (cp.async global→shared fills), LDS/STS.128, packed FP32 pairs (FADD2/FMUL2/FFMA2), three-input FMNMX3, warp shuffles, barriers."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "waterlily.jl_b200/csrc/libwl_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["fm_conv4ILi0ELb0", "f_vsmoothILb1ELb0ELi1", "f_vsmoothILb0ELb0ELi2", "f_correct_cflILb1", "f_divres_uni", "f_jacobi_uni2", "k_tiny_uni", "k_small_levelsILb1",
        "fm_convILi0ELb0ELb0", "k_halo_push", "k_allreduce", "k_bcast_planes"]
cur, mix = None, collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        mix[cur][base] += 1
        if base in ("LDS", "STS", "LDG", "STG") and ".128" in op:
            mix[cur][base + ".128"] += 1
keys = ["total", "FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2", "FMNMX", "FMNMX3", "FSEL", "FSETP", "LDGSTS", "LDS", "LDS.128", "STS", "STS.128", "LDG", "LDG.128", "STG", "STG.128",
        "SHFL", "BAR", "LDC", "LDCU", "IMAD", "IADD3", "ISETP", "UBLKCP", "UTMALDG", "SYNCS", "DADD", "DFMA", "DMUL", "MUFU", "ATOMG", "RED", "MEMBAR", "ERRBAR"]
print("static SASS instruction mix, sm_100a build of", lib)
print("%-44s" % "kernel" + "".join("%8s" % k for k in keys))
for fn, c in mix.items():
    if not any(w in fn for w in want):
        continue
    c["total"] = sum(v for k, v in c.items() if k != "total" and "." not in k)
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0][:43]
    print("%-44s" % name + "".join("%8d" % c.get(k, 0) for k in keys))
