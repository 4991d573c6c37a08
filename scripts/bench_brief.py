"""Prints a short per-kernel summary of a bench.py JSON line (stdin or file)."""
import json, signal, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads(txt.strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "launches", d["gpu_launches"], "iters",
      d["config"].get("poisson_iters_per_step"), "clk", d.get("clocks", {}).get("sm_mhz"))
for k, v in d.get("kernel_times", {}).items():
    print(" ", k.ljust(16), v["launches"], v["ms"], v["share"], v.get("frac"))
