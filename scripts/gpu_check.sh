# quick GPU gate: parity tests + the headline bench (no CPU leg); summary of per-kernel times
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 400 gpurun_out/bench_quick.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e",{}).get("value"), "launches", d["gpu_launches"], "iters", d["config"].get("poisson_iters_per_step"))
for k,v in d["kernel_times"].items(): print(k.ljust(16), v["launches"], v["ms"], v["share"], v.get("frac"))
PY
