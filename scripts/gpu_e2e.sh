# host-transfer check: set-up probe, all GPU tests, the headline bench with its end-to-end legs
python scripts/setup_probe.py 2>&1 | tail -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
mkdir -p gpurun_out/q
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/q/bench_tgv512.json 2> gpurun_out/q/bench_tgv512.err; tail -3 gpurun_out/q/bench_tgv512.err
python - <<PY
import json
d=json.loads(open("gpurun_out/q/bench_tgv512.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"])
PY
