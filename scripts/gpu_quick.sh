# quick GPU gate: parity tests + bench lines (no CPU leg)
mkdir -p gpurun_out/q
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/q/bench_tgv512.json 2> gpurun_out/q/bench_tgv512.err; tail -3 gpurun_out/q/bench_tgv512.err
python scripts/bench_brief.py gpurun_out/q/bench_tgv512.json
python bench.py --workload tgv128 --steps 100 --warmup 5 --no-cpu --no-e2e > gpurun_out/q/bench_tgv128.json 2> gpurun_out/q/bench_tgv128.err; tail -3 gpurun_out/q/bench_tgv128.err
python scripts/bench_brief.py gpurun_out/q/bench_tgv128.json
if [ -n "$SPHERE" ]; then
python bench.py --workload sphere --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/q/bench_sphere.json 2> gpurun_out/q/bench_sphere.err; tail -3 gpurun_out/q/bench_sphere.err
python scripts/bench_brief.py gpurun_out/q/bench_sphere.json
fi
if [ -n "$NCU_K" ]; then
ncu --set full --clock-control none --import-source on -k regex:$NCU_K -s ${NCU_S:-6} -c ${NCU_C:-1} -o gpurun_out/q/ncu_k python bench.py --workload ${NCU_W:-tgv512} --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/q/ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/q/ncu_k.ncu-rep > gpurun_out/q/ncu_k.txt; rm -f gpurun_out/q/ncu_k.ncu-rep
fi
