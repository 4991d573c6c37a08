# round-2 first GPU pass: all parity tests, bench lines of tgv512 / tgv128 / sphere, launch list at 128³, full capture of f_vsmooth
set -x
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2a/pytest.txt; cat gpurun_out/r2a/pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a/bench_tgv512.json 2> gpurun_out/r2a/bench_tgv512.err
python scripts/bench_brief.py gpurun_out/r2a/bench_tgv512.json
python bench.py --workload tgv128 --steps 100 --warmup 5 --no-cpu > gpurun_out/r2a/bench_tgv128.json 2> gpurun_out/r2a/bench_tgv128.err
python scripts/bench_brief.py gpurun_out/r2a/bench_tgv128.json
python bench.py --workload sphere --steps 20 --warmup 3 --no-cpu > gpurun_out/r2a/bench_sphere.json 2> gpurun_out/r2a/bench_sphere.err
python scripts/bench_brief.py gpurun_out/r2a/bench_sphere.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r2a/launches_tgv128.csv python bench.py --workload tgv128 --steps 4 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2a/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:f_vsmooth -s 8 -c 1 -o gpurun_out/r2a/vsmooth python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a/b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fm_conv4 -s 6 -c 1 -o gpurun_out/r2a/fm_conv4 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a/b2.log 2>&1
ls -la gpurun_out/r2a
