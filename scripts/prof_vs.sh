python -m pytest tests -m gpu -x -q -k "fused" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null > gpurun_out/bvs.json; python scripts/bench_brief.py gpurun_out/bvs.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"f_vsmooth|f_jacobi|k_small" -s 30 -c 12 --csv --log-file gpurun_out/vs_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
grep -E "f_vsmooth|f_jacobi|k_small" gpurun_out/vs_launches.csv | awk -F'","' '{print substr($5,1,30), $9, $NF}'
ncu --set full --clock-control none --import-source on -k regex:f_vsmooth -s 6 -c 1 -o gpurun_out/r1_v3_vsmooth python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b6.log 2>&1
