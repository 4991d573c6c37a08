# round-1 profile set (run under gpurun, 1 GPU): launch list + full captures of the two dominant kernels of tgv512
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_tgv512.json 2> gpurun_out/bench_tgv512.err
python scripts/bench_brief.py gpurun_out/bench_tgv512.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r1_v3_launches_tgv512.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fm_conv4 -s 6 -c 2 -o gpurun_out/r1_v3_fm_conv4 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:f_vsmooth -s 8 -c 1 -o gpurun_out/r1_v3_vsmooth python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b3.log 2>&1
ncu --set full --clock-control none -k regex:"f_correct|f_div_residual|f_jacobi|f_cfl" -s 8 -c 5 -o gpurun_out/r1_v3_march python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b4.log 2>&1
