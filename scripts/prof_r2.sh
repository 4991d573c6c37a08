# round-2 profile set, one GPU: bench lines (tgv512 headline, tgv128 = configs[1], sphere = configs[2]), launch list, full captures
set -x
O=gpurun_out/r2
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_tgv512.json 2> $O/bench_tgv512.err
python scripts/bench_brief.py $O/bench_tgv512.json
python bench.py --workload tgv128 --steps 100 --warmup 5 > $O/bench_tgv128.json 2> $O/bench_tgv128.err
python scripts/bench_brief.py $O/bench_tgv128.json
python bench.py --workload sphere --steps 20 --warmup 3 > $O/bench_sphere.json 2> $O/bench_sphere.err
python scripts/bench_brief.py $O/bench_sphere.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_tgv512.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fm_conv4 -s 4 -c 3 -o $O/fm_conv4 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $O/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:f_vsmooth -s 8 -c 1 -o $O/vsmooth python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $O/b3.log 2>&1
ncu --set full --clock-control none -k regex:"f_correct_cfl|f_divres_uni|f_jacobi_uni2|k_tiny_uni|k_small_levels" -s 8 -c 7 -o $O/march python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $O/b4.log 2>&1
ncu --set full --clock-control none -k regex:"fm_conv|f_gs_half|k_bdim2|f_increment|f_jacobi|f_div_residual" -s 30 -c 12 -o $O/sphere_kernels python bench.py --workload sphere --steps 1 --warmup 2 --no-cpu --no-e2e > $O/b5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file $O/launches_tgv128.csv python bench.py --workload tgv128 --steps 4 --warmup 5 --no-cpu --no-e2e > $O/b6.log 2>&1
for r in fm_conv4 vsmooth march sphere_kernels; do python scripts/ncu_summary.py $O/$r.ncu-rep > $O/$r.txt; done
rm -f $O/vsmooth.ncu-rep $O/march.ncu-rep $O/sphere_kernels.ncu-rep   # (gpurun brings back at most 64 MiB: the summaries travel, one report stays for source-level reading)
ls -la $O
