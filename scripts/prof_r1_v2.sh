set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r1_v2_launches_tgv512.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fm_conv -s 4 -c 2 -o gpurun_out/r1_v2_fm_conv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"f_gs_half|f_increment|f_jacobi|f_gs_a" -s 0 -c 10 -o gpurun_out/r1_v2_vcycle python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bc_vec" -s 4 -c 2 -o gpurun_out/r1_v2_bc python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b4.log 2>&1
ls -la gpurun_out
