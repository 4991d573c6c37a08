# compute-sanitizer passes over small cases that reach every production kernel family: memcheck, racecheck (shared-memory
# hazards of the fm_conv4 / f_vsmooth rings and k_tiny_uni), initcheck on the fused paths
mkdir -p gpurun_out/san
T="tests/test_gpu_parity.py"
K='test_fused_upstroke_matches_oracle or (test_mom_step_two_steps and quick) or test_device_measure_matches_oracle or test_tiny_velocities or (test_limiters_on_the_hot_flux_kernels and quick) or (test_tiny_level_kernels_equal_the_cooperative_kernel and tgv_general_coeff) or (test_sgs_udf_matches_oracle and sphere_exit_3d)'
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 400 python -m pytest $T -q -x -k "$K" > gpurun_out/san/$tool.txt 2>&1
  echo "$tool rc=$?"; tail -6 gpurun_out/san/$tool.txt
done
