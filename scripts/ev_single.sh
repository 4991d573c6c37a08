# round-2 evidence, one GPU: the profile set, then the big single-GPU rows (TGV 1024³, donut)
bash scripts/prof_r2.sh
O=gpurun_out/ev
mkdir -p $O
for w in tgv1024 donut; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > $O/bench_${w}_1gpu.json 2> $O/bench_${w}_1gpu.err; tail -2 $O/bench_${w}_1gpu.err | cut -c1-300
  python scripts/bench_brief.py $O/bench_${w}_1gpu.json | head -8
done
