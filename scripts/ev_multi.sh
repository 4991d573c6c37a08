# round-2 evidence, N GPUs (gpurun --gpus N): bit-identity against one GPU, then bench lines of the z-slab workloads
N=${N:-2}
O=gpurun_out/ev
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
python -m pytest tests -m gpu -q -x -k "two_rank" 2>&1 | tail -3 > $O/two_rank_n$N.txt; cat $O/two_rank_n$N.txt
$TR scripts/dist_check.py 64 10 tgv 2>/dev/null | grep -v "^\*\|OMP_NUM" > $O/dist_check_n$N.txt
$TR scripts/dist_check.py 128 10 tgv 2>/dev/null | grep -v "^\*\|OMP_NUM" >> $O/dist_check_n$N.txt
$TR scripts/dist_check.py 64 10 sphere 2>/dev/null | grep -v "^\*\|OMP_NUM" >> $O/dist_check_n$N.txt
cat $O/dist_check_n$N.txt
for w in ${WORKLOADS:-tgv512 tgv1024 sphere}; do
  $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --workload $w > $O/bench_${w}_${N}gpu.json 2> $O/bench_${w}_${N}gpu.err; tail -2 $O/bench_${w}_${N}gpu.err | cut -c1-300
  python scripts/bench_brief.py $O/bench_${w}_${N}gpu.json | head -8
done
