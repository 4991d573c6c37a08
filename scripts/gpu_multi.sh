# multi-GPU pass (run with gpurun --gpus N): 2-rank bit-identity tests when N>=2, then the bench at N ranks
N=${N:-2}
mkdir -p gpurun_out/m
python -m pytest tests -m gpu -q -x -k "two_rank" 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/m/bench_n$N.json 2> gpurun_out/m/bench_n$N.err; tail -3 gpurun_out/m/bench_n$N.err
python scripts/bench_brief.py gpurun_out/m/bench_n$N.json
