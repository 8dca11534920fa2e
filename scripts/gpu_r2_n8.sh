N=${N:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/bench_n${N}.err > gpurun_out/bench_n${N}.log
tail -c 300 gpurun_out/bench_n${N}.err
python scripts/show_bench.py < gpurun_out/bench_n${N}.log | cut -c1-330
