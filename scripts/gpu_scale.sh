# weak-scaling bench at N GPUs (short run; the driver's own scaling run is the reference)
N=${N:-4}
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-30} --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.log | python scripts/show_bench.py
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
