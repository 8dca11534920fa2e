# Round-2 multi-GPU validation (gpurun --gpus N): world-2 sharded-equals-single test over CUDA IPC and NCCL, bench at N ranks.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 800 -k "world2 or peer or exchange" 2>&1 | grep -v Warning | grep -v "data = " | tail -40 > gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
for T in auto nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --transport $T 2> gpurun_out/bench_n${N}_$T.err > gpurun_out/bench_n${N}_$T.log
tail -c 400 gpurun_out/bench_n${N}_$T.err
python scripts/show_bench.py < gpurun_out/bench_n${N}_$T.log
done
