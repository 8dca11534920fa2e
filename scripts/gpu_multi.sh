# N-GPU checks on one box: sharded correctness, weak-scaling bench, reference arm under torchrun.
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | tail -3 | tee gpurun_out/mgpu_check.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.log | python scripts/show_bench.py
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_ref_n$N.err | tee gpurun_out/bench_ref_n$N.log | python scripts/show_bench.py
