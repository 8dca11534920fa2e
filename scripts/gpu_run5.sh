mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py > gpurun_out/mgpu_check.log 2>&1; echo "check exit $?" >> gpurun_out/mgpu_check.log; tail -3 gpurun_out/mgpu_check.log
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; tail -1 gpurun_out/bench_n$N.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_ref_n$N.log 2>&1; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
