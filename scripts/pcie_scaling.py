"""What the host links give when N ranks copy a config-2 step's inputs in and results out at the same time, with NO kernels:
1,392,704 B host->device + 1,310,720 B device->host per step from/to pinned memory, two streams per rank (as prk_pipeline_host).
torchrun --nproc-per-node N scripts/pcie_scaling.py   (prints the slowest rank's microseconds per step)"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); lr = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
dev = torch.device('cuda', lr)
h_in = torch.empty(1392704, dtype=torch.uint8).pin_memory(); d_in = torch.empty(1392704, dtype=torch.uint8, device=dev)
h_out = torch.empty(1310720, dtype=torch.uint8).pin_memory(); d_out = torch.empty(1310720, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def loop(n, both=True):
    for _ in range(n):
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
for mode, both in (('h2d + d2h', True), ('h2d only', False)):
    loop(50, both); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(500, both)
    torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / 500 * 1e6
    if world > 1:
        t = torch.tensor([us], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); us = float(t.item())
    if rank == 0:
        gb = (1392704 + (1310720 if both else 0)) * world / us / 1e3
        print(f'{world} ranks, {mode}: {us:.1f} us per step (slowest rank) = {gb:.1f} GB/s aggregate', flush=True)
if world > 1: dist.destroy_process_group()
