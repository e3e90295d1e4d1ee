"""Bandwidth of the final-gather collectives at the bench's message size (26 MB per rank), device-timed:
torchrun --nproc-per-node N tools/micro/nccl_gather_bw.py"""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = 4096 * 200 * 4  # doubles per rank: K (2) + k (1) + sigK (1) of 4096 x 200 pendulum controllers = 26 MB
x = torch.randn(n, dtype=torch.float64, device="cuda")
out = torch.empty(n * world, dtype=torch.float64, device="cuda")
for name in ("all_gather", "gather0"):
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            if name == "all_gather":
                dist.all_gather_into_tensor(out, x)
            else:
                dist.gather(x, [out[i * n:(i + 1) * n] for i in range(world)] if rank == 0 else None, dst=0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        print(f"{name}: world {world}, {n * 8 / 1e6:.1f} MB per rank: {ms:.3f} ms per call, "
              f"{(world - 1) * n * 8 / 1e9 / (ms * 1e-3):.0f} GB/s into rank 0", flush=True)
dist.destroy_process_group()
