"""torchrun diagnostic: what does the result gather cost at N ranks?  NCCL all_gather_into_tensor of the packed
result block vs nothing, with CUDA events; prints one line per rank-0 measurement."""
import os, sys, time
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for rows in (1250, 10000):
    nbytes = rows * 10 * 8
    mine = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    allb = torch.zeros(world * nbytes, dtype=torch.uint8, device=dev)
    for _ in range(5):
        dist.all_gather_into_tensor(allb, mine)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        dist.all_gather_into_tensor(allb, mine)
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / 50], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[diag] world {world} all_gather {nbytes} B/rank: {float(t[0])*1e3:.1f} us per call", flush=True)
    # the same with a dependent tiny kernel in between (launch-rate check)
    s.record()
    for _ in range(50):
        mine.add_(1)
        dist.all_gather_into_tensor(allb, mine)
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / 50], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[diag] world {world} kernel+all_gather {nbytes} B/rank: {float(t[0])*1e3:.1f} us per step", flush=True)
dist.barrier(); dist.destroy_process_group()
