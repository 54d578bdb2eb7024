#!/usr/bin/env python3
"""Developer tool (multi-GPU box, under torchrun): NCCL all-gather latency for frame-sized buffers + peer-access facts."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    print("can_access_peer", [[torch.cuda.can_device_access_peer(i, j) if i != j else None for j in range(world)] for i in range(world)], flush=True)
for mb in (0.25, 1, 4, 8, 32):
    n = int(mb * 1024 * 1024 / 4)
    shard = torch.zeros(n, dtype=torch.int32, device="cuda"); out = torch.zeros(n * world, dtype=torch.int32, device="cuda")
    for _ in range(5): dist.all_gather_into_tensor(out, shard)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_gather_into_tensor(out, shard)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"all_gather {mb} MB/rank x{world}: {e0.elapsed_time(e1)/20*1000:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
