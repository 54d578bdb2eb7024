#!/usr/bin/env python3
"""Multi-GPU parity check of the pipelined frame assembly (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank renders its row-cyclic shard of 9 orbit frames with 3 frames in flight (renderer_b200.dist.FramePipeline: render
streams + ONE NCCL all-gather + de-interleave per frame on a communication stream) and compares every assembled frame, bit for
bit, with the same frame rendered whole on its own GPU by the blocking call. Prints one line per rank; exit code 1 on a mismatch.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import renderer_b200 as rb
    from renderer_b200.dist import FramePipeline
    from oracle import pyport          # model staging paths only
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    from renderer_b200.dist import init_nccl
    init_nccl(local)
    model = pyport.model_path("chessboard.tri")
    scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
    gpu = rb.Renderer(local)
    gpu.upload(scene)
    W, H, n, depth = 1920, 1080, 9, 3
    cams = rb.Orbit.cameras(range(n))
    whole = [gpu.render(rb.make_frame(rb.MODE_RAYTRACE, W, H, cams[k], flags=1 | 4, frame_index=k)).copy() for k in range(n)]
    pipe = FramePipeline(gpu, W, H, rank=rank, world=world, depth=depth, to_host=(rank == 0))
    bad = 0
    for base in range(0, n, depth):
        slots = [pipe.submit(rb.make_frame(rb.MODE_RAYTRACE, W, H, cams[k], flags=1 | 4, frame_index=k,
                                           row_first=rank, row_step=world)) for k in range(base, base + depth)]
        pipe.drain()
        for k, d in zip(range(base, base + depth), slots):
            got = pipe.full[d].cpu().numpy().view(np.uint32)
            bad += int((got != whole[k]).sum())
            if rank == 0:
                bad += int((pipe.host[d].numpy().view(np.uint32) != whole[k]).sum())
    t = torch.tensor([bad], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    print(f"dist_check rank {rank}/{world}: {n} frames, {depth} in flight, differing pixels on this rank: {bad}, on all ranks: {int(t.item())}",
          flush=True)
    dist.barrier()
    dist.destroy_process_group()
    gpu.close()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
