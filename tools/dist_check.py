#!/usr/bin/env python3
"""Multi-GPU parity check of the frame pipeline (b200r_pipeline_*), run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [c2|c5] [mlaa]

Every rank renders its row-cyclic shard of 9 orbit frames with 3 frames in flight, once per assembly mode (one NCCL all-gather +
de-interleave; peer stores over NVLink), and compares every assembled frame - on every rank - bit
for bit (through the copy-out to page-locked host memory, on every rank) with the same frame rendered whole on its own GPU by the blocking call. `mlaa`: with the MLAA filter, which then runs on
the assembled frame. Prints one line per rank and mode; exit code 1 on a mismatch.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import renderer_b200 as rb
    from bench import WORKLOADS
    from oracle import pyport          # model staging paths only
    wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in WORKLOADS else "c2"]
    mlaa = "mlaa" in sys.argv[1:]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    from renderer_b200.dist import init_nccl
    init_nccl(local)
    model = pyport.model_path(wl["model"])
    scene = rb.Scene(model).UpdateBoundingVolumeHierarchy(model + ".bvh")
    gpu = rb.Renderer(local)
    gpu.upload(scene)
    W, H, depth = wl["W"], wl["H"], 3
    n = 9 if wl["ao"] == 0 else 3
    flags = wl["flags"] | (rb.F_MLAA if mlaa else 0)
    cams = rb.Orbit.cameras(range(n))
    frames = [rb.make_frame(wl["mode"], W, H, cams[k], flags=flags, ao_samples=wl["ao"] or 32, frame_index=k) for k in range(n)]
    whole = [gpu.render(f).copy() for f in frames]
    total_bad = 0
    for mode, name in ((rb.ASSEMBLE_NCCL, "nccl all-gather"), (rb.ASSEMBLE_PUSH, "peer push")):
        uid = [rb.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pipe = rb.Pipeline(gpu, W, H, depth=depth, rank=rank, world=world, unique_id=uid[0], assemble=mode)
        hosts = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(depth)]
        bad = 0
        for base in range(0, n, depth):
            ks = list(range(base, min(n, base + depth)))
            for k in ks:
                pipe.submit(frames[k], hosts[k % depth].data_ptr())      # the assembled frame, copied out on EVERY rank
            pipe.drain()
            dist.barrier()             # every rank has pushed AND consumed these frames before anyone moves on
            for k in ks:
                bad += int((hosts[k % depth].numpy().view(np.uint32) != whole[k]).sum())
            dist.barrier()
        t = torch.tensor([bad], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        print(f"dist_check rank {rank}/{world} [{wl['desc'][:40]}{' +MLAA' if mlaa else ''}] {name}: {n} frames, {depth} in flight, "
              f"differing pixels on this rank: {bad}, on all ranks: {int(t.item())}", flush=True)
        total_bad += int(t.item())
        dist.barrier()
        pipe.close()
    dist.barrier()
    dist.destroy_process_group()
    gpu.close()
    sys.exit(1 if total_bad else 0)


if __name__ == "__main__":
    main()
