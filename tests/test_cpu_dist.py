"""CPU: host-side logic of the N>1 path (row-cyclic sharding, one all-gather, de-interleave) with gloo, world_size 2.
The renderer inside each rank is the CPU restatement (no GPU here); on GPUs bench.py drives the same steps over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_and_deinterleave_roundtrip():
    from renderer_b200 import dist
    H, W = 1080, 64
    img = np.arange(H * W, dtype=np.uint32).reshape(H, W)
    for P in (1, 2, 3, 4, 7, 8):
        shards = [dist.pack_shard(img[dist.shard_rows(H, P, r)], H, P) for r in range(P)]
        assert all(s.shape[0] == dist.rows_per_shard(H, P) for s in shards)
        assert np.array_equal(dist.deinterleave(np.concatenate(shards, 0), H, P), img)
    assert sorted(sum((dist.shard_rows(10, 3, r) for r in range(3)), [])) == list(range(10))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as td
    import renderer_b200 as rb
    from oracle import pyport
    from renderer_b200 import dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        path = pyport.model_path("torus.ply")
        s = rb.Scene(path).UpdateBoundingVolumeHierarchy(path + ".bvh")
        cam = rb.Orbit.cameras([4])[4]
        W, H = 96, 70                                             # H % world != 0: exercises the padding
        f = rb.make_frame(9, W, H, cam, row_first=rank, row_step=world)
        mine = pyport.render(s, f).view(np.int32)
        shard = torch.from_numpy(dist.pack_shard(mine, H, world).copy())
        gathered = dist.all_gather_rows(shard)
        frame = dist.deinterleave(gathered.numpy().view(np.uint32), H, world)
        full = pyport.render(s, rb.make_frame(9, W, H, cam))
        q.put((rank, bool(np.array_equal(frame, full)), int((frame != 0).sum())))
    finally:
        td.destroy_process_group()


def test_two_rank_gloo_frame_assembly(pyport):
    import torch.multiprocessing as mp
    if not os.path.exists(pyport.model_path("torus.ply")):
        pytest.skip("model not staged")
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0)); port = so.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    assert all(n > 0 for _, _, n in res)


def test_pipeline_api_refuses_bad_arguments_without_a_device(rb):
    """The C-ABI frame pipeline (b200r_pipeline_*, csrc/cuda/dist.cu) checks its arguments before touching a device; without a
    context there is nothing it could run on - it must fail, not fall back."""
    import ctypes as C
    L = rb.lib()
    out = C.c_void_p()
    assert L.b200r_pipeline_create(None, 64, 48, 2, 0, 1, None, rb.ASSEMBLE_PUSH, C.byref(out)) != 0
    assert not out.value
    assert L.b200r_pipeline_submit(None, None, None) != 0
    assert L.b200r_pipeline_drain(None) != 0
    L.b200r_pipeline_destroy(None)                     # no-op


def test_row_ownership_matches_the_device_convention():
    """b200r_pipeline stamps row_first = rank, row_step = world into every frame (dist.cu); the host mirror deals the same rows."""
    from renderer_b200 import dist
    for H, P in ((1080, 8), (2160, 8), (603, 4), (7, 3)):
        for r in range(P):
            assert dist.shard_rows(H, P, r) == [r + k * P for k in range((H - r + P - 1) // P)]
