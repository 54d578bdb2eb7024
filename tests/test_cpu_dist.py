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


def test_frame_pipeline_orders_its_stages(monkeypatch):
    """renderer_b200.dist.FramePipeline on a recording stand-in for torch.cuda / NCCL (no GPU): per frame the render stream waits
    for its slot's previous frame, the all-gather follows the render through an event on the communication stream, the
    de-interleave follows the all-gather, the copy-out follows the assembly on its own stream, and the slot is only released
    after the copy - for every slot of a 3-deep pipeline on rank 0 of 2."""
    import sys
    import types
    log = []

    class Event:
        n = 0

        def __init__(self, **kw):
            Event.n += 1
            self.id = Event.n

        def record(self, stream):
            log.append(("record", self.id, stream.name))

    class Stream:
        n = 0

        def __init__(self, priority=0):
            Stream.n += 1
            self.name = f"s{Stream.n}"
            self.cuda_stream = 1000 + Stream.n

        def wait_event(self, ev):
            log.append(("wait_event", ev.id, self.name))

        def wait_stream(self, other):
            log.append(("wait_stream", other.name, self.name))

        def synchronize(self):
            log.append(("sync", self.name))

    class Tensor:
        def __init__(self, name):
            self.name = name

        def data_ptr(self):
            return hash(self.name) & 0xffff

        def pin_memory(self):
            return self

        def copy_(self, src, non_blocking=False):
            log.append(("copy", src.name, self.name, current[-1].name))

    current = [types.SimpleNamespace(name="default")]
    counter = {"t": 0}

    class StreamCtx:
        def __init__(self, s):
            self.s = s

        def __enter__(self):
            current.append(self.s)

        def __exit__(self, *a):
            current.pop()

    def zeros(shape, **kw):
        counter["t"] += 1
        return Tensor(f"t{counter['t']}_{shape[0]}x{shape[1]}")

    cuda = types.SimpleNamespace(Stream=Stream, Event=Event, stream=StreamCtx, current_device=lambda: 0, synchronize=lambda: None)
    torch = types.ModuleType("torch")
    torch.cuda, torch.int32, torch.zeros, torch.device = cuda, "int32", zeros, (lambda *a: "cuda:0")
    tdist = types.ModuleType("torch.distributed")
    tdist.all_gather_into_tensor = lambda out, inp, group=None: log.append(("all_gather", inp.name, out.name, current[-1].name))
    torch.distributed = tdist
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.distributed", tdist)

    class Gpu:
        def render_device_slot(self, frame, ptr, stream, slot):
            log.append(("render", frame, stream, slot))

        def deinterleave_device(self, g, f, W, H, P, stream):
            log.append(("deinterleave", stream))

        def last_launches(self):
            return 2

    from renderer_b200.dist import FramePipeline
    D = 3
    pipe = FramePipeline(Gpu(), 64, 48, rank=0, world=2, depth=D, to_host=True)
    comm, copy = pipe.comm, pipe.copy
    for i in range(2 * D + 1):
        del log[:]
        d = pipe.submit(f"frame{i}")
        assert d == i % D
        rs = pipe.render_streams[d]
        kinds = [e[0] for e in log]
        # the order of everything enqueued for this frame
        assert kinds == (["wait_event"] if i >= D else []) + ["render", "record", "wait_event", "all_gather", "deinterleave",
                                                               "record", "wait_event", "copy", "record"]
        it = iter(log)
        if i >= D:
            assert next(it) == ("wait_event", pipe.free[d].id, rs.name)              # slot d's previous frame has been copied out
        assert next(it) == ("render", f"frame{i}", rs.cuda_stream, d)
        assert next(it) == ("record", pipe.rendered[d].id, rs.name)
        assert next(it) == ("wait_event", pipe.rendered[d].id, comm.name)
        assert next(it) == ("all_gather", pipe.shard[d].name, pipe.gathered[d].name, comm.name)
        assert next(it) == ("deinterleave", comm.cuda_stream)
        assert next(it) == ("record", pipe.assembled[d].id, comm.name)
        assert next(it) == ("wait_event", pipe.assembled[d].id, copy.name)
        assert next(it) == ("copy", pipe.full[d].name, pipe.host[d].name, copy.name)
        assert next(it) == ("record", pipe.free[d].id, copy.name)
    assert pipe.launches == (2 + 1) * (2 * D + 1)
    # one GPU, frames staying on the device: no communication stream work at all, the slot is released by the render stream
    pipe1 = FramePipeline(Gpu(), 64, 48, depth=2)
    del log[:]
    pipe1.submit("f0")
    assert [e[0] for e in log] == ["render", "record", "record"] and log[-1] == ("record", pipe1.free[0].id, pipe1.render_streams[0].name)
