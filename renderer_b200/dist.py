"""Multi-GPU frame assembly: row-cyclic sharding + ONE all-gather + de-interleave.

Rank r of P renders rows r, r+P, r+2P, ... (the orbit view is ~90 % background, so contiguous bands would
be badly unbalanced) into a packed buffer of ceil(H/P) rows; one all-gather of the packed rows (NCCL over
NVLink on GPUs; gloo on CPU for the host-logic tests) gives every rank all shards; a de-interleave pass
restores scan order. Nothing else is exchanged (the scene is replicated, the Z-buffer never leaves a GPU).
"""
import numpy as np


def rows_per_shard(height, n_shards):
    return (height + n_shards - 1) // n_shards


def shard_rows(height, n_shards, rank):
    """The screen rows rank `rank` owns."""
    return list(range(rank, height, n_shards))


def pack_shard(rows_img, height, n_shards):
    """Pad a rank's packed rows to rows_per_shard (all-gather needs equal counts)."""
    rps = rows_per_shard(height, n_shards)
    if rows_img.shape[0] == rps:
        return rows_img
    out = np.zeros((rps, rows_img.shape[1]), dtype=rows_img.dtype)
    out[: rows_img.shape[0]] = rows_img
    return out


def deinterleave(gathered, height, n_shards):
    """gathered: (n_shards*rows_per_shard, W) as produced by the all-gather -> (height, W) in scan order.
    Host-side mirror of b200r_deinterleave_device."""
    rps = rows_per_shard(height, n_shards)
    g = np.asarray(gathered).reshape(n_shards, rps, -1)
    out = np.empty((height, g.shape[2]), dtype=g.dtype)
    for s in range(n_shards):
        rows = shard_rows(height, n_shards, s)
        out[rows] = g[s, : len(rows)]
    return out


def all_gather_rows(shard, group=None):
    """torch.distributed all-gather of equally sized packed shards -> stacked tensor (P*rps, W)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * shard.shape[0], shard.shape[1]), dtype=shard.dtype, device=shard.device)
    try:
        dist.all_gather_into_tensor(out, shard.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard.contiguous(), group=group)
        out = torch.cat(parts, 0)
    return out


class FramePipeline:
    """Frames in flight on one rank of P (P = 1 included): the multi-GPU path of SURVEY.md §8e with its stages overlapped.

    Frame i uses slot d = i % depth. Per slot: a render stream, a set of the library's per-frame scratch buffers
    (b200r_render_device_slot), a packed shard, the all-gathered shards and the assembled frame. Stage order per frame:

        render stream d : [wait: slot d's previous frame has left its buffers] -> this rank's rows of frame i
        comm stream     : [wait: rendered] -> ONE all-gather of the packed rows (NCCL) -> de-interleave
        copy stream     : (optional, the `to_host` rank) [wait: assembled] -> the frame to page-locked host memory

    so the all-gather of frame i is on the wire while frames i+1 .. i+depth-1 render, and the head of each frame's
    persistent kernel fills the SMs that the tail of the previous frame (its last few long rays) leaves idle.
    Every rank enqueues the collectives in frame order on its one comm stream, so they match up across ranks.
    `pre_frame` (optional callable(stream)) is enqueued on the render stream before each frame (bench.py: the L2 flush).
    PyTorch is plumbing only (streams, events, the NCCL communicator).
    """

    def __init__(self, gpu, width, height, rank=0, world=1, depth=2, group=None, to_host=False, pre_frame=None):
        import torch
        self.torch = torch
        self.gpu, self.W, self.H, self.rank, self.P, self.D = gpu, width, height, rank, world, depth
        self.group, self.pre_frame = group, pre_frame
        self.rps = rows_per_shard(height, world)
        dev = torch.device("cuda", torch.cuda.current_device())
        i32 = dict(dtype=torch.int32, device=dev)
        self.render_streams = [torch.cuda.Stream() for _ in range(depth)]
        # high priority: when a render CTA retires, the pending all-gather / de-interleave CTAs get its SM before the next
        # frame's persistent CTAs do (those would hold it for a whole frame)
        self.comm = torch.cuda.Stream(priority=-1)
        self.rendered = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.assembled = [torch.cuda.Event() for _ in range(depth)]
        self.copy = torch.cuda.Stream()
        self.full = [torch.zeros((height, width), **i32) for _ in range(depth)]
        if world > 1:
            self.shard = [torch.zeros((self.rps, width), **i32) for _ in range(depth)]
            self.gathered = [torch.zeros((world * self.rps, width), **i32) for _ in range(depth)]
        self.host = [torch.zeros((height, width), dtype=torch.int32).pin_memory() for _ in range(depth)] if to_host else None
        self.submitted = 0
        self.launches = 0
        torch.cuda.synchronize()                        # the buffers above were zero-filled on the current stream

    def frame_rows(self, frame):
        """Stamp this rank's row-cyclic shard into a b200r_frame."""
        frame.row_first, frame.row_step = (self.rank, self.P) if self.P > 1 else (0, 1)
        return frame

    def submit(self, frame):
        """Enqueue frame (already stamped by frame_rows); returns its slot. Never blocks the host."""
        torch = self.torch
        i = self.submitted
        d = i % self.D
        rs = self.render_streams[d]
        if i >= self.D:
            rs.wait_event(self.free[d])                 # slot d's previous frame: gathered (+ copied out)
        if self.pre_frame is not None:
            self.pre_frame(rs)
        target = self.shard[d] if self.P > 1 else self.full[d]
        self.gpu.render_device_slot(frame, target.data_ptr(), rs.cuda_stream, d)
        self.launches += self.gpu.last_launches()
        self.rendered[d].record(rs)
        if self.P > 1 or self.host is not None:
            self.comm.wait_event(self.rendered[d])
            with torch.cuda.stream(self.comm):
                if self.P > 1:
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(self.gathered[d], self.shard[d], group=self.group)
                    self.gpu.deinterleave_device(self.gathered[d].data_ptr(), self.full[d].data_ptr(), self.W, self.H,
                                                 self.P, self.comm.cuda_stream)
                    self.launches += 1
            if self.host is not None:                   # copy-out on its own stream: the next frame's all-gather does not wait for it
                self.assembled[d].record(self.comm)
                self.copy.wait_event(self.assembled[d])
                with torch.cuda.stream(self.copy):
                    self.host[d].copy_(self.full[d], non_blocking=True)
                self.free[d].record(self.copy)
            else:
                self.free[d].record(self.comm)
        else:
            self.free[d].record(rs)
        self.submitted += 1
        return d

    def join(self, stream):
        """Make `stream` wait for everything submitted so far (device-side; the host does not block)."""
        for s in self.render_streams:
            stream.wait_stream(s)
        stream.wait_stream(self.comm)
        stream.wait_stream(self.copy)

    def start_after(self, stream):
        """Nothing submitted from now on starts before `stream`'s current tail (e.g. the start event of a timed region)."""
        for s in self.render_streams:
            s.wait_stream(stream)
        self.comm.wait_stream(stream)
        self.copy.wait_stream(stream)

    def drain(self):
        for s in self.render_streams:
            s.synchronize()
        self.comm.synchronize()
        self.copy.synchronize()


def init_nccl(local_rank):
    """torch.distributed over NCCL with the communicator's internal stream at high priority (see FramePipeline.comm)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    except (AttributeError, TypeError):
        dist.init_process_group("nccl", device_id=dev)
    return dist
