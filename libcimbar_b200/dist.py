"""Multi-GPU plumbing for the decode path (one process per GPU): frames shard f % world, the only exchange is the hand-over
of the decoded fountain chunk records to rank 0 -- the role of concurrent_fountain_decoder_sink in the reference
(src/lib/fountain/concurrent_fountain_decoder_sink.h:58-84).

The exchange itself lives behind the C ABI (include/cb200.h, csrc/gather.cu); torch.distributed is only the host channel
that carries the 64-byte IPC handle / the 128-byte NCCL id between the processes at start-up:

  RecordExchange(ctx, "window")  every rank's RS / chunk-mask kernels store straight into a window in rank 0's HBM over NVLink
                                 (CUDA IPC peer mapping); epochs are published / awaited with system-scope flags on the device
  RecordExchange(ctx, "nccl")    cb200_gather_chunks: ncclSend / ncclRecv on a side stream, overlapping the next decode
  gather_records(...)            plain torch.distributed.gather of host or device tensors (gloo in the CPU tests)
"""
import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world):
    """frame f belongs to rank f % world: indices of this rank's frames"""
    return list(range(rank, n_frames, world))


def gather_records(chunks, mask, dst=0):
    """chunks: (n, chunks_per_frame*chunk_size) uint8, mask: (n,) int32 -- same n on every rank.
    Returns (list_of_chunks, list_of_masks) on dst, (None, None) elsewhere."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    gc = [torch.empty_like(chunks) for _ in range(world)] if rank == dst else None
    gm = [torch.empty_like(mask) for _ in range(world)] if rank == dst else None
    dist.gather(chunks, gc, dst=dst)
    dist.gather(mask, gm, dst=dst)
    return gc, gm


class RecordExchange:
    """Double-buffered hand-over of `n` chunk records per rank and step to rank 0.

    step s (1, 2, 3, ...) uses buffer s & 1:
        d_chunks, d_mask = ex.begin(s)        # where this rank's decode of step s must write (device addresses)
        ctx.decode_chunks_dev(frames, n, d_chunks, d_mask, ...)
        ex.end(s)                             # publish / start the exchange of step s
        ... on rank 0, whenever the records of step s are needed (typically one step later, so that the exchange overlaps
        the next decode):  ex.collect(s) -> (d_all_chunks, d_all_masks) device addresses, rank-major, valid on the context's
        stream; ex.release(s) when rank 0 is done with them."""

    def __init__(self, ctx, kind, n, rank=None, world=None):
        import libcimbar_b200 as cb
        self.ctx, self.kind, self.n = ctx, kind, n
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        info = ctx.info
        self.rec_bytes = info.data_bytes
        self.send = None
        if kind in ("window", "window-direct"):
            box = [ctx.gather_root_create(self.world) if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            if self.rank != 0:
                ctx.gather_peer_open(self.world, self.rank, box[0])
                if kind == "window":       # push form: the decode writes locally, a copy engine moves the records to rank 0
                    dev = torch.device("cuda", torch.cuda.current_device())
                    self.send = [(torch.empty((n, self.rec_bytes), dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.int32, device=dev))
                                 for _ in range(2)]
        elif kind == "nccl":
            box = [cb.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0], self.world, self.rank)
            dev = torch.device("cuda", torch.cuda.current_device())
            self.send = [(torch.empty((n, self.rec_bytes), dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.int32, device=dev))
                         for _ in range(2)]
            self.recv = [(torch.empty((self.world, n, self.rec_bytes), dtype=torch.uint8, device=dev),
                          torch.empty((self.world, n), dtype=torch.int32, device=dev)) if self.rank == 0 else (None, None)
                         for _ in range(2)]
        else:
            raise ValueError("kind must be 'window', 'window-direct' or 'nccl'")

    def begin(self, step):
        b = step & 1
        if self.kind == "window" and self.rank != 0:
            if step > 2:
                self.ctx.gather_chunks_wait(b)                 # the transfer of step - 2 out of these local buffers is done
            c, m = self.send[b]
            return c.data_ptr(), m.data_ptr()
        if self.kind in ("window", "window-direct"):
            if step > 2:
                self.ctx.gather_acquire(b, step - 2)           # rank 0 has let go of the records of step - 2
            return self.ctx.gather_slot(b)
        if step > 2:
            self.ctx.gather_chunks_wait(b)                     # the send buffers of step - 2 are free again (side stream done)
        c, m = self.send[b]
        return c.data_ptr(), m.data_ptr()

    def end(self, step):
        b = step & 1
        if self.kind == "window" and self.rank != 0:
            c, m = self.send[b]
            self.ctx.gather_push(b, c.data_ptr(), m.data_ptr(), self.n, step, step - 2 if step > 2 else 0)
        elif self.kind in ("window", "window-direct"):
            self.ctx.gather_publish(b, step)
        else:
            c, m = self.send[b]
            ac, am = self.recv[b]
            self.ctx.gather_chunks(self.world, self.rank, b, c.data_ptr(), m.data_ptr(), self.n,
                                   ac.data_ptr() if ac is not None else None, am.data_ptr() if am is not None else None)

    def collect(self, step):
        """rank 0: make the context's stream wait for every rank's records of `step`; returns their device addresses"""
        b = step & 1
        if self.kind in ("window", "window-direct"):
            self.ctx.gather_wait(b, step)
            return self.ctx.gather_slot(b, 0)                  # rank r's records: + r * slot stride (ctx.gather_slot(b, r))
        self.ctx.gather_chunks_wait(b)
        ac, am = self.recv[b]
        return ac.data_ptr(), am.data_ptr()

    def release(self, step):
        if self.kind in ("window", "window-direct") and self.rank == 0:
            self.ctx.gather_release(step & 1, step)
