"""Multi-GPU plumbing for the decode path: frames shard one-per-rank, the only exchange is a gather of the decoded
fountain chunk records to rank 0 (torch.distributed; NCCL on GPUs, gloo in the CPU tests)."""
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
