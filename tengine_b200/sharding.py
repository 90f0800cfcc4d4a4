"""Host-side logic of the multi-GPU path: the batch dimension is sharded contiguously over ranks, the
packed weight arena is broadcast ONCE from rank 0 at prerun, and there is no collective in the steady state.
torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_images, world, rank):
    """Images [start, start+count) of a batch of n_images belong to `rank`: contiguous slices of dim 0 of the NCHW input,
    the first (n_images % world) ranks take one extra image."""
    base, extra = divmod(int(n_images), int(world))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def broadcast_arena(arena, src=0):
    """One broadcast of the packed weight/bias/scale arena (a 1-D uint8 torch tensor: CUDA for NCCL, CPU for gloo)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a few scalars over all ranks (timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def merge_shards(shards):
    """Inverse of the sharding for outputs: concatenate per-rank result slices along dim 0 in rank order."""
    return np.concatenate(list(shards), axis=0)
