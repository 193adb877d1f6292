"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous only).

The data path has exactly one exchange step per dense quasi-Newton iteration (all-gather of the
row-block slices of h = H y and u = H' g, done by the CUDA library over NCCL); everything here is
bookkeeping: which rows / problems / samples a rank owns, and how the NCCL id reaches every rank.
"""
import os


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_rows(n, rank, world):
    """Row block [row0, row0 + nrows) of the n x n inverse-Hessian approximation owned by `rank`.
    The CUDA kernels tile 8 rows per CTA, hence the divisibility requirement."""
    if n % (8 * world) != 0:
        raise ValueError("row-sharded H needs n divisible by 8 * world (n=%d, world=%d)" % (n, world))
    nrows = n // world
    return rank * nrows, nrows


def shard_problems(n_problems, rank, world):
    """Contiguous slice [p0, p0 + count) of a batch of independent problems (no collective needed)."""
    base, rem = divmod(n_problems, world)
    count = base + (1 if rank < rem else 0)
    p0 = rank * base + min(rank, rem)
    return p0, count


def shard_indices(n, rank, world, block=2):
    """Contiguous coordinate range [i0, i0 + count) of an n-vector owned by `rank` (index-range sharded GD / PGD / SPG).
    Ranges are multiples of `block` (the objective functor's block size: 2 for extended Rosenbrock) and equal-sized,
    because the all-gather of the per-rank scalars assumes nothing about them but the rank order."""
    if n % (block * world) != 0:
        raise ValueError("index-range sharding needs n divisible by block * world (n=%d, block=%d, world=%d)" % (n, block, world))
    count = n // world
    return rank * count, count


def shard_samples(m, rank, world):
    if m % world != 0:
        raise ValueError("sample count must divide by the number of ranks")
    return rank * (m // world), m // world


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object (the 128-byte ncclUniqueId) from `src` over the default process group."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def make_context(osb, backend_init=True):
    """Create the per-rank CUDA context of the library from torchrun's environment."""
    import torch.distributed as dist
    rank, world, local_rank = env_rank_world()
    if world == 1:
        return osb.Context(local_rank)
    if backend_init and not dist.is_initialized():
        dist.init_process_group("nccl")
    uid = broadcast_bytes(osb.Context.nccl_unique_id() if rank == 0 else None)
    ctx = osb.Context(local_rank, rank, world, uid)
    ctx.connect_peers()
    return ctx
