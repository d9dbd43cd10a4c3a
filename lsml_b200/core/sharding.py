"""Multi-GPU partitioning of the hot path.

Batches of independent items (configs C2/C4, and bench.py's one-volume-per-GPU weak scaling)
shard across ranks as contiguous index ranges with NO data-path collective: every item has its
own level set, band and global features, the regressors are read-only and replicated.  The only
communication is the final reduction of counters (band-voxel-iterations, timings), done with
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of ``n_items`` for ``rank``: the first ``n_items % world``
    ranks get one extra item; ranks beyond ``n_items`` get an empty range."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world (%d, %d)" % (rank, world))
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_counters(total_units, elapsed_ms, device=None):
    """(sum of units over ranks, max of elapsed time over ranks) -- the aggregation bench.py
    uses: whole-job units divided by the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()):
        return int(total_units), float(elapsed_ms)
    dev = device if device is not None else (
        torch.device("cuda", torch.cuda.current_device())
        if dist.get_backend() == "nccl" else torch.device("cpu"))
    units = torch.tensor([int(total_units)], dtype=torch.int64, device=dev)
    ms = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return int(units.item()), float(ms.item())


def segment_batch_sharded(model, imgs, dx=None, n_iters=None, device=None):
    """Evolve this rank's contiguous shard of ``imgs`` (all ranks hold / can read the full
    list).  Returns ``(lo, hi, u_final, mask_final)`` for the shard; no collective is issued."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(len(imgs), rank, world)
    if hi == lo:
        return lo, hi, None, None
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    u, m = model.segment_batch(imgs[lo:hi], dx=dx, n_iters=n_iters, device=dev,
                               iterate_until_validation_max=False)
    return lo, hi, u, m
