"""Data-parallel plumbing for the learner (one process per GPU, torch.distributed).

Episodes are independent units: rank r trains on a contiguous slice of the sampled batch and
contributes the UN-normalised gradient of sum((td * mask)^2) plus the five loss sums, packed into one
fp32 buffer; ONE all-reduce(sum) per step, after which every rank applies the identical clip + RMSprop
update (parameters / optimizer state stay replicated - no broadcast).  Replay indices are drawn
once (rank 0's numpy stream) and sliced, so sampling stays identical to the 1-GPU run."""
import torch as th
import torch.distributed as dist


def is_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_slice(batch_size, rank, world):
    """Contiguous episode range [lo, hi) of this rank (the first `batch_size % world` ranks get one more)."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_fields(fields, rank, world):
    """Slice every batch-major field to this rank's episodes (views, no copy)."""
    some = next(iter(fields.values()))
    lo, hi = shard_slice(some.shape[0], rank, world)
    return {k: v[lo:hi] for k, v in fields.items()}


def rank():
    return dist.get_rank() if is_active() else 0


def world_size():
    return dist.get_world_size() if is_active() else 1


def allreduce_step(flat_grad_and_sums):
    """THE exchange of a step: sum `[un-normalised flat gradient | loss sums as (hi, lo) floats]` (one fp32 buffer, see
    pmb_dp_pack in include/pymarl_b200.h) over all ranks, in place, in ONE collective."""
    if not is_active():
        return
    dist.all_reduce(flat_grad_and_sums, op=dist.ReduceOp.SUM)
