"""Batch sharding for the data-parallel runs (host-side logic only; no device code here).

Only the batch dimension is partitioned: rank r owns a contiguous, EVEN-sized slice of the poses so that
  * the consecutive row pairs (2k, 2k+1) of the pairwise deformation loss (train_leg_torso_lifter.py:250-254) and
  * the pose pairs mixed by split_data_left_right_3d (utils/helpers.py:81-91)
never straddle ranks.  Evaluation keeps per-rank double sums and does one final reduction."""
import torch


def shard_bounds(n_items, rank, world, multiple=2, equal=False):
    """[begin, end) of rank's contiguous slice of n_items; every slice length is a multiple of `multiple` except that
    the last rank also takes the remainder -- unless equal=True, which drops the remainder so that every rank owns
    exactly the same number of items (training: every rank must issue the same number of steps / collectives).
    Raises if n_items cannot give every rank at least `multiple` items."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %r / world %r" % (rank, world))
    per = (n_items // world) // multiple * multiple
    if per < multiple and world > 1:
        raise ValueError("%d items cannot be sharded over %d ranks in multiples of %d" % (n_items, world, multiple))
    begin = rank * per
    end = n_items if (rank == world - 1 and not equal) else begin + per
    return begin, end


def shard_rows(t, rank, world, multiple=2):
    b, e = shard_bounds(t.shape[0], rank, world, multiple)
    return t[b:e]


def reduce_eval_sums(sums, count, group=None):
    """(per-rank double sums [k], per-rank pose count) -> global means [k] (eval_h36m.py:83-97 aggregated over ranks).
    One all-reduce of k + 1 doubles; works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    buf = torch.cat((sums.to(torch.float64).flatten(), torch.tensor([float(count)], dtype=torch.float64, device=sums.device)))
    if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
        torch.distributed.all_reduce(buf, group=group)
    total = buf[-1].item()
    return (buf[:-1] / total).tolist(), int(total)


def average_gradients(flat_grad, world, group=None):
    """SUM all-reduce of a flat gradient buffer followed by 1/world (what the fused Adam kernel applies as grad_scale)."""
    if world > 1:
        torch.distributed.all_reduce(flat_grad, group=group)
    return flat_grad.mul_(1.0 / world)
