"""Batch-sharded RNN-T loss: one process per GPU, lattices partitioned across ranks, ONE scalar
all-reduce per step.

The reference has no distributed code at all (single GPU, single stream).  Lattices are
independent (the batch is blockIdx.z in core.cu:49,377), so the path shards with no data-path
collective: every rank runs the single-GPU op on its rank-local log_probs / labels / lengths and
keeps its rank-local gradient.  The only exchange is the loss value itself: a 1- or 2-element
all-reduce (sum of costs [, sample count]) over NCCL / NVLink -- latency-bound, never bandwidth.

Host logic only; the arithmetic is the CUDA op in ``warp_rnnt_b200.rnnt_loss``.  ``loss_fn`` is
injectable so that the sharding / reduction logic can be tested on CPU ranks (gloo).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, world_size, rank):
    """Contiguous partition of ``n_total`` lattices: ranks [0, n_total % world) get one extra."""
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(log_probs, labels, frames_lengths, labels_lengths, world_size, rank):
    """Slice a (host or device) dense batch down to this rank's lattices."""
    lo, hi = shard_range(log_probs.shape[0], world_size, rank)
    return log_probs[lo:hi], labels[lo:hi], frames_lengths[lo:hi], labels_lengths[lo:hi]


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Each rank's x only influences the (replicated) y once, so for
    the usual "every rank calls backward on the same replicated loss" convention the local gradient
    is the incoming gradient unchanged (no second collective)."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_loss(local_sum, local_count, reduction="mean", group=None):
    """Global 'sum' or 'mean' of per-sample costs from rank-local partial sums.

    local_sum: 0-d tensor (sum of this rank's costs); local_count: python int.  One all-reduce of
    2 elements for 'mean' (sum, count), 1 element for 'sum'."""
    if reduction not in ("sum", "mean"):
        raise ValueError("sharded loss supports reduction 'sum' or 'mean', got %r" % (reduction,))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_sum if reduction == "sum" else local_sum / max(local_count, 1)
    if reduction == "sum":
        return _AllReduceSum.apply(local_sum.reshape(1), group)[0]
    packed = torch.stack([local_sum, local_sum.new_tensor(float(local_count))])
    total = _AllReduceSum.apply(packed, group)
    return total[0] / total[1].detach()


def rnnt_loss_sharded(log_probs, labels, frames_lengths, labels_lengths, average_frames=False,
                      reduction="mean", blank=0, gather=False, fastemit_lambda=0.0, compact=False,
                      group=None, loss_fn=None):
    """Loss over a batch that is sharded across the ranks of ``group``.

    Arguments are this rank's shard (same meaning as ``rnnt_loss``).  Returns the GLOBAL reduced
    loss, identical on every rank; ``.backward()`` leaves d(global loss)/d(local log_probs) in
    ``log_probs.grad`` (gradients never cross GPUs)."""
    if loss_fn is None:
        from . import rnnt_loss as loss_fn
    costs = loss_fn(log_probs, labels, frames_lengths, labels_lengths, average_frames=average_frames,
                    reduction="none", blank=blank, gather=gather, fastemit_lambda=fastemit_lambda,
                    compact=compact)
    return all_reduce_loss(costs.sum(), int(costs.shape[0]), reduction, group)
