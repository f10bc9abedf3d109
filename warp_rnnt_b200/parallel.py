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


def _local_weighted_sum(log_probs, labels, frames_lengths, labels_lengths, scale, average_frames, blank, gather,
                        fastemit_lambda, compact, loss_fn):
    """scale * sum_n cost_n [/ frames_n] of this rank's lattices.  On the default dense path the weights go INTO the
    kernel launch (the gradient it writes is final and backward touches nothing); otherwise the scale reaches the
    gradient through the upstream gradient of backward (one more dense pass)."""
    if loss_fn is None and not gather and not compact:
        from . import RNNTLoss
        n = int(frames_lengths.shape[0])
        weights = torch.full((n,), float(scale), dtype=torch.float32, device=log_probs.device)
        if average_frames:
            weights = weights / frames_lengths.to(torch.float32)
        return RNNTLoss.apply(log_probs, labels, frames_lengths, labels_lengths, blank, fastemit_lambda, weights, True)
    if loss_fn is None:
        from . import rnnt_loss as loss_fn
    local = loss_fn(log_probs, labels, frames_lengths, labels_lengths, average_frames=average_frames,
                    reduction="sum", blank=blank, gather=gather, fastemit_lambda=fastemit_lambda, compact=compact)
    return local * scale if scale != 1.0 else local


def all_reduce_loss_async(local_sum, group=None):
    """Start the one collective of the sharded path without blocking the compute stream: in-place SUM all-reduce of the
    0-d / 1-element ``local_sum`` on NCCL's own stream.  Returns the work handle (``.wait()`` before reading the value)
    or None when there is nothing to reduce.  The gradients never depend on the reduced VALUE (only on the weights,
    which are known up front), so a training loop calls ``backward()`` on the LOCAL loss and lets this overlap with the
    next step's kernel -- a blocking all-reduce sits on the critical path for ~20-30 us per step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(local_sum, op=dist.ReduceOp.SUM, group=group, async_op=True)


def rnnt_loss_sharded(log_probs, labels, frames_lengths, labels_lengths, average_frames=False,
                      reduction="mean", blank=0, gather=False, fastemit_lambda=0.0, compact=False,
                      group=None, loss_fn=None, global_batch=None):
    """Loss over a batch that is sharded across the ranks of ``group``.

    Arguments are this rank's shard (same meaning as ``rnnt_loss``).  Returns the GLOBAL reduced
    loss, identical on every rank; ``.backward()`` leaves d(global loss)/d(local log_probs) in
    ``log_probs.grad`` (gradients never cross GPUs).

    The local part is ONE call of ``rnnt_loss(reduction='sum')`` -- on the dense path one kernel launch that also
    returns the local sum -- followed by one all-reduce of 1 element ('sum', or 'mean' with ``global_batch`` =
    the number of lattices over all ranks known to the caller) or 2 elements ('mean' otherwise: sum and count)."""
    if reduction not in ("sum", "mean"):
        raise ValueError("sharded loss supports reduction 'sum' or 'mean', got %r" % (reduction,))
    n_local = int(frames_lengths.shape[0])
    if reduction == "mean" and global_batch is not None:
        local = _local_weighted_sum(log_probs, labels, frames_lengths, labels_lengths, 1.0 / float(global_batch),
                                    average_frames, blank, gather, fastemit_lambda, compact, loss_fn)
        return all_reduce_loss(local, n_local, "sum", group)
    local = _local_weighted_sum(log_probs, labels, frames_lengths, labels_lengths, 1.0, average_frames, blank, gather,
                                fastemit_lambda, compact, loss_fn)
    return all_reduce_loss(local, n_local, reduction, group)


def rnnt_loss_microbatches(batches, global_batch, reduction="mean", average_frames=False, blank=0, gather=False,
                           fastemit_lambda=0.0, compact=False, group=None, loss_fn=None):
    """One training step over a rank-local batch that is too large for one call (BASELINE configs[4]: 256 lattices
    of 600 x 150 x 1024 per GPU are 94 GB of log-probs plus 94 GB of gradients), processed as micro-batches.

    ``batches`` yields ``(log_probs, labels, frames_lengths, labels_lengths)`` micro-batches of this rank (lattices are
    independent).  Each one runs ``rnnt_loss(..., reduction='sum')`` and its ``backward()`` at once, so
    only one micro-batch's gradient is being produced at a time; the weight 1/global_batch of reduction='mean' is
    applied to the returned value AND, through the upstream gradient of each backward, to the gradients.  The step ends
    with ONE all-reduce of the scalar.  Returns the global loss (detached; the gradients are already in
    ``log_probs.grad`` of every micro-batch that required grad)."""
    if reduction not in ("sum", "mean"):
        raise ValueError("reduction must be 'sum' or 'mean', got %r" % (reduction,))
    scale = 1.0 / float(global_batch) if reduction == "mean" else 1.0
    total = None
    for log_probs, labels, frames_lengths, labels_lengths in batches:
        loss = _local_weighted_sum(log_probs, labels, frames_lengths, labels_lengths, scale, average_frames, blank,
                                   gather, fastemit_lambda, compact, loss_fn)
        if loss.requires_grad:
            loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    if total is None:
        raise ValueError("no micro-batches")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        total = total.reshape(1).clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        total = total[0]
    return total
