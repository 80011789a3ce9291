"""Multi-GPU plumbing of the heatmap path: one process per GPU, batch sharded by image.

Every (b, k) heatmap is independent in encode and decode, and the loss is a sum over
heatmaps, so there is no data-path collective.  The only exchanges are
  * the loss all-reduce (a few floats), and
  * the gather of the decoded keypoint records (B/G x K x 7 values per rank),
both latency-bound over NVLink/NVSwitch.  ``exchange_step_results`` folds them into ONE
``all_gather`` per step (the loss partial rides in the same buffer), which halves the
collective launch latency of small steps.  torch.distributed is used as plumbing (NCCL on
GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def world() -> tuple[int, int]:
    """(world_size, rank); (1, 0) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split of ``batch`` images over ``world_size`` ranks (earlier ranks take the remainder)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, world_size: int | None = None, rank: int | None = None):
    """This rank's contiguous slice (by image) of a global batch tensor / array."""
    ws, rk = world()
    ws = ws if world_size is None else world_size
    rk = rk if rank is None else rank
    lo, hi = shard_bounds(len(t), ws, rk)
    return t[lo:hi]


def all_reduce_loss(loss_sum: Tensor, count: int | Tensor) -> Tensor:
    """Global mean loss from per-rank partial sums: all-reduce(SUM) of (sum, count)."""
    buf = torch.stack([loss_sum.detach().reshape(()).double(),
                       torch.as_tensor(count, dtype=torch.float64, device=loss_sum.device).reshape(())])
    if world()[0] > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return (buf[0] / buf[1]).to(loss_sum.dtype)


def all_gather_keypoints(records: Tensor) -> Tensor:
    """Gather per-rank keypoint records (B_local, K, C) into (sum B_local, K, C), rank order.
    Ranks may hold different B_local (ragged split): shorter shards are padded for the collective."""
    ws, _ = world()
    if ws == 1:
        return records
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    if records.shape[0] < m:
        pad = records.new_zeros((m - records.shape[0],) + tuple(records.shape[1:]))
        records = torch.cat([records, pad])
    out = records.new_empty((ws * m,) + tuple(records.shape[1:]))
    dist.all_gather_into_tensor(out.view(-1), records.contiguous().view(-1))
    if all(s == m for s in sizes):
        return out
    return torch.cat([out[r * m: r * m + s] for r, s in enumerate(sizes)])


class StepExchange:
    """Result of :func:`exchange_step_results`: ``records`` (gathered, rank order) and ``loss`` (global
    mean).  With ``async_op=True`` the collective runs on NCCL's own stream, overlapped with whatever
    the caller enqueues next; ``wait()`` makes the current stream wait for it before the results are read."""

    def __init__(self, gathered: Tensor, n_local: int, rec_shape, loss_dtype, work=None):
        self._gathered, self._n, self._shape, self._dtype, self._work = gathered, n_local, rec_shape, loss_dtype, work

    def wait(self) -> "StepExchange":
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self

    @property
    def records(self) -> Tensor:
        self.wait()
        ws = self._gathered.shape[0]
        return self._gathered[:, :-1].reshape((ws * self._n,) + tuple(self._shape[1:]))

    @property
    def loss(self) -> Tensor:
        self.wait()
        return self._gathered[:, -1].mean().to(self._dtype)

    def __iter__(self):   # (records, loss) = exchange_step_results(...)
        yield self.records
        yield self.loss


def exchange_step_results(records: Tensor, loss_mean_local: Tensor, async_op: bool = False):
    """One collective per step: every rank contributes its (B_local, K, C) records and its local
    mean loss (equal B_local on all ranks); yields the gathered records and the global mean loss.
    Returns a :class:`StepExchange` (unpacks like a ``(records, loss)`` tuple)."""
    ws, _ = world()
    if ws == 1:
        return records, loss_mean_local
    flat = torch.cat([records.reshape(-1), loss_mean_local.detach().reshape(1).to(records.dtype)])
    out = flat.new_empty(ws * flat.numel())
    work = dist.all_gather_into_tensor(out, flat, async_op=async_op)
    return StepExchange(out.view(ws, flat.numel()), records.shape[0], records.shape, loss_mean_local.dtype,
                        work if async_op else None)


class BucketExchange:
    """Result of :func:`exchange_bucket`: ``records`` (S, B_global, K, C) in rank order and ``loss`` (S,) global means
    of the S bucketed steps.  ``wait()`` as for :class:`StepExchange`."""

    def __init__(self, gathered: Tensor, steps: int, rec_shape, loss_dtype, work=None):
        self._gathered, self._steps, self._shape, self._dtype, self._work = gathered, steps, rec_shape, loss_dtype, work

    def wait(self) -> "BucketExchange":
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self

    @property
    def records(self) -> Tensor:
        self.wait()
        ws, S = self._gathered.shape[0], self._steps
        rec = self._gathered[:, :-S].reshape((ws, S) + tuple(self._shape))          # (ws, S, B_local, K, C)
        return rec.transpose(0, 1).reshape((S, ws * self._shape[0]) + tuple(self._shape[1:]))

    @property
    def loss(self) -> Tensor:
        self.wait()
        return self._gathered[:, -self._steps:].mean(dim=0).to(self._dtype)

    def __iter__(self):
        yield self.records
        yield self.loss


def exchange_bucket(records: Sequence[Tensor], losses: Sequence[Tensor], async_op: bool = False):
    """The per-step exchange, bucketed: the records and local mean losses of S consecutive steps travel in ONE
    all-gather.  The exchange is latency bound (a few hundred KB per rank over NVSwitch), so its cost per step
    falls as 1 / S; results arrive at most S steps late, which is what loss logging and evaluation tolerate.
    Equal ``B_local`` on all ranks; with one process the inputs come back stacked."""
    S = len(records)
    assert S >= 1 and len(losses) == S
    ws, _ = world()
    rec = torch.stack([r for r in records])
    loss = torch.stack([l.detach().reshape(()) for l in losses])
    if ws == 1:
        return rec, loss
    flat = torch.cat([rec.reshape(-1), loss.to(rec.dtype)])
    out = flat.new_empty(ws * flat.numel())
    work = dist.all_gather_into_tensor(out, flat, async_op=async_op)
    return BucketExchange(out.view(ws, flat.numel()), S, tuple(records[0].shape), losses[0].dtype, work if async_op else None)
