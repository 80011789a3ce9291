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


class PeerMailbox:
    """Record / loss exchange over NVLink peer memory, without a collective call.

    Every rank owns a mailbox of ``slots x world`` blocks (a symmetric allocation: the same buffer exists on every
    rank and all of them are mapped into every rank's address space).  A step is published into a slot by two parties,
    in any order and on any streams:

    * ``Codec.decode_device(..., mailbox=mb, slot=s)`` / ``Codec.pack_records`` pack the step's keypoint records and,
      in the same kernel, store them into block ``(s, rank)`` of EVERY rank's mailbox (``pp_pack_records``);
    * the loss: ``OKSHeatmapLoss.forward_mean[_encoded](..., publish=mb.descriptor(s))`` stores it from the loss'
      own finalize kernel, or ``commit(s, loss)`` from a one-thread kernel (``pp_mailbox_commit``).

    Whichever finishes last raises the block's flag on every rank.  ``read(s)`` / ``read_async(s)`` wait for the flags
    of all sources, copy the slot and acknowledge it; they return ``(records (B_global, K, 7), losses (world,))`` --
    what one all-gather of records + loss would have delivered.

    Flow control (default): a producer does not rewrite a slot before every rank has acknowledged the slot's previous
    publication, so a reader never sees a block change under it (no torn records, no mixing of steps) -- every rank
    must then consume every publication (``read``, ``read_async`` or ``skip``) before the slot comes round again.
    ``flow_control=False`` drops the acknowledgements: a consumer that is too late gets an error (the sequence number it
    finds is newer than the one it expected), never silently mixed data.  CUDA-graph capturable on the producer side:
    the sequence numbers live in device memory.

    Plumbing: ``torch.distributed._symmetric_memory`` allocates and maps the buffers (world > 1); with one process an
    ordinary tensor plays every peer.
    """

    def __init__(self, batch_local: int, num_keypoints: int, slots: int, device: torch.device, group=None,
                 flow_control: bool = True):
        from . import _lib
        self.world, self.rank = world()
        self.slots, self.device = int(slots), torch.device(device)
        self.flow_control = bool(flow_control)
        self.shape = (int(batch_local), int(num_keypoints), 7)
        self.n_records = int(batch_local) * int(num_keypoints)
        L = _lib.lib()
        self.block_bytes = int(L.pp_mailbox_block_bytes(self.n_records))
        nbytes = int(L.pp_mailbox_bytes(self.n_records, self.world, self.slots))
        self._handle = None
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.buf.zero_()
            self._handle = symm_mem.rendezvous(self.buf, grp)
            ptrs = [int(p) for p in self._handle.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier()                      # every mailbox is zeroed before anybody publishes
        else:
            self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            ptrs = [self.buf.data_ptr()]
        self._peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self._state = torch.zeros(int(L.pp_mailbox_state_words(self.slots)), dtype=torch.int32, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._published = [0] * self.slots       # host copy of the sequence numbers (launch order)
        # the consumer's private, compact copies of every slot (allocated up front: read_async may be captured in a graph)
        self._st_rec = torch.zeros((self.slots, self.world, self.n_records * 7), dtype=torch.float64, device=self.device)
        self._st_loss = torch.zeros((self.slots, self.world, 1), dtype=torch.float64, device=self.device)
        self._loss_slots = torch.zeros((self.slots, 1), dtype=torch.float32, device=self.device)   # see loss_slot()
        self._pending_reads: list = []
        self._lib = _lib

    def descriptor(self, slot: int):
        assert 0 <= slot < self.slots
        return self._lib.Mailbox(self._peer_ptrs.data_ptr(), self._state.data_ptr(), self.world, self.rank, self.slots,
                                 int(slot), self.block_bytes, int(self.flow_control), 0)

    def commit(self, slot: int, loss: Tensor | None = None) -> None:
        """The loss party of the publication of ``slot`` as a kernel of its own: the local loss joins the records (stored
        by ``pp_pack_records`` into the same slot on any stream); the flags go up when both have finished."""
        if loss is not None:
            loss = loss.detach().reshape(()).to(torch.float32)
        with torch.cuda.device(self.device):
            rc = self._lib.lib().pp_mailbox_commit(self.descriptor(slot), self.n_records, self._lib.ptr(loss),
                                                   self._lib.stream_ptr(self.device))
        self._lib.check(rc, "pp_mailbox_commit")
        self.loss_enqueued(slot)

    def loss_slot(self, slot: int) -> Tensor:
        """A persistent (1,) float32 tensor for the loss of ``slot``: pass it as ``loss_out=`` to
        ``OKSHeatmapLoss.forward_mean[_encoded]`` and publish it one step late with ``commit_deferred(slot)``."""
        return self._loss_slots[slot]

    def commit_deferred(self, slot: int, loss: Tensor | None = None) -> None:
        """The loss party of ``slot`` one step late (``pp_mailbox_commit_deferred``): call it at the START of the next
        step -- typically on a side branch of that step's CUDA graph -- so that the NVLink round trips of the
        publication do not sit at the end of the step that produced the loss.  ``loss`` defaults to ``loss_slot(slot)``.
        Publishes only if the slot's records have been packed and are waiting for their loss; otherwise (very first
        step, or the slot was flushed already) nothing happens, so it is safe to call unconditionally.  Flow control
        only (the host keeps no count of these publications)."""
        assert self.flow_control, "deferred commits need flow_control=True"
        src = self._loss_slots[slot] if loss is None else loss.detach().reshape(1).to(torch.float32)
        with torch.cuda.device(self.device):
            rc = self._lib.lib().pp_mailbox_commit_deferred(self.descriptor(slot), self.n_records, self._lib.ptr(src),
                                                            self._lib.stream_ptr(self.device))
        self._lib.check(rc, "pp_mailbox_commit_deferred")

    def loss_enqueued(self, slot: int) -> None:
        """Book-keeping after the loss party of ``slot`` has been enqueued (``commit`` calls it; call it yourself after
        ``forward_mean(..., publish=descriptor(slot))``): outside CUDA-graph capture that completes one publication."""
        if not torch.cuda.is_current_stream_capturing():
            self.published(slot)      # a captured step publishes when (and as often as) its graph is replayed

    def published(self, slot: int, times: int = 1) -> None:
        """Book-keeping: ``slot`` was published ``times`` more times (a replayed CUDA graph publishes without
        passing through Python -- call this after each replay)."""
        self._published[slot] += times

    def _views(self, buf: Tensor):
        blocks = buf[:self.world * self.block_bytes].view(self.world, self.block_bytes)
        rec = blocks[:, :self.n_records * 56].contiguous().view(torch.float64)
        rec = rec.view((self.world * self.shape[0],) + self.shape[1:])
        loss = blocks[:, self.block_bytes - 16:self.block_bytes - 8].contiguous().view(torch.float64).reshape(self.world)
        return rec, loss

    def _consume(self, slot: int, timeout_us: int, copy: bool):
        L = self._lib.lib()
        st = self._lib.stream_ptr(self.device)
        out = None
        if copy:
            out = (self._st_rec[slot].view((self.world * self.shape[0],) + self.shape[1:]), self._st_loss[slot].view(self.world))
        if self.flow_control:
            # one kernel (pp_mailbox_consume), sequence numbers on the device: nothing host-side is baked into the launch,
            # so the consumer side can sit in the same CUDA graph as the step that publishes
            with torch.cuda.device(self.device):
                rc = L.pp_mailbox_consume(self.descriptor(slot), self.n_records,
                                          self._lib.ptr(self._st_rec[slot]) if copy and self.n_records else None,
                                          self._lib.ptr(self._st_loss[slot]) if copy else None,
                                          int(timeout_us), self._lib.ptr(self._status), st)
            self._lib.check(rc, "pp_mailbox_consume")
        else:
            expected = self._published[slot]
            assert expected > 0, "nothing published into this slot yet"
            with torch.cuda.device(self.device):
                rc = L.pp_mailbox_wait(self._lib.ptr(self.buf), self.world, int(slot), self.n_records, expected & 0xFFFFFFFF,
                                       int(timeout_us), self._lib.ptr(self._status), st)
            self._lib.check(rc, "pp_mailbox_wait")
            if copy:
                bb = self.block_bytes
                blocks = self.buf[slot * self.world * bb:(slot + 1) * self.world * bb].view(self.world, bb)
                if self.n_records:
                    self._st_rec[slot].copy_(blocks[:, :self.n_records * 56].view(torch.float64))
                self._st_loss[slot].copy_(blocks[:, bb - 16:bb - 8].view(torch.float64))
        self._pending_reads.append(slot)
        return out

    def read_async(self, slot: int, timeout_us: int = 2_000_000):
        """Enqueue on the current stream, without a host synchronisation, the consumer side for ``slot``: wait on the
        device until every source rank has published it (with flow control: the oldest publication this rank has not
        consumed yet, nothing if there is none; otherwise the latest), copy the slot's blocks into a private per-slot
        buffer, acknowledge.  CUDA-graph capturable (flow control): the step's graph can consume an earlier step's slot.
        Returns that buffer's ``(records, losses)`` views (valid once the stream has run, until the slot is read again);
        ``check_async()`` reports time-outs / overwritten blocks of everything enqueued so far."""
        return self._consume(slot, timeout_us, True)

    def skip(self, slot: int, timeout_us: int = 2_000_000) -> None:
        """Consume a publication of ``slot`` without copying it (wait + acknowledge)."""
        self._consume(slot, timeout_us, False)

    def check_async(self) -> None:
        """Synchronise and raise if a consumer wait timed out / found a newer sequence number, or a producer of this
        rank ran out of patience waiting for an acknowledgement."""
        code = int(self._status.item())
        prod = int(self._state[5 * self.slots].item())
        pending, self._pending_reads = self._pending_reads, []
        self._status.zero_()
        self._state[5 * self.slots].zero_()
        if code > 0:
            raise RuntimeError(f"PeerMailbox: rank {code - 1} did not publish in time (consumed: {pending})")
        if code < 0:
            raise RuntimeError(f"PeerMailbox: the block of rank {-code - 1} was overwritten before it was read "
                               f"(consumed: {pending}); read every slot within `slots` steps or use flow_control=True")
        if prod:
            raise RuntimeError(f"PeerMailbox: rank {prod - 1} did not acknowledge a slot before it came round again; with "
                               "flow control every rank must consume every publication")

    def read(self, slot: int, timeout_us: int = 2_000_000):
        """(records (world * B_local, K, 7) float64 in rank order, losses (world,) float64) of ``slot`` as published by
        every rank (see ``read_async`` for which publication); synchronises."""
        rec, loss = self.read_async(slot, timeout_us)
        self.check_async()
        return rec.clone(), loss.clone()

    def peek(self, slot: int):
        """The slot's current contents without waiting, acknowledging or checking (debugging / tests)."""
        torch.cuda.synchronize(self.device)
        rec, loss = self._views(self.buf[slot * self.world * self.block_bytes:(slot + 1) * self.world * self.block_bytes])
        return rec.clone(), loss.clone()
