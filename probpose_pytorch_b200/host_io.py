"""Host-buffer front end of the hot path: double-buffered H2D / D2H staging around a device step.

The reference keeps targets and predictions in host memory at both ends of the path (NumPy in
``codec.py:214-239``, ``.cpu()`` in ``loss.py:531,578``).  A caller that has to start from host buffers is
bound by the PCIe copy of the prediction tensor, not by the kernels, so the only thing worth doing is to
keep the copy engine busy: the inputs of step i+1 travel while step i computes, and the (small) results of
step i are read back while step i+1 computes.

Plumbing only -- streams, events and pinned buffers from PyTorch; the compute is whatever ``step_fn``
launches (the C-ABI kernels through the codec / loss shims).
"""

from __future__ import annotations

import collections
from typing import Callable, Iterable, Iterator, Sequence, Tuple

import torch
from torch import Tensor


class _Slot:
    __slots__ = ("dev_in", "host_out", "h2d_done", "step_done")

    def __init__(self):
        self.dev_in = None
        self.host_out = None
        self.h2d_done = torch.cuda.Event()
        self.step_done = None


def pipelined_steps(host_batches: Iterable[Sequence[Tensor]], step_fn: Callable[..., Sequence[Tensor]],
                    device: torch.device, depth: int = 2) -> Iterator[Tuple[Tensor, ...]]:
    """Run ``step_fn(*device_inputs) -> device_results`` over batches of *pinned host* tensors.

    Yields, in order, one tuple of host tensors per batch (the results of ``step_fn`` copied back).
    Every batch's inputs are copied host->device and its results device->host; copies of neighbouring
    steps overlap the compute.  ``step_fn`` must not keep references to its inputs (their buffers are
    reused ``depth`` steps later) and must return tensors of the same shapes every step.
    """
    if depth < 2:
        raise ValueError("depth must be >= 2")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pipelined_steps stages into CUDA memory; there is no CPU path")
    cur = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device=device)
    slots = [_Slot() for _ in range(depth)]
    in_flight: collections.deque = collections.deque()

    def issue(slot: _Slot, batch: Sequence[Tensor]) -> None:
        for h in batch:
            if not h.is_pinned():
                raise ValueError("host batches must be pinned (tensor.pin_memory())")
        if slot.dev_in is None:
            slot.dev_in = [torch.empty(h.shape, dtype=h.dtype, device=device) for h in batch]
        with torch.cuda.stream(copy):
            if slot.step_done is not None:
                copy.wait_event(slot.step_done)      # the step that last read this slot's inputs
            for d, h in zip(slot.dev_in, batch):
                d.copy_(h, non_blocking=True)
            slot.h2d_done.record(copy)

    def collect(slot: _Slot) -> Tuple[Tensor, ...]:
        slot.step_done.synchronize()
        return tuple(h.clone() for h in slot.host_out)

    it = iter(host_batches)
    batch = next(it, None)
    i = 0
    if batch is not None:
        issue(slots[0], batch)
    while batch is not None:
        slot = slots[i % depth]
        ahead = next(it, None)
        if ahead is not None:
            nxt = slots[(i + 1) % depth]
            if in_flight and in_flight[0] is nxt:      # its results have not been handed out yet
                yield collect(in_flight.popleft())
            issue(nxt, ahead)
        cur.wait_event(slot.h2d_done)
        outs = step_fn(*slot.dev_in)
        if slot.host_out is None:
            slot.host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        for h, o in zip(slot.host_out, outs):
            h.copy_(o.detach(), non_blocking=True)
        slot.step_done = torch.cuda.Event()
        slot.step_done.record(cur)
        in_flight.append(slot)
        batch = ahead
        i += 1
    while in_flight:
        yield collect(in_flight.popleft())
