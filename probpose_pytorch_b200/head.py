"""Post-convolution tail of ``ProbMapHead.forward_heatmap`` (head.py:526-532, ``normalize=None``):
``clamp(x / temperature, 0, 1)``.  Stand-alone kernel here; the decoders can also fuse it into
their load (``decode_device(..., temperature=...)``) so the clamped maps never touch HBM."""

from __future__ import annotations

import torch
from torch import Tensor

from . import _lib


def heatmap_tail(x: Tensor, temperature: float = 0.5, out: Tensor | None = None) -> Tensor:
    """``clamp(x / temperature, 0, 1)`` of a CUDA tensor (float32 or bfloat16), one read + one write."""
    _lib.require_cuda()
    if not x.is_cuda:
        raise RuntimeError("heatmap_tail needs a CUDA tensor; there is no CPU fallback")
    src = x.detach().contiguous()
    dst = torch.empty_like(src) if out is None else out
    assert dst.is_contiguous() and dst.shape == src.shape and dst.dtype == src.dtype
    with torch.cuda.device(src.device):
        rc = _lib.lib().pp_heatmap_tail(_lib.ptr(src), _lib.ptr(dst), _lib.dtype_code(src.dtype), src.numel(),
                                        float(temperature), _lib.stream_ptr(src.device))
    _lib.check(rc, "pp_heatmap_tail")
    return dst
