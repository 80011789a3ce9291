"""Post-convolution tail of ``ProbMapHead.forward_heatmap`` (head.py:526-532, ``normalize=None``):
``clamp(x / temperature, 0, 1)``.

The head's deconvolution / convolution / BatchNorm stacks are dense contractions and stay in
cuDNN / cuBLAS; what this module replaces is the per-pixel tail:

* ``heatmap_tail(x, t)``              -- one read + one write, differentiable (backward is one kernel);
* ``HeatmapTail``                     -- the same as an ``nn.Module``;
* ``patch_probmap_head(head)``        -- makes an existing reference ``ProbMapHead`` (``normalize=None``)
  use the fused tail, leaving its layers and its 5-tuple output contract (head.py:487-511) untouched;
* ``decode_device(..., temperature=t)`` on the codecs fuses the tail into the decoder's load, so at
  inference the clamped maps never touch HBM.
"""

from __future__ import annotations

import types

import torch
from torch import Tensor, nn

from . import _lib


def _tail_forward(x: Tensor, temperature: float, out: Tensor | None = None) -> Tensor:
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


class _TailFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, temperature: float):
        ctx.save_for_backward(x)
        ctx.temperature = temperature
        return _tail_forward(x, temperature)

    @staticmethod
    def backward(ctx, grad_y: Tensor):
        (x,) = ctx.saved_tensors
        xs = x.detach().contiguous()
        gy = grad_y.detach().to(xs.dtype).contiguous()
        gx = torch.empty_like(xs)
        with torch.cuda.device(xs.device):
            rc = _lib.lib().pp_heatmap_tail_backward(_lib.ptr(xs), _lib.ptr(gy), _lib.ptr(gx), _lib.dtype_code(xs.dtype),
                                                     xs.numel(), float(ctx.temperature), _lib.stream_ptr(xs.device))
        _lib.check(rc, "pp_heatmap_tail_backward")
        return gx, None


def heatmap_tail(x: Tensor, temperature: float = 0.5, out: Tensor | None = None) -> Tensor:
    """``clamp(x / temperature, 0, 1)`` of a CUDA tensor (float32 or bfloat16).  Differentiable unless
    ``out`` is given."""
    if out is not None or not (torch.is_grad_enabled() and x.requires_grad):
        return _tail_forward(x, temperature, out)
    return _TailFunction.apply(x, float(temperature))


class HeatmapTail(nn.Module):
    """``nn.Module`` form of :func:`heatmap_tail` (``temperature`` is a plain float, as in head.py:107)."""

    def __init__(self, temperature: float = 0.5):
        super().__init__()
        self.temperature = temperature

    def forward(self, x: Tensor) -> Tensor:
        return heatmap_tail(x, self.temperature)


def patch_probmap_head(head: nn.Module) -> nn.Module:
    """Give a reference ``ProbMapHead`` instance the fused tail.

    Only ``forward_heatmap`` changes: the layer stacks run as before, then the tail
    ``reshape -> / temperature -> clamp(0, 1) -> reshape`` (head.py:526-532) is one kernel.  Heads
    built with ``normalize != None`` use Sparsemax and are left alone.
    """
    if getattr(head, "normalize", None) is not None:
        raise NotImplementedError("patch_probmap_head: Sparsemax-normalised heads are not covered (normalize=None only)")
    for name in ("deconv_layers", "conv_layers", "final_layer", "temperature"):
        if not hasattr(head, name):
            raise TypeError(f"patch_probmap_head: {type(head).__name__} has no attribute {name!r}")

    def forward_heatmap(self, x: Tensor) -> Tensor:
        x = self.final_layer(self.conv_layers(self.deconv_layers(x)))
        return heatmap_tail(x, float(self.temperature))

    head.forward_heatmap = types.MethodType(forward_heatmap, head)
    return head
