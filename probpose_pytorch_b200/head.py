"""Post-convolution tail of ``ProbMapHead.forward_heatmap`` (head.py:526-532):
``clamp(x / temperature, 0, 1)`` for ``normalize=None`` and
``clamp(Sparsemax(dim=-1)(x / temperature) * normalize, 0, 1)`` over the H*W pixels of each heatmap otherwise
(head.py:237-245; the projection is the PyPI package ``sparsemax==0.1.9`` in the reference).

The head's deconvolution / convolution / BatchNorm stacks are dense contractions and stay in
cuDNN / cuBLAS; what this module replaces is the per-pixel tail:

* ``heatmap_tail(x, t)``              -- one read + one write, differentiable (backward is one kernel);
* ``HeatmapTail``                     -- the same as an ``nn.Module``;
* ``Sparsemax(dim=-1)``                -- the normalisation layer on its own, same constructor as the package's;
* ``patch_probmap_head(head)``        -- makes an existing reference ``ProbMapHead`` use the fused tail,
  leaving its layers and its 5-tuple output contract (head.py:487-511) untouched;
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


def _sparsemax_forward(x: Tensor, temperature: float, normalize: float, want_aux: bool):
    """One launch: (..., n) logits -> projected, scaled, clamped values of the same shape (+ (rows, 2) aux)."""
    _lib.require_cuda()
    if not x.is_cuda:
        raise RuntimeError("the Sparsemax tail needs a CUDA tensor; there is no CPU fallback")
    src = x.detach().contiguous()
    n = src.shape[-1]
    rows = src.numel() // n if n else 0
    dst = torch.empty_like(src)
    aux = torch.empty((rows, 2), dtype=torch.float32, device=src.device) if want_aux else None
    with torch.cuda.device(src.device):
        rc = _lib.lib().pp_sparsemax_tail(_lib.ptr(src), _lib.ptr(dst), _lib.ptr(aux), _lib.dtype_code(src.dtype), rows, n,
                                          float(temperature), float(normalize), _lib.stream_ptr(src.device))
    _lib.check(rc, "pp_sparsemax_tail")
    return dst, aux


class _SparsemaxTailFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, temperature: float, normalize: float):
        y, aux = _sparsemax_forward(x, temperature, normalize, True)
        ctx.save_for_backward(x, aux)
        ctx.temperature, ctx.normalize = temperature, normalize
        return y

    @staticmethod
    def backward(ctx, grad_y: Tensor):
        x, aux = ctx.saved_tensors
        xs = x.detach().contiguous()
        gy = grad_y.detach().to(xs.dtype).contiguous()
        gx = torch.empty_like(xs)
        n = xs.shape[-1]
        with torch.cuda.device(xs.device):
            rc = _lib.lib().pp_sparsemax_tail_backward(
                _lib.ptr(xs), _lib.ptr(gy), _lib.ptr(aux), _lib.ptr(gx), _lib.dtype_code(xs.dtype), xs.numel() // n, n,
                float(ctx.temperature), float(ctx.normalize), _lib.stream_ptr(xs.device))
        _lib.check(rc, "pp_sparsemax_tail_backward")
        return gx, None, None


def _sparsemax_rows(x: Tensor, temperature: float, normalize: float) -> Tensor:
    if torch.is_grad_enabled() and x.requires_grad:
        return _SparsemaxTailFunction.apply(x, float(temperature), float(normalize))
    return _sparsemax_forward(x, temperature, normalize, False)[0]


def heatmap_tail(x: Tensor, temperature: float = 0.5, out: Tensor | None = None,
                 normalize: float | None = None) -> Tensor:
    """Tail of ``forward_heatmap`` on a CUDA tensor (float32 or bfloat16).

    ``normalize=None``: ``clamp(x / temperature, 0, 1)``, any shape.  Otherwise ``x`` is ``(B, K, H, W)`` and
    each heatmap is projected onto the simplex first: ``clamp(sparsemax(x / temperature) * normalize, 0, 1)``.
    Differentiable unless ``out`` is given."""
    if normalize is not None:
        if out is not None:
            raise ValueError("heatmap_tail: `out` is only supported with normalize=None")
        if x.ndim < 2:
            raise ValueError("heatmap_tail: the Sparsemax tail needs (..., H, W) heatmaps")
        flat = x.reshape(*x.shape[:-2], x.shape[-2] * x.shape[-1])
        return _sparsemax_rows(flat, temperature, normalize).reshape(x.shape)
    if out is not None or not (torch.is_grad_enabled() and x.requires_grad):
        return _tail_forward(x, temperature, out)
    return _TailFunction.apply(x, float(temperature))


class HeatmapTail(nn.Module):
    """``nn.Module`` form of :func:`heatmap_tail` (``temperature`` / ``normalize`` are plain floats, as in
    head.py:107,237)."""

    def __init__(self, temperature: float = 0.5, normalize: float | None = None):
        super().__init__()
        self.temperature = temperature
        self.normalize = normalize

    def forward(self, x: Tensor) -> Tensor:
        return heatmap_tail(x, self.temperature, normalize=self.normalize)


class Sparsemax(nn.Module):
    """Drop-in for ``sparsemax.Sparsemax`` (sparsemax==0.1.9) on CUDA tensors: Euclidean projection of the
    logits along ``dim`` onto the probability simplex.  The reference only uses ``dim=-1`` (head.py:241)."""

    def __init__(self, dim: int = -1):
        super().__init__()
        self.dim = dim

    def forward(self, x: Tensor) -> Tensor:
        dim = self.dim if self.dim >= 0 else x.ndim + self.dim
        if dim != x.ndim - 1:
            return _sparsemax_rows(x.movedim(dim, -1).contiguous(), 1.0, 1.0).movedim(-1, dim)
        return _sparsemax_rows(x, 1.0, 1.0)


def patch_probmap_head(head: nn.Module) -> nn.Module:
    """Give a reference ``ProbMapHead`` instance the fused tail.

    Only ``forward_heatmap`` changes: the layer stacks run as before, then the tail
    ``reshape -> / temperature -> [Sparsemax -> * normalize] -> clamp(0, 1) -> reshape`` (head.py:526-532)
    is one kernel.
    """
    for name in ("deconv_layers", "conv_layers", "final_layer", "temperature"):
        if not hasattr(head, name):
            raise TypeError(f"patch_probmap_head: {type(head).__name__} has no attribute {name!r}")

    def forward_heatmap(self, x: Tensor) -> Tensor:
        x = self.final_layer(self.conv_layers(self.deconv_layers(x)))
        normalize = getattr(self, "normalize", None)
        return heatmap_tail(x, float(self.temperature), normalize=None if normalize is None else float(normalize))

    head.forward_heatmap = types.MethodType(forward_heatmap, head)
    return head
