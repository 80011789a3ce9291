"""CPU oracle of the Sparsemax-normalised head tail  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference calls ``sparsemax.Sparsemax(dim=-1)`` from the PyPI package
``sparsemax==0.1.9`` (requirements.txt:4; call sites head.py:11, 241, 528).  The package is neither under
/root/reference nor installed here, the reference has no test or fixture for it, and a head with
``normalize != None`` cannot even be constructed in this container.  What follows restates the *published*
algorithm of that package (Martins & Astudillo 2016, Algorithm 1, as implemented by the package's
``SparsemaxFunction``): shift by the maximum, sort descending, ``k* = max{k : 1 + k z_(k) > cumsum_k}``,
``tau = (sum_{k <= k*} z_(k) - 1) / k*``, ``p = max(0, z - tau)``; backward
``g_z = [p != 0] * (g_p - sum(g_p [p != 0]) / |{p != 0}|)``.  It is checked against the defining properties of
the projection (non-negative, sums to one, KKT conditions) and a float64 evaluation in
tests/test_oracle_golden.py, not against reference outputs.

``head_tail_sparsemax`` is the tail of ``ProbMapHead.forward_heatmap`` (head.py:526-532) around it.
"""

from __future__ import annotations

import numpy as np
import torch


class _SparsemaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z: torch.Tensor) -> torch.Tensor:      # projection along the last axis
        z = z - z.max(-1, keepdim=True).values
        zs = z.sort(-1, descending=True).values
        k = torch.arange(1, z.shape[-1] + 1, dtype=z.dtype).expand_as(z)
        bound = 1 + k * zs
        is_gt = (bound > zs.cumsum(-1)).to(z.dtype)
        k_star = (is_gt * k).max(-1, keepdim=True).values
        tau = ((is_gt * zs).sum(-1, keepdim=True) - 1) / k_star
        out = torch.clamp_min(z - tau, 0)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor) -> torch.Tensor:
        (out,) = ctx.saved_tensors
        nz = out != 0
        mean = (g * nz).sum(-1, keepdim=True) / nz.sum(-1, keepdim=True)
        return nz * (g - mean)


def sparsemax(z: torch.Tensor) -> torch.Tensor:
    """Sparsemax along the last axis (differentiable, torch CPU, dtype of ``z``)."""
    return _SparsemaxFunction.apply(z)


def sparsemax_f64(z: np.ndarray) -> np.ndarray:
    """Exact-arithmetic yardstick: the same projection in float64 NumPy."""
    z = np.asarray(z, dtype=np.float64)
    z = z - z.max(-1, keepdims=True)
    zs = -np.sort(-z, axis=-1)
    k = np.arange(1, z.shape[-1] + 1, dtype=np.float64)
    cs = np.cumsum(zs, axis=-1)
    k_star = ((1 + k * zs) > cs).sum(-1, keepdims=True)
    tau = (np.take_along_axis(cs, k_star - 1, -1) - 1) / k_star
    return np.maximum(z - tau, 0)


def head_tail_sparsemax(x: torch.Tensor, temperature: float = 0.5, normalize: float = 1.0) -> torch.Tensor:
    """``clamp(Sparsemax(dim=-1)(x.reshape(B, C, H*W) / temperature) * normalize, 0, 1)`` reshaped back
    (head.py:526-532 with ``normalize`` set); differentiable."""
    B, C, H, W = x.shape
    y = sparsemax(x.reshape(B, C, H * W) / temperature)
    y = y * normalize
    return torch.clamp(y, 0, 1).reshape(B, C, H, W)
