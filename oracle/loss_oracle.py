"""CPU oracle: OKS heatmap loss forward (+ autograd backward) and a closed-form gradient.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates
``OKSHeatmapLoss.forward`` / ``_get_mask`` (loss.py:55-143, 145-191) with
torch-CPU float32 ops, so the backward is whatever autograd derives from the
same graph the reference builds.  ``oks_heatmap_loss_grad_closed_form`` is an
independent NumPy float64 statement of d(mean per-pixel loss)/d(output) used to
cross-check both autograd and the CUDA kernel.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# cross-correlation stencils of loss.py:106-107
_SOBEL_X = [[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]]
_SOBEL_Y = [[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]]


def _combined_mask(target, target_weights, mask, skip_empty_channel):
    """loss.py:145-191: product of the spatial mask, the (B,K) or (B,K,H,W)
    target weights and, optionally, the non-empty-channel indicator."""
    m = mask
    if target_weights is not None:
        assert target_weights.ndim in (2, 4)
        assert target_weights.shape == target.shape[: target_weights.ndim]
        w = target_weights.view(target_weights.shape + (1,) * (target.ndim - target_weights.ndim))
        m = w if m is None else m * w
    if skip_empty_channel:
        ne = (target != 0).flatten(2).any(dim=2)
        ne = ne.view(ne.shape + (1, 1))
        m = ne if m is None else m * ne
    return m


def oks_heatmap_loss(output, target, target_weights=None, mask=None, per_pixel=False,
                     per_keypoint=False, *, skip_empty_channel=False, smoothing_weight=0.2,
                     gaussian_weight=0.0, loss_weight=1.0, oks_type="minus"):
    """Forward of ``OKSHeatmapLoss`` (loss.py:55-143); differentiable w.r.t. ``output``.

    per_pixel -> (B,K,H,W) ``w_s*grad + w_o*oks + w_g*mse`` (loss.py:122-127);
    per_keypoint -> (B,K) ``w_o*sum(oks) + w_s*max(grad) + w_g*mean(mse)``
    (loss.py:128-134); default -> mean of that (loss.py:135-141).
    """
    assert target.max() <= 1 and target.min() >= 0, "target should be normalized"  # loss.py:85-86
    B, K, H, W = output.shape
    m = _combined_mask(target, target_weights, mask, skip_empty_channel)

    minus = output * (1 - target)
    plus = (1 - output) * target
    oks = {"minus": minus, "plus": plus, "both": (minus + plus) / 2}[oks_type]
    mse = (output - target) ** 2

    flat = output.reshape(B * K, 1, H, W)
    kx = torch.tensor(_SOBEL_X, dtype=torch.float32).view(1, 1, 3, 3)
    ky = torch.tensor(_SOBEL_Y, dtype=torch.float32).view(1, 1, 3, 3)
    gx = F.conv2d(flat, kx, padding=1)
    gy = F.conv2d(flat, ky, padding=1)
    grad = (gx ** 2 + gy ** 2).reshape(B, K, H, W)

    if m is not None:
        oks, mse, grad = oks * m, mse * m, grad * m

    w_s, w_g = smoothing_weight, gaussian_weight
    w_o = 1 - w_s - w_g
    if per_pixel:
        loss = w_s * grad + w_o * oks + w_g * mse
    else:
        peak = grad.reshape(B, K, H * W).max(dim=-1)[0]
        loss = w_o * oks.sum(dim=(2, 3)) + w_s * peak + w_g * mse.mean(dim=(2, 3))
        if not per_keypoint:
            loss = loss.mean()
    return loss * loss_weight


def _corr3(a: np.ndarray, k: np.ndarray) -> np.ndarray:
    """3x3 cross-correlation with zero 'same' padding over the last two axes."""
    H, W = a.shape[-2:]
    p = np.pad(a, [(0, 0)] * (a.ndim - 2) + [(1, 1), (1, 1)])
    out = np.zeros_like(a)
    for i in range(3):
        for j in range(3):
            if k[i][j] != 0.0:
                out += k[i][j] * p[..., i:i + H, j:j + W]
    return out


def oks_heatmap_loss_grad_closed_form(output, target, weight_map=None, *, smoothing_weight=0.2,
                                      gaussian_weight=0.0, loss_weight=1.0, oks_type="minus"):
    """d/d(output) of ``mean(per_pixel loss)`` in float64:

        [ w_o * d(oks)/d(out) * m + w_g * 2 (out - tgt) m
          - w_s * ( Sx * (2 gx m) + Sy * (2 gy m) ) ] * loss_weight / (B K H W)

    where ``*`` is zero-padded cross-correlation; the minus sign comes from the
    adjoint of an antisymmetric-under-flip stencil (flip(Sx) = -Sx, flip(Sy) = -Sy).
    ``weight_map`` broadcasts to (B, K, H, W).
    """
    o = np.asarray(output, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    m = np.ones_like(o) if weight_map is None else np.broadcast_to(np.asarray(weight_map, np.float64), o.shape)
    w_s, w_g = smoothing_weight, gaussian_weight
    w_o = 1 - w_s - w_g
    d_oks = {"minus": 1 - t, "plus": -t, "both": (1 - 2 * t) / 2}[oks_type]
    gx = _corr3(o, _SOBEL_X)
    gy = _corr3(o, _SOBEL_Y)
    smooth = -(_corr3(2 * gx * m, _SOBEL_X) + _corr3(2 * gy * m, _SOBEL_Y))
    g = w_o * d_oks * m + w_g * 2 * (o - t) * m + w_s * smooth
    return g * (loss_weight / o.size)
