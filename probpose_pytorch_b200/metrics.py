"""Validation metrics of the reference on the device.

Same names, arguments and return structures as the reference:

* ``pose_pck_accuracy`` / ``keypoint_pck_accuracy`` (loss.py:767-866, over heatmap.py:55-111);
* ``get_pose_accuracy`` / ``get_binary_accuracy`` / ``get_mae`` -- the ``ProbPoseLoss`` methods of
  loss.py:642-712 as free functions (they do not use ``self``).

The reference moves the (N, K, H, W) stacks to the host and takes two NumPy argmax passes per call; here
both passes are ``pp_heatmap_maximum`` launches and the (N, K) arithmetic is one small kernel.  The
``*_device`` variants return tensors without synchronising; the reference-named functions return NumPy /
Python scalars like the reference does (one device -> host read of K + 2 numbers).
"""

from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .heatmap import heatmap_maximum_device

_THRESHOLDS = np.arange(0.1, 1.0, 0.05)      # loss.py:685


def _cuda(x, dtype=None, device=None) -> Tensor:
    t = torch.as_tensor(x)
    if not t.is_cuda:
        _lib.require_cuda()
        t = t.to(device if device is not None else "cuda")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def keypoint_pck_accuracy_device(pred, gt, mask, thr, norm_factor, return_distances: bool = False):
    """Device form of :func:`keypoint_pck_accuracy`: ``(acc (K,) float64, avg_acc () float64, cnt () int32)``
    tensors (+ the (K, N) float32 distances)."""
    pred = _cuda(pred, torch.float32)
    dev = pred.device
    gt = _cuda(gt, torch.float32, dev)
    N, K, D = pred.shape
    if D != 2 or gt.shape != pred.shape:
        raise ValueError(f"pred {tuple(pred.shape)} / gt {tuple(gt.shape)} must both be (N, K, 2)")
    mask = _cuda(mask, None, dev).reshape(N, K).to(torch.bool).to(torch.uint8).contiguous()
    norm = torch.as_tensor(norm_factor)
    # NumPy promotion of (float32 coordinates) / norm_factor: float32 stays float32, everything else is float64
    norm = _cuda(norm, torch.float32 if norm.dtype == torch.float32 else torch.float64, dev).reshape(N, 2)
    acc = torch.empty((K,), dtype=torch.float64, device=dev)
    avg = torch.empty((), dtype=torch.float64, device=dev)
    cnt = torch.empty((), dtype=torch.int32, device=dev)
    dist = torch.empty((K, N), dtype=torch.float32, device=dev) if return_distances else None
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_pck_accuracy(_lib.ptr(pred), _lib.ptr(gt), _lib.ptr(mask), _lib.ptr(norm),
                                        _lib.dtype_code(norm.dtype), N, K, float(thr), _lib.ptr(acc), _lib.ptr(avg),
                                        _lib.ptr(cnt), _lib.ptr(dist), _lib.stream_ptr(dev))
    _lib.check(rc, "pp_pck_accuracy")
    return (acc, avg, cnt, dist) if return_distances else (acc, avg, cnt)


def keypoint_pck_accuracy(pred, gt, mask, thr, norm_factor) -> tuple:
    """PCK of coordinates (loss.py:825-866): ``(acc np.ndarray[K], avg_acc float, cnt int)``; ``acc[k] = -1``
    for keypoints without a valid instance; unlike the reference the caller's ``norm_factor`` is not modified."""
    acc, avg, cnt = keypoint_pck_accuracy_device(pred, gt, mask, thr, norm_factor)
    packed = torch.cat([acc, avg.reshape(1), cnt.to(torch.float64).reshape(1)]).cpu().numpy()
    return packed[:-2].copy(), float(packed[-2]), int(packed[-1])


def pose_pck_accuracy_device(output, target, mask, thr: float = 0.05, normalize=None):
    output = _cuda(output)
    target = _cuda(target, None, output.device)
    N, K, H, W = output.shape
    if normalize is None:
        normalize = np.tile(np.array([[H, W]]), (N, 1))      # loss.py:814 (applied to (x, y) as is)
    pred = heatmap_maximum_device(output)[0]
    gt = heatmap_maximum_device(target)[0]
    return keypoint_pck_accuracy_device(pred, gt, mask, thr, normalize)


def pose_pck_accuracy(output, target, mask, thr: float = 0.05, normalize=None, method: str = "argmax") -> tuple:
    """PCK from heatmaps (loss.py:767-822).  ``output`` / ``target``: (N, K, H, W) NumPy arrays or tensors.
    Only ``method="argmax"`` is callable in the reference (its ``"expected"`` branch omits the required
    ``sigmas`` argument, loss.py:820-821) and only that is provided."""
    method = method.lower()
    if method not in ["argmax", "expected"]:
        raise ValueError(f"Invalid method: {method}")
    if method == "expected":
        raise TypeError("get_heatmap_expected_value() missing 1 required positional argument: 'sigmas'")
    if output.shape[1] == 0:
        return None, 0, 0
    acc, avg, cnt = pose_pck_accuracy_device(output, target, mask, thr, normalize)
    packed = torch.cat([acc, avg.reshape(1), cnt.to(torch.float64).reshape(1)]).cpu().numpy()
    return packed[:-2].copy(), float(packed[-2]), int(packed[-1])


def get_pose_accuracy(dt, gt, mask) -> Tensor:
    """``ProbPoseLoss.get_pose_accuracy`` (loss.py:642-651): average argmax-PCK@0.05 as a tensor on ``gt.device``."""
    _, avg, _ = pose_pck_accuracy_device(dt.detach(), gt.detach(), mask)
    return avg


def _select(dt, gt, mask):
    dt = _cuda(dt.detach() if isinstance(dt, Tensor) else dt, torch.float32)
    dev = dt.device
    gt = _cuda(gt.detach() if isinstance(gt, Tensor) else gt, torch.float32, dev)
    assert dt.shape == gt.shape
    mask = _cuda(mask, None, dev).to(torch.bool).expand(dt.shape).to(torch.uint8).contiguous()
    return dt, gt, mask, dev


def get_binary_accuracy(dt, gt, mask, force_balanced: bool = False):
    """``ProbPoseLoss.get_binary_accuracy`` (loss.py:653-697): ``(best_acc, best_threshold)`` float32 tensors on
    the device.  ``force_balanced=True`` draws a random subset with ``np.random.shuffle`` in the reference
    (never used by its training loop); not provided."""
    if force_balanced:
        raise NotImplementedError("get_binary_accuracy(force_balanced=True) is a host-side random subsample; not provided")
    dt, gt, mask, dev = _select(dt, gt, mask)
    thr = torch.from_numpy(_THRESHOLDS).to(dev)
    out = torch.empty((2,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_binary_accuracy(_lib.ptr(dt), _lib.ptr(gt), _lib.ptr(mask), dt.numel(), _lib.ptr(thr),
                                           thr.numel(), _lib.ptr(out), None, _lib.stream_ptr(dev))
    _lib.check(rc, "pp_binary_accuracy")
    return out[0], out[1]


def get_mae(dt, gt, mask) -> Tensor:
    """``ProbPoseLoss.get_mae`` (loss.py:699-712): mean absolute error over the mask, float32 tensor on the device."""
    dt, gt, mask, dev = _select(dt, gt, mask)
    out = torch.empty((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_masked_mae(_lib.ptr(dt), _lib.ptr(gt), _lib.ptr(mask), dt.numel(), _lib.ptr(out),
                                      _lib.stream_ptr(dev))
    _lib.check(rc, "pp_masked_mae")
    return out
