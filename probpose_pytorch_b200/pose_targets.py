"""Device-resident replacements for the two host round trips inside ``ProbPoseLoss.forward``
(SURVEY.md section 8 f-1):

* ``ProbPoseLoss._oks_from_heatmaps`` (loss.py:550-640): DARK-decode the target and the predicted
  heatmaps, then per-keypoint OKS (``compute_oks(use_area=False, per_kpt=True)``, loss.py:715-764);
* ``ProbPoseLoss._error_from_heatmaps`` (loss.py:512-548): Euclidean distance between the two decodes.

The reference copies both (B, K, H, W) stacks to the host and decodes sample by sample with NumPy /
OpenCV on every training step; here both stacks are decoded by ``pp_decode_argmax_dark`` and the
(B, K) arithmetic is finished by ``pp_pose_targets``, without leaving the GPU or synchronising.
"""

from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import _lib


def _decode_pair(codec, gt_heatmaps: Tensor, dt_heatmaps: Tensor):
    probmap = getattr(codec, "probmap", codec)
    gt = probmap.decode_device(gt_heatmaps.detach())["keypoints"]
    dt = probmap.decode_device(dt_heatmaps.detach())["keypoints"]
    return probmap, gt, dt


def _launch(gt: Tensor, dt: Tensor, weight, sigmas, heatmap_size, want_oks: bool, want_err: bool):
    B, K, _ = gt.shape
    dev = gt.device
    oks = torch.empty((B, K), dtype=torch.float32, device=dev) if want_oks else None
    okw = torch.empty((B,), dtype=torch.float32, device=dev) if want_oks else None
    err = torch.empty((B, K), dtype=torch.float64, device=dev) if want_err else None
    w = sg = None
    if want_oks:
        w = torch.as_tensor(weight).to(device=dev, dtype=torch.float32).reshape(B, K).contiguous()
        sg = torch.as_tensor(np.asarray(sigmas, dtype=np.float64)[:K]).to(dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_pose_targets(_lib.ptr(gt), _lib.ptr(dt), _lib.ptr(w), _lib.ptr(sg), B, K,
                                        float(heatmap_size[0]), float(heatmap_size[1]), _lib.ptr(oks), _lib.ptr(okw),
                                        _lib.ptr(err), _lib.stream_ptr(dev))
    _lib.check(rc, "pp_pose_targets")
    return oks, okw, err


def oks_from_heatmaps(codec, gt_heatmaps: Tensor, dt_heatmaps: Tensor, weight: Tensor, heatmap_size=(48, 64)):
    """``ProbPoseLoss._oks_from_heatmaps``: returns ``(target_oks (B, K) float32, oks_weights (B,) float32)``
    on the device of the heatmaps.  ``codec`` is a ``Codec(ArgMaxProbMap(...))`` (or the probmap itself)."""
    probmap, gt, dt = _decode_pair(codec, gt_heatmaps, dt_heatmaps)
    oks, okw, _ = _launch(gt, dt, weight, probmap.sigmas, heatmap_size, True, False)
    return oks, okw


def error_from_heatmaps(codec, gt_heatmaps: Tensor, dt_heatmaps: Tensor) -> Tensor:
    """``ProbPoseLoss._error_from_heatmaps``: (B, K) float64 Euclidean distance between the decoded target and
    predicted keypoints (input-image pixels)."""
    _, gt, dt = _decode_pair(codec, gt_heatmaps, dt_heatmaps)
    return _launch(gt, dt, None, None, (1, 1), False, True)[2]
