"""Device-side parts for the reference's own ``ProbPoseLoss`` (loss.py:342-712).

``ProbPoseLoss`` is the training-step *caller* of the hot path, not part of it, so this package does not
re-implement it.  :func:`patch_probpose_loss` takes an instance of the reference's class and swaps the
pieces that sit on the hot path for their CUDA versions; the reference's ``forward`` and its four
scalar-head losses keep running unchanged:

=====================================  ================================================================
reference member (loss.py)             replaced by
=====================================  ================================================================
``keypoint_loss_module`` (:348-352)    :class:`FusedOKSHeatmapLoss`: ``module(out, tgt, w, per_pixel=True).mean()``
                                       (:428-431) becomes ONE fused forward+backward kernel
``_oks_from_heatmaps`` (:550-640)      two DARK decodes + ``pp_pose_targets`` on the device, no host copy
``_error_from_heatmaps`` (:512-548)    same; the (B, K) result is returned as a NumPy array because the
                                       caller wraps it with ``torch.from_numpy`` (:383-384)
``get_pose_accuracy`` (:642-651),      ``metrics`` kernels
``get_binary_accuracy`` (:653-697),
``get_mae`` (:699-712)
=====================================  ================================================================

:func:`ground_truth_from_keypoints` builds the ground-truth dictionary the reference's ``forward`` reads
(dataset.py:130-135) from (B, K, 2) keypoints on the device, so a loader can ship keypoints instead of
(B, K, H, W) target planes.
"""

from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import metrics
from .codec import ArgMaxProbMap, Codec
from .loss import OKSHeatmapLoss
from .pose_targets import error_from_heatmaps, oks_from_heatmaps

__all__ = ["FusedOKSHeatmapLoss", "LazyPerPixelLoss", "patch_probpose_loss", "ground_truth_from_keypoints"]


class LazyPerPixelLoss:
    """The un-reduced per-pixel loss of ``OKSHeatmapLoss(..., per_pixel=True)``, not materialised yet.

    ``.mean()`` -- the only thing ``ProbPoseLoss.forward`` does with it (loss.py:431) -- runs the fused
    forward+backward kernel and never writes the (B, K, H, W) map.  Anything else (indexing, arithmetic,
    ``.sum()``, passing it to torch functions) first materialises the real tensor with the per-pixel kernel.
    """

    def __init__(self, module: OKSHeatmapLoss, output, target, target_weights, mask):
        self._module, self._args = module, (output, target, target_weights, mask)
        self._tensor = None

    def mean(self, *args, **kwargs):
        if args or kwargs or self._tensor is not None:
            return self.materialize().mean(*args, **kwargs)
        return self._module.forward_mean(*self._args)

    def materialize(self) -> Tensor:
        if self._tensor is None:
            self._tensor = OKSHeatmapLoss.forward(self._module, *self._args, per_pixel=True)
        return self._tensor

    @property
    def shape(self):
        return self._args[0].shape

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.materialize() if isinstance(a, LazyPerPixelLoss) else a
        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in (kwargs or {}).items()})

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __len__(self):
        return self.shape[0]


for _op in ("add", "radd", "sub", "rsub", "mul", "rmul", "truediv", "rtruediv", "neg", "pow"):
    def _delegate(self, *a, _name=f"__{_op}__"):
        return getattr(self.materialize(), _name)(*a)
    setattr(LazyPerPixelLoss, f"__{_op}__", _delegate)


class FusedOKSHeatmapLoss(OKSHeatmapLoss):
    """``OKSHeatmapLoss`` whose ``per_pixel=True`` result is lazy (see :class:`LazyPerPixelLoss`)."""

    def forward(self, output, target, target_weights=None, mask=None, per_pixel=False, per_keypoint=False):
        if per_pixel:
            return LazyPerPixelLoss(self, output, target, target_weights, mask)
        return super().forward(output, target, target_weights, mask, per_pixel, per_keypoint)


def _device_codec(codec) -> Codec:
    """This package's ``Codec(ArgMaxProbMap)`` with the parameters of the reference codec object."""
    pm = getattr(codec, "probmap", codec)
    if isinstance(pm, ArgMaxProbMap):
        return codec if isinstance(codec, Codec) else Codec(pm)
    return Codec(ArgMaxProbMap(pm.input_size, pm.heatmap_size, pm.sigmas, sigma=getattr(pm, "sigma", -1),
                               blur_kernel_size=getattr(pm, "blur_kernel_size", 11)))


def _balanced_subset(gt, mask) -> np.ndarray | None:
    """The random balanced subsample of loss.py:666-682, drawn like the reference does (NumPy's global generator,
    positives shuffled first) on the (B, K) booleans; returns the selection as a flat boolean array over all entries."""
    m = np.asarray(torch.as_tensor(mask).detach().cpu().numpy(), dtype=bool).reshape(-1)
    g = torch.as_tensor(gt).detach().cpu().numpy().reshape(-1)[m].astype(bool)
    n_pos = int(g.sum())
    num = min(n_pos, int(g.size) - n_pos)
    if num == 0:
        return None
    pos, neg = np.flatnonzero(g), np.flatnonzero(~g)
    np.random.shuffle(pos)
    np.random.shuffle(neg)
    chosen = np.zeros(g.size, dtype=bool)
    chosen[pos[:num]] = True
    chosen[neg[:num]] = True
    sel = np.zeros(m.size, dtype=bool)
    sel[np.flatnonzero(m)[chosen]] = True
    return sel


def patch_probpose_loss(loss_module):
    """Swap the hot-path members of a reference ``ProbPoseLoss`` instance for their CUDA versions, in place.

    ``loss_module`` is the reference's object (anything with its member names works); its ``codec`` may be the
    reference's ``Codec(ArgMaxProbMap(...))`` -- an equivalent device codec is built from its parameters.
    Returns ``loss_module``.
    """
    old = loss_module.keypoint_loss_module
    loss_module.keypoint_loss_module = FusedOKSHeatmapLoss(
        use_target_weight=getattr(old, "use_target_weight", True), skip_empty_channel=getattr(old, "skip_empty_channel", False),
        smoothing_weight=getattr(old, "smoothing_weight", 0.05), gaussian_weight=getattr(old, "gaussian_weight", 0.0),
        loss_weight=getattr(old, "loss_weight", 1.0), oks_type=getattr(old, "oks_type", "minus"))
    dev_codec = _device_codec(loss_module.codec)

    def _oks(gt_heatmaps, dt_heatmaps, weight, heatmap_size=(48, 64)):
        return oks_from_heatmaps(dev_codec, gt_heatmaps, dt_heatmaps, weight, heatmap_size=heatmap_size)

    def _error(gt_heatmaps, dt_heatmaps):
        return error_from_heatmaps(dev_codec, gt_heatmaps, dt_heatmaps).cpu().numpy()

    def _binary_accuracy(dt, gt, mask, force_balanced=False):
        if force_balanced:
            sel = _balanced_subset(gt, mask)
            if sel is None:
                zero = torch.tensor([0.0], device=gt.device)
                return zero, zero.clone()
            return metrics.get_binary_accuracy(torch.as_tensor(dt).reshape(-1), torch.as_tensor(gt).reshape(-1).to(torch.float32),
                                               torch.from_numpy(sel))
        return metrics.get_binary_accuracy(dt, gt, mask)

    loss_module._oks_from_heatmaps = _oks
    loss_module._error_from_heatmaps = _error
    loss_module.get_pose_accuracy = metrics.get_pose_accuracy
    loss_module.get_binary_accuracy = _binary_accuracy
    loss_module.get_mae = lambda dt, gt, mask: metrics.get_mae(dt, torch.as_tensor(gt).to(torch.float32), mask)
    loss_module.device_codec = dev_codec
    return loss_module


def ground_truth_from_keypoints(codec, keypoints, keypoints_visible, keypoints_visibility, *,
                                dtype: torch.dtype = torch.float32, device: torch.device | None = None) -> dict:
    """The ground-truth dictionary of one batch (what the reference's collate yields from ``Codec.encode``,
    dataset.py:116-135) built on the device from (B, K, 2) input-space keypoints: ``heatmaps`` (B, K, H, W),
    ``in_image``, ``keypoints_visible`` (= annotated) and ``keypoints_visibility``.  ``codec``: this package's
    ``Codec(ArgMaxProbMap)`` or the object :func:`patch_probpose_loss` returned."""
    codec = getattr(codec, "device_codec", codec)
    enc = codec.probmap.encode_batch(torch.as_tensor(keypoints), torch.as_tensor(keypoints_visible).to(torch.float32),
                                     dtype=dtype, device=device)
    dev = enc["heatmaps"].device
    return dict(heatmaps=enc["heatmaps"], in_image=enc["in_image"], keypoints_visible=enc["annotated"],
                keypoints_visibility=torch.as_tensor(keypoints_visibility).to(dev), keypoint_weights=enc["keypoint_weights"])
