"""``ProbPoseLoss`` -- the training-step caller of the hot path (loss.py:342-510) -- on the device.

Same constructor, ``forward`` arguments, returned dictionaries and loss modules as the reference.  What
changes is where the work happens:

* heatmap loss: ``OKSHeatmapLoss(per_pixel=True).mean()`` is one fused forward+backward kernel
  (``OKSHeatmapLoss.forward_mean``) instead of ~20 elementwise / convolution launches and a (B, K, H, W)
  temporary per launch;
* OKS / error targets (``_oks_from_heatmaps`` / ``_error_from_heatmaps``, loss.py:512-640): two DARK
  decodes + a (B, K) kernel on the device instead of a device->host copy of both heatmap stacks and
  2 B per-sample NumPy / OpenCV decodes per step (``pose_targets``);
* accuracy read-outs (``compute_acc=True``, loss.py:463-508): ``metrics`` kernels;
* the ground-truth dict (dataset.py:130-135) may carry ``keypoints`` (B, K, 2) instead of ``heatmaps``: the
  targets are then encoded on the device (``encode_batch``), so the DataLoader workers no longer encode and the
  per-step host->device copy of the (B, K, H, W) target stack (loss.py:376) disappears.

The four scalar heads' losses (BCE / MSE / smooth-L1-of-logs on (B, K) values, loss.py:194-339) are a few
torch element-wise calls on 4 K numbers: plumbing, kept in torch on the device.
"""

from __future__ import annotations

from functools import partial
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import metrics
from .loss import OKSHeatmapLoss
from .pose_targets import error_from_heatmaps, oks_from_heatmaps


class BCELoss(nn.Module):
    """Binary cross entropy on probabilities (``use_sigmoid=True``) or logits; loss.py:194-260."""

    def __init__(self, use_target_weight=False, loss_weight=1.0, reduction="mean", use_sigmoid=False):
        super().__init__()
        assert reduction in ("mean", "sum", "none"), (
            f"the argument `reduction` should be either 'mean', 'sum' or 'none', but got {reduction}")
        self.reduction = reduction
        self.use_sigmoid = use_sigmoid
        self.criterion = partial(F.binary_cross_entropy if use_sigmoid else F.binary_cross_entropy_with_logits,
                                 reduction="none")
        self.use_target_weight = use_target_weight
        self.loss_weight = loss_weight

    def forward(self, output, target, target_weight=None):
        loss = self.criterion(output, target)
        if self.use_target_weight:
            assert target_weight is not None
            if target_weight.dim() == 1:
                target_weight = target_weight[:, None]
            loss = loss * target_weight
        if self.reduction == "sum":
            loss = loss.sum()
        elif self.reduction == "mean":
            loss = loss.mean()
        return loss * self.loss_weight


class MSELoss(nn.Module):
    """``mse_loss(output * w, target * w)``; loss.py:263-292."""

    def __init__(self, use_target_weight=False, loss_weight=1.0):
        super().__init__()
        self.criterion = F.mse_loss
        self.use_target_weight = use_target_weight
        self.loss_weight = loss_weight

    def forward(self, output, target, target_weight=None):
        if self.use_target_weight:
            assert target_weight is not None
            loss = self.criterion(output * target_weight, target * target_weight)
        else:
            loss = self.criterion(output, target)
        return loss * self.loss_weight


class L1LogLoss(nn.Module):
    """Smooth-L1 between ``log(1 + output)`` and ``log(1 + target)``; loss.py:295-339."""

    def __init__(self, use_target_weight=False, loss_weight=1.0):
        super().__init__()
        self.criterion = F.smooth_l1_loss
        self.use_target_weight = use_target_weight
        self.loss_weight = loss_weight

    def forward(self, output, target, target_weight=None):
        output = torch.log(1 + output)
        target = torch.log(1 + target)
        if self.use_target_weight:
            assert target_weight is not None
            assert output.ndim >= target_weight.ndim
            for _ in range(output.ndim - target_weight.ndim):
                target_weight = target_weight.unsqueeze(-1)
            loss = self.criterion(output * target_weight, target * target_weight)
        else:
            loss = self.criterion(output, target)
        return loss * self.loss_weight


class ProbPoseLoss(nn.Module):
    """Drop-in for the reference's ``ProbPoseLoss`` (loss.py:342-510).  ``codec`` is a
    ``Codec(ArgMaxProbMap(...))`` of this package; predictions must live on a CUDA device."""

    def __init__(self, codec, freeze_error: bool = True):
        super().__init__()
        self.codec = codec
        self.keypoint_loss_module = OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus")
        self.probability_loss_module = BCELoss(use_target_weight=False, use_sigmoid=True)
        self.visibility_loss_module = BCELoss(use_target_weight=False, use_sigmoid=True)
        self.oks_loss_module = MSELoss(use_target_weight=True)
        self.error_loss_module = L1LogLoss(use_target_weight=True)
        self.freeze_error = freeze_error
        self.freeze_oks = False

    def forward(self, gt, pred, keypoint_weights: Tensor | None = None, learn_heatmaps_from_zeros: bool = False,
                compute_acc: bool = False):
        dt_heatmaps, dt_probs, dt_vis, dt_oks, dt_errs = pred
        device = dt_heatmaps.device
        B, C, H, W = dt_heatmaps.shape
        if keypoint_weights is None:
            keypoint_weights = torch.ones((B, C), device=device, dtype=dt_heatmaps.dtype)

        def dev(x, dtype):
            return torch.as_tensor(x).to(device, dtype=dtype)

        if "heatmaps" not in gt:
            # keypoint front end: the loader ships (B, K, 2) input-space keypoints and the
            # target planes are encoded here, on the device -- B*K*3 numbers cross PCIe instead of B*K*H*W
            enc = self.codec.probmap.encode_batch(torch.as_tensor(gt["keypoints"]).reshape(B, C, -1),
                                                  torch.as_tensor(gt["keypoints_visible"]).reshape(B, C).to(torch.float32),
                                                  dtype=dt_heatmaps.dtype, device=device)
            gt = dict(gt, heatmaps=enc["heatmaps"], in_image=gt.get("in_image", enc["in_image"]))
        gt_heatmaps = dev(gt["heatmaps"], dt_heatmaps.dtype).view((B, C, H, W))
        gt_probs = dev(gt["in_image"], torch.int64).view((B, C))
        gt_annotated = dev(gt["keypoints_visible"], torch.int64).view((B, C))
        gt_vis = dev(gt["keypoints_visibility"], torch.int64).view((B, C))
        dt_heatmaps = dt_heatmaps.view((B, C, H, W))

        # targets of the error / OKS heads, from the two stacks of heatmaps, without leaving the device
        if self.freeze_error:
            gt_errs = torch.zeros((B, C), device=device, dtype=dt_errs.dtype)
        else:
            gt_errs = self._error_from_heatmaps(gt_heatmaps, dt_heatmaps).to(dt_errs.dtype)
        if self.freeze_oks:
            gt_oks = torch.zeros((B, C), device=device, dtype=dt_oks.dtype)
        else:
            gt_oks, _ = self._oks_from_heatmaps(gt_heatmaps, dt_heatmaps, gt_probs & gt_annotated, heatmap_size=(W, H))
            gt_oks = gt_oks.to(dt_oks.dtype).view((B, C))

        dt_probs, dt_vis = dt_probs.view((B, C)), dt_vis.view((B, C))
        dt_oks, dt_errs = dt_oks.view((B, C)), dt_errs.view((B, C))
        keypoint_weights = keypoint_weights.view((B, C))
        annotated_in = gt_annotated & (gt_probs > 0.5)

        heatmap_weights = gt_annotated if learn_heatmaps_from_zeros else keypoint_weights
        # == keypoint_loss_module(dt, gt, w, per_pixel=True).mean(), fused forward + backward
        heatmap_loss = self.keypoint_loss_module.forward_mean(dt_heatmaps, gt_heatmaps, heatmap_weights.to(torch.float32))
        probability_loss = self.probability_loss_module(dt_probs, gt_probs.float())

        # loss.py:438-452: weights that balance visible / invisible keypoints.  The visibility module is built
        # with use_target_weight=False, so they do not enter the loss; they are still formed, because an
        # annotated-free batch fails here in the reference (min() of an empty tensor) and must fail here too.
        invisible_in = (gt_vis == 0) & (gt_annotated > 0.5)
        visible_in = (gt_vis > 0) & (gt_annotated > 0.5)
        weighted = annotated_in.clone().to(torch.float64)
        weighted[invisible_in] = (1 / (invisible_in.sum() + 1e-10)).to(weighted.dtype)
        weighted[visible_in] = (1 / (visible_in.sum() + 1e-10)).to(weighted.dtype)
        weighted = (weighted / weighted[weighted > 0].min()).to(dt_vis.dtype)

        visibility_loss = self.visibility_loss_module(dt_vis, gt_vis.float(), weighted)
        oks_loss = self.oks_loss_module(dt_oks, gt_oks, annotated_in)
        error_loss = self.error_loss_module(dt_errs, gt_errs, annotated_in)
        losses = dict(kpt=heatmap_loss, probability=probability_loss, visibility=visibility_loss, oks=oks_loss,
                      error=error_loss)
        if not compute_acc:
            return losses
        acc = {
            "kpt": self.get_pose_accuracy(dt_heatmaps, gt_heatmaps, keypoint_weights > 0.5),
            "probability": self.get_binary_accuracy(dt_probs, gt_probs, gt_annotated > 0.5, force_balanced=True)[0],
            "visibility": self.get_binary_accuracy(dt_vis, gt_vis, annotated_in > 0.5, force_balanced=True)[0],
            "oks": self.get_mae(dt_oks, gt_oks, annotated_in > 0.5),
            "error": self.get_mae(dt_errs, gt_errs, annotated_in > 0.5),
        }
        return losses, acc

    # ---- loss.py:512-640 ------------------------------------------------------------------------------
    def _error_from_heatmaps(self, gt_heatmaps: Tensor, dt_heatmaps: Tensor) -> Tensor:
        """(B, K) float64 distance between the DARK-decoded target and prediction, on the device."""
        return error_from_heatmaps(self.codec, gt_heatmaps, dt_heatmaps)

    def _oks_from_heatmaps(self, gt_heatmaps: Tensor, dt_heatmaps: Tensor, weight: Tensor,
                           heatmap_size: Sequence[int] = (48, 64)):
        return oks_from_heatmaps(self.codec, gt_heatmaps, dt_heatmaps, weight, heatmap_size=heatmap_size)

    # ---- loss.py:642-712 ------------------------------------------------------------------------------
    def get_pose_accuracy(self, dt, gt, mask):
        return metrics.get_pose_accuracy(dt, gt, mask)

    def get_binary_accuracy(self, dt, gt, mask, force_balanced=False):
        """With ``force_balanced`` the reference keeps an equal number of randomly chosen positives and
        negatives (``np.random.shuffle`` on the host, loss.py:666-682).  The same draws are made here from
        NumPy's global generator -- the selection is (B, K) booleans -- and the counting stays on the device."""
        if not force_balanced:
            return metrics.get_binary_accuracy(dt, gt, mask)
        device = gt.device
        m = np.asarray(torch.as_tensor(mask).detach().cpu().numpy(), dtype=bool).reshape(-1)
        g = torch.as_tensor(gt).detach().cpu().numpy().reshape(-1)[m].astype(bool)
        num = min(int(g.sum()), int(len(g) - g.sum()))
        if num == 0:
            return torch.tensor([0.0], device=device), torch.tensor([0.0], device=device)
        pos_idx, neg_idx = np.where(g)[0], np.where(~g)[0]
        np.random.shuffle(pos_idx)
        np.random.shuffle(neg_idx)
        chosen = np.zeros(len(g), dtype=bool)
        chosen[np.concatenate([pos_idx[:num], neg_idx[:num]])] = True
        sel = np.zeros(m.shape, dtype=bool)
        sel[np.where(m)[0][chosen]] = True
        return metrics.get_binary_accuracy(torch.as_tensor(dt).reshape(-1), torch.as_tensor(gt).reshape(-1).to(torch.float32),
                                           torch.from_numpy(sel))

    def get_mae(self, dt, gt, mask):
        return metrics.get_mae(dt, torch.as_tensor(gt).to(torch.float32), mask)
