"""CPU oracle: data flow of the reference's ``ProbPoseLoss.forward`` (loss.py:360-510)  --  TEST INFRASTRUCTURE ONLY.

``/root/reference`` does not exist on the GPU box, so the GPU tests cannot instantiate the reference's class to
hand it to ``probpose_pytorch_b200.patch_probpose_loss``.  This module supplies a stand-in with the same member
names (:class:`ProbPoseLossLayout`) and a functional restatement of the forward's data flow
(:func:`training_losses`) that calls those members exactly where the reference does.  It is pinned against the
reference itself: ``tests/golden/probpose_loss.npz`` holds the reference's outputs (written by
``oracle/make_golden_probpose_loss.py``, which imports the reference), ``tests/test_oracle_golden.py`` replays them
through this restatement on the CPU, and -- in the build container, where the reference is importable --
``tests/test_reference_patch.py`` patches the reference's own instance.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import metrics_oracle
from .loss_oracle import oks_heatmap_loss
from .targets_oracle import error_from_heatmaps, oks_from_heatmaps

HEAD_NAMES = ("kpt", "probability", "visibility", "oks", "error")


class _CpuHeatmapLoss:
    """``OKSHeatmapLoss(smoothing_weight=0.05, oks_type="minus")`` (loss.py:348-352) through the loss oracle."""

    use_target_weight, skip_empty_channel = True, False
    smoothing_weight, gaussian_weight, loss_weight, oks_type = 0.05, 0.0, 1.0, "minus"

    def __call__(self, output, target, target_weights=None, mask=None, per_pixel=False, per_keypoint=False):
        return oks_heatmap_loss(output, target, target_weights, mask, per_pixel=per_pixel, per_keypoint=per_keypoint,
                                smoothing_weight=self.smoothing_weight, oks_type=self.oks_type)


class ProbPoseLossLayout:
    """Stand-in carrying the member names of the reference's ``ProbPoseLoss`` (loss.py:342-358, 512-712) with CPU
    oracle implementations; ``patch_probpose_loss`` replaces the hot-path ones the same way it does on the real
    object.  ``codec`` needs ``probmap.input_size / heatmap_size / sigmas`` only."""

    def __init__(self, codec, freeze_error: bool = True):
        self.codec = codec
        self.keypoint_loss_module = _CpuHeatmapLoss()
        self.freeze_error, self.freeze_oks = freeze_error, False

    def _pm(self):
        return getattr(self.codec, "probmap", self.codec)

    def _oks_from_heatmaps(self, gt_heatmaps, dt_heatmaps, weight, heatmap_size=(48, 64)):
        pm = self._pm()
        o, w = oks_from_heatmaps(gt_heatmaps.detach().numpy(), dt_heatmaps.detach().numpy(), weight.numpy(), pm.sigmas,
                                 pm.input_size, pm.heatmap_size, area_size=heatmap_size, backend="cv2")
        return torch.from_numpy(o), torch.from_numpy(w)

    def _error_from_heatmaps(self, gt_heatmaps, dt_heatmaps):
        pm = self._pm()
        return error_from_heatmaps(gt_heatmaps.detach().numpy(), dt_heatmaps.detach().numpy(), pm.input_size,
                                   pm.heatmap_size, backend="cv2")

    def get_pose_accuracy(self, dt, gt, mask):
        _, avg, _ = metrics_oracle.pose_pck_accuracy(dt.detach().numpy(), gt.detach().numpy(), np.asarray(mask))
        return torch.tensor(avg)

    def get_binary_accuracy(self, dt, gt, mask, force_balanced=False):
        d, g, m = (np.asarray(t.detach()) for t in (dt, gt, mask))
        d, g = d[m], g[m].astype(bool)
        if force_balanced:      # loss.py:666-682: equal numbers of randomly kept positives and negatives
            num = min(int(g.sum()), int(g.size - g.sum()))
            if num == 0:
                return torch.tensor([0.0]), torch.tensor([0.0])
            pos, neg = np.flatnonzero(g), np.flatnonzero(~g)
            np.random.shuffle(pos)
            np.random.shuffle(neg)
            keep = np.concatenate([pos[:num], neg[:num]])
            d, g = d[keep], g[keep]
        acc, thr = metrics_oracle.binary_accuracy(d, g, np.ones(g.shape, dtype=bool))
        return torch.tensor(acc).float(), torch.tensor(thr).float()

    def get_mae(self, dt, gt, mask):
        return torch.tensor(metrics_oracle.masked_mae(np.asarray(dt.detach()), np.asarray(gt.detach()), np.asarray(mask)))


def _scalar_head_losses(dt_probs, dt_vis, dt_oks, dt_errs, gt_probs, gt_vis, gt_oks, gt_errs, use):
    """The four (B, K)-sized losses as the reference configures them (loss.py:353-356): BCE on probabilities with no
    weighting for presence and visibility (``BCELoss(use_target_weight=False, use_sigmoid=True)``, :194-260),
    MSE of the weighted values for OKS (``MSELoss(use_target_weight=True)``, :263-292) and smooth-L1 between the
    weighted ``log(1 + .)`` for the error head (``L1LogLoss(use_target_weight=True)``, :295-339)."""
    use = use.to(dt_oks.dtype) if use.dtype != dt_oks.dtype else use
    return {
        "probability": F.binary_cross_entropy(dt_probs, gt_probs.float()),
        "visibility": F.binary_cross_entropy(dt_vis, gt_vis.float()),
        "oks": F.mse_loss(dt_oks * use, gt_oks * use),
        "error": F.smooth_l1_loss(torch.log(1 + dt_errs) * use, torch.log(1 + gt_errs) * use),
    }


def training_losses(parts, gt: dict, pred, keypoint_weights=None, learn_heatmaps_from_zeros=False, compute_acc=False):
    """Data flow of ``ProbPoseLoss.forward`` (loss.py:360-510) around the members of ``parts``.  ``gt`` holds
    ``heatmaps``, ``in_image``, ``keypoints_visible``, ``keypoints_visibility``; ``pred`` is the model's 5-tuple."""
    hm, probs, vis, oks, errs = pred
    dev, (B, K, H, W) = hm.device, hm.shape
    flat = lambda t: t.reshape(B, K)
    w_kpt = flat(keypoint_weights) if keypoint_weights is not None else torch.ones((B, K), device=dev, dtype=hm.dtype)

    tgt = gt["heatmaps"].to(dev, dtype=hm.dtype).reshape(B, K, H, W)                               # :375-376
    present = flat(gt["in_image"].to(dev, dtype=torch.int64))                                     # :377
    annotated = flat(gt["keypoints_visible"].to(dev, dtype=torch.int64))                          # :378
    visible = flat(gt["keypoints_visibility"].to(dev, dtype=torch.int64))                         # :379

    if parts.freeze_error:                                                                        # :381-385
        err_t = torch.zeros((B, K), device=dev, dtype=errs.dtype)
    else:
        e = parts._error_from_heatmaps(tgt, hm)
        err_t = flat(torch.from_numpy(e).to(dev, dtype=errs.dtype))
    if parts.freeze_oks:                                                                          # :386-396
        oks_t = torch.zeros((B, K), device=dev, dtype=oks.dtype)
    else:
        oks_t, _ = parts._oks_from_heatmaps(tgt, hm, present & annotated, heatmap_size=(W, H))
        oks_t = flat(oks_t.to(dev).to(oks.dtype))

    use = annotated & (present > 0.5)                                                             # :422
    w_hm = annotated if learn_heatmaps_from_zeros else w_kpt                                      # :425-428
    # the visible / invisible balancing weights of :438-452 are formed by the reference and then ignored (its
    # visibility module has use_target_weight=False): not restated

    losses = {"kpt": parts.keypoint_loss_module(hm, tgt, w_hm, per_pixel=True).mean()}            # :428-431
    losses.update(_scalar_head_losses(flat(probs), flat(vis), flat(oks), flat(errs), present, visible, oks_t, err_t, use))
    losses = {k: losses[k] for k in HEAD_NAMES}
    if not compute_acc:
        return losses
    on = use > 0.5
    acc = {                                                                                       # :463-508
        "kpt": parts.get_pose_accuracy(hm, tgt, w_kpt > 0.5),
        "probability": parts.get_binary_accuracy(flat(probs), present, annotated > 0.5, force_balanced=True)[0],
        "visibility": parts.get_binary_accuracy(flat(vis), visible, on, force_balanced=True)[0],
        "oks": parts.get_mae(flat(oks), oks_t, on),
        "error": parts.get_mae(flat(errs), err_t, on),
    }
    return losses, acc
