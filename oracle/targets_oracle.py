"""CPU oracle: training targets derived from decoded heatmaps.  TEST INFRASTRUCTURE ONLY.

Restates ``ProbPoseLoss._oks_from_heatmaps`` (loss.py:550-640) with the per-keypoint branch of
``compute_oks(use_area=False, per_kpt=True)`` (loss.py:715-764) and
``ProbPoseLoss._error_from_heatmaps`` (loss.py:512-548)."""

from __future__ import annotations

import numpy as np

from .codec_oracle import decode_argmax_dark


def _decode_both(gt_heatmaps, dt_heatmaps, input_size, heatmap_size, backend):
    B, K = gt_heatmaps.shape[:2]
    gt = np.zeros((B, K, 2))
    dt = np.zeros((B, K, 2))
    for i in range(B):  # per sample, like loss.py:576-585
        gt[i] = decode_argmax_dark(gt_heatmaps[i], input_size, heatmap_size, backend=backend)[0].squeeze()
        dt[i] = decode_argmax_dark(dt_heatmaps[i], input_size, heatmap_size, backend=backend)[0].squeeze()
    return gt, dt


def error_from_heatmaps(gt_heatmaps, dt_heatmaps, input_size, heatmap_size, backend="numpy"):
    gt, dt = _decode_both(gt_heatmaps, dt_heatmaps, input_size, heatmap_size, backend)
    gt[np.isnan(gt)] = -1                                   # loss.py:541
    return np.linalg.norm(gt - dt, axis=2)                  # loss.py:544


def oks_from_heatmaps(gt_heatmaps, dt_heatmaps, weight, sigmas, input_size, heatmap_size, area_size=(48, 64),
                      backend="numpy"):
    """Returns ``(target_oks (B, K) float32, oks_weights (B,) float32)``."""
    B, K = gt_heatmaps.shape[:2]
    gt, dt = _decode_both(gt_heatmaps, dt_heatmaps, input_size, heatmap_size, backend)
    w = np.asarray(weight).squeeze().reshape((B, K, 1))
    gt[np.isnan(gt)] = 0                                    # loss.py:588
    gt, dt = gt * w, dt * w                                 # loss.py:591-592
    vis = (w * 2)[..., 0]
    variances = (np.asarray(sigmas)[:K] * 2) ** 2
    area = area_size[0] * area_size[1] * 0.53               # bbox[3] * bbox[2] * 0.53, loss.py:751
    oks = np.zeros((B, K))
    weights = np.zeros(B)
    for i in range(B):
        valid = vis[i] > 0
        if not valid.any():                                 # loss.py:600-604
            continue
        dx, dy = dt[i, :, 0] - gt[i, :, 0], dt[i, :, 1] - gt[i, :, 1]
        e = (dx ** 2 + dy ** 2) / variances / (area + np.spacing(1)) / 2
        o = np.exp(-e)
        o[~valid] = 0                                       # loss.py:757-758
        oks[i] = o
        weights[i] = 1
    return oks.astype(np.float32), weights.astype(np.float32)
