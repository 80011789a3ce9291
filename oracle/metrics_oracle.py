"""CPU oracle of the validation metrics  --  TEST INFRASTRUCTURE ONLY.

Restates, in NumPy, ``_calc_distances`` / ``_distance_acc`` (heatmap.py:55-111), ``keypoint_pck_accuracy`` /
``pose_pck_accuracy`` (loss.py:767-866) and the mask-select metrics ``ProbPoseLoss.get_binary_accuracy``
(force_balanced=False) / ``get_mae`` (loss.py:653-712).  Pinned against the reference's own outputs on seeded
inputs: tests/golden/metrics.npz (written by oracle/make_golden_metrics.py, which imports the reference).
"""

from __future__ import annotations

import numpy as np

from .codec_oracle import heatmap_maximum


def calc_distances(preds, gts, mask, norm_factor) -> np.ndarray:
    """(K, N) float32 normalised distances, -1 where masked out (heatmap.py:55-90).  Rows whose
    normalisation has a zero component are masked (:80-81); non-positive components become 1e6 (:85).
    The arithmetic runs in the promoted dtype of ``preds`` and ``norm_factor`` (float64 for the default
    integer ``[H, W]`` factor) and is rounded to float32 on store (:83,86)."""
    preds, gts = np.asarray(preds), np.asarray(gts)
    norm = np.array(norm_factor, copy=True)
    m = np.array(mask, dtype=bool, copy=True)
    m[(norm == 0).sum(1) > 0, :] = False
    norm[norm <= 0] = 1e6
    d = np.full(m.shape, -1, dtype=np.float32)
    q = (preds - gts) / norm[:, None, :]
    d[m] = np.sqrt((np.abs(q) ** 2).sum(-1))[m]
    return d.T


def distance_acc(distances: np.ndarray, thr: float = 0.5):
    """Fraction of valid (!= -1) distances below ``thr`` (compared in float32), -1 if none (heatmap.py:93-111)."""
    valid = distances != -1
    n = valid.sum()
    if n > 0:
        return (distances[valid] < np.float32(thr)).sum() / n
    return -1


def keypoint_pck_accuracy(pred, gt, mask, thr, norm_factor):
    """(acc (K,) float64, avg_acc, cnt) -- loss.py:825-866."""
    d = calc_distances(pred, gt, mask, norm_factor)
    acc = np.array([distance_acc(row, thr) for row in d])
    valid = acc[acc >= 0]
    cnt = len(valid)
    return acc, (valid.mean() if cnt > 0 else 0.0), cnt


def pose_pck_accuracy(output, target, mask, thr: float = 0.05, normalize=None):
    """PCK from heatmaps by plain argmax (loss.py:767-822, method="argmax"; the default normalisation is
    ``[[H, W]]`` applied to (x, y) -- x is divided by H, as the reference does)."""
    N, K, H, W = output.shape
    if K == 0:
        return None, 0, 0
    if normalize is None:
        normalize = np.tile(np.array([[H, W]]), (N, 1))
    pred, _ = heatmap_maximum(output)
    gt, _ = heatmap_maximum(target)
    return keypoint_pck_accuracy(pred, gt, mask, thr, normalize)


def binary_accuracy(dt, gt, mask):
    """Best accuracy over thresholds arange(0.1, 1.0, 0.05) and that threshold, both float32
    (loss.py:653-697 with force_balanced=False)."""
    dt, gt, mask = np.asarray(dt), np.asarray(gt), np.asarray(mask, dtype=bool)
    d, g = dt[mask], gt[mask].astype(bool)
    thresholds = np.arange(0.1, 1.0, 0.05)
    counts = ((d[:, None] > thresholds) == g[:, None]).sum(axis=0)
    best = int(np.argmax(counts))
    return np.float32(counts[best] / len(g)), np.float32(thresholds[best])


def masked_mae(dt, gt, mask):
    """mean |dt - gt| over the mask (loss.py:699-712)."""
    dt, gt, mask = np.asarray(dt), np.asarray(gt), np.asarray(mask, dtype=bool)
    return np.abs(dt[mask] - gt[mask]).mean()
