"""Seeded synthetic inputs for the five BASELINE.json configurations.

NumPy only.  The same arrays are handed to the CUDA path and (in tests / the
CPU-baseline leg of bench.py) to the oracle.  Nothing here computes a heatmap:
predicted "blob" heatmaps are built by the caller from an encode function, so
that the product can use its own GPU encoder and the tests can use the oracle.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

COCO17_SIGMAS = np.array(
    [0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072,
     0.062, 0.062, 0.107, 0.107, 0.087, 0.087, 0.089, 0.089]
)

_FOOT = [0.068, 0.066, 0.066, 0.092, 0.094, 0.094]
_FACE = [0.042, 0.043, 0.044, 0.043, 0.040, 0.035, 0.031, 0.025, 0.020, 0.023, 0.029, 0.032,
         0.037, 0.038, 0.043, 0.041, 0.045, 0.013, 0.012, 0.011, 0.011, 0.012, 0.012, 0.011,
         0.011, 0.013, 0.015, 0.009, 0.007, 0.007, 0.007, 0.012, 0.009, 0.008, 0.016, 0.010,
         0.017, 0.011, 0.009, 0.011, 0.009, 0.007, 0.013, 0.008, 0.011, 0.012, 0.010, 0.034,
         0.008, 0.008, 0.009, 0.008, 0.008, 0.007, 0.010, 0.008, 0.009, 0.009, 0.009, 0.007,
         0.007, 0.008, 0.011, 0.008, 0.008, 0.008, 0.010, 0.008]
_HAND = [0.029, 0.022, 0.035, 0.037, 0.047, 0.026, 0.025, 0.024, 0.035, 0.018, 0.024, 0.022,
         0.026, 0.017, 0.021, 0.021, 0.032, 0.020, 0.019, 0.022, 0.031]
#: COCO-WholeBody: 17 body + 6 feet + 68 face + 2 x 21 hands = 133 keypoints.
WHOLEBODY133_SIGMAS = np.array(list(COCO17_SIGMAS) + _FOOT + _FACE + _HAND + _HAND)
assert WHOLEBODY133_SIGMAS.shape == (133,)


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration: shapes of the heatmap path."""
    config_id: int
    name: str
    batch: int
    num_keypoints: int
    input_size: tuple   # (w, h)
    heatmap_size: tuple  # (W, H)
    oob: bool           # keypoints drawn from the padded box (out-of-image cases)
    crowd: bool = False

    @property
    def sigmas(self) -> np.ndarray:
        return WHOLEBODY133_SIGMAS if self.num_keypoints == 133 else COCO17_SIGMAS[: self.num_keypoints]

    @property
    def heatmaps(self) -> int:
        return self.batch * self.num_keypoints


WORKLOADS = {
    1: Workload(1, "C1 codec round-trip COCO17 256x192->64x48 B=32", 32, 17, (192, 256), (48, 64), False),
    2: Workload(2, "C2 head decode COCO17 256x192->64x48 B=256", 256, 17, (192, 256), (48, 64), True),
    3: Workload(3, "C3 train step COCO17 64x48 B=1024 (128/GPU x 8)", 1024, 17, (192, 256), (48, 64), True),
    4: Workload(4, "C4 high-res 384x288->96x72 B=512 out-of-image", 512, 17, (288, 384), (72, 96), True),
    5: Workload(5, "C5 COCO-WholeBody133 64x48 B=512 crowd", 512, 133, (192, 256), (48, 64), True, True),
}


def make_keypoints(wl: Workload, batch: int | None = None, seed: int | None = None, dtype=np.float32):
    """Keypoints in input-image pixels (B, K, 2), ``keypoints_visible`` (B, K)
    float32 ~ Bernoulli(0.9) and ``keypoints_visibility`` ~ Bernoulli(0.7)."""
    B = wl.batch if batch is None else batch
    K = wl.num_keypoints
    rng = np.random.default_rng(1000 + wl.config_id if seed is None else seed)
    w, h = wl.input_size
    lo, hi = (-0.125, 1.125) if wl.oob else (0.0, 1.0)
    if wl.crowd:
        centre = rng.uniform(lo, hi, size=(B, 1, 2)) * [w, h]
        kps = centre + rng.normal(0.0, 0.15, size=(B, K, 2)) * [w, h]
    else:
        kps = rng.uniform(lo, hi, size=(B, K, 2)) * [w, h]
    visible = (rng.random((B, K)) < 0.9).astype(np.float32)
    visibility = (rng.random((B, K)) < 0.7).astype(np.float32)
    return kps.astype(dtype), visible, visibility


def jitter_keypoints(wl: Workload, keypoints: np.ndarray, seed: int):
    """GT keypoints moved by N(0, 1.5 heatmap px) -- centres of the predicted blobs."""
    rng = np.random.default_rng(seed)
    scale = (np.array(wl.input_size) - 1) / (np.array(wl.heatmap_size) - 1)
    return (keypoints + rng.normal(0.0, 1.5, size=keypoints.shape) * scale).astype(keypoints.dtype)


def blob_params(shape_bk: tuple, seed: int):
    """Per-heatmap amplitude U(0.3, 1) and the noise seed for blob predictions."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.3, 1.0, size=shape_bk).astype(np.float32)


def blob_predictions_numpy(target_like: np.ndarray, amplitude: np.ndarray, seed: int) -> np.ndarray:
    """``clip(amplitude * maps + U(0, 0.02), 0, 1)`` float32 -- NumPy flavour for tests."""
    rng = np.random.default_rng(seed)
    noise = rng.uniform(0.0, 0.02, size=target_like.shape).astype(np.float32)
    out = target_like * amplitude[..., None, None] + noise
    return np.clip(out, 0.0, 1.0).astype(np.float32)


def uniform_predictions_numpy(shape, seed: int) -> np.ndarray:
    """U(0,1) float32 heatmaps in the style of the reference's tests/test_heatmap.py:6."""
    return np.random.default_rng(seed).random(shape, dtype=np.float32)
