"""Host-side constant tables of a codec (computed once in NumPy, uploaded once per device).

These are the per-keypoint constants the reference recomputes on every call:
the OKS variance ``s`` (codec.py:48,60-62; heatmap.py:171,176-178), the
normalised d x d OKS kernels (heatmap.py:170-194) and the 11-tap Gaussian of the
DARK blur (``cv2.getGaussianKernel(11, 0)`` behind codec.py:310).
"""

from __future__ import annotations

import numpy as np
import torch

from ._lib import PP_MAX_BLUR_KSIZE, PP_MAX_OKS_RADIUS, PP_OKS_TAPS


def oks_variance(sigmas, H: int, W: int) -> np.ndarray:
    """``s_k = clip((2 sigma_k)^2 * sqrt(H/1.25 * W/1.25) * 2, 0.55, 3.0)`` in float64."""
    sig = np.asarray(sigmas)
    box = np.sqrt(H / 1.25 * W / 1.25)
    return np.clip((sig * 2) ** 2 * box * 2, 0.55, 3.0).astype(np.float64)


def encode_divisors(sigmas, sigma, K: int, H: int, W: int) -> np.ndarray:
    """The divisor ``2 s`` of codec.py:65 per keypoint; a positive scalar ``sigma``
    overrides the per-keypoint table (codec.py:63-64)."""
    if sigma is not None and sigma > 0:
        return np.full(K, 2 * float(sigma), dtype=np.float64)
    if sigmas is None:
        raise ValueError("per-keypoint sigmas are required when sigma <= 0")
    sig = np.asarray(sigmas)
    if sig.shape[0] < K:
        raise IndexError(f"{sig.shape[0]} sigmas for {K} keypoints")
    return 2 * oks_variance(sig[:K], H, W)


class OksKernelTable:
    """Device-resident table for the expected-OKS decoder of one (K, H, W, sigmas)."""

    def __init__(self, sigmas, K: int, H: int, W: int, device: torch.device):
        sig = np.asarray(sigmas)
        if sig.shape[0] < K:
            raise IndexError(f"{sig.shape[0]} sigmas for {K} keypoints")
        s = oks_variance(sig[:K], H, W)
        radius = np.ceil(s * 3).astype(np.int32)
        if radius.max() > PP_MAX_OKS_RADIUS or radius.min() < 1:
            raise ValueError("OKS kernel radius out of range")
        taps = np.zeros((K, PP_OKS_TAPS), dtype=np.float32)
        k2d = np.zeros((K, PP_OKS_TAPS * PP_OKS_TAPS), dtype=np.float64)
        for k in range(K):
            r = int(radius[k])
            ax = np.arange(-r, r + 1)
            # 2-D table with the reference's operation order (sqrt, then square; heatmap.py:186-189)
            dist = np.sqrt(ax[None, :] ** 2 + ax[:, None] ** 2)
            ker = np.exp(-(dist ** 2) / (2 * s[k]))
            ker = ker / ker.sum()
            k2d[k, : ker.size] = ker.ravel()
            one = np.exp(-(ax.astype(np.float64) ** 2) / (2 * s[k]))
            taps[k, : 2 * r + 1] = (one / one.sum()).astype(np.float32)
        self.radius = torch.from_numpy(radius).to(device)
        self.taps = torch.from_numpy(taps).to(device)
        self.kernel2d = torch.from_numpy(k2d).to(device)
        self.max_radius = int(radius.max())
        # work-queue order of the decoder: widest kernels (most expensive heatmaps) first
        self.order = torch.from_numpy(np.argsort(-radius, kind="stable").astype(np.int32)).to(device)
        # operand tables of the tensor-core prefilter (csrc/pp_decode_mma.cuh): one per distinct (radius, taps) row,
        # built on the device by the library for the shapes it has such a kernel for
        self.mma_tables = self.mma_index = None
        self.shape = (H, W)
        uniq, index = np.unique(np.concatenate([radius[:, None].astype(np.float32), taps], axis=1), axis=0,
                                return_inverse=True)
        U = int(uniq.shape[0])
        from . import _lib
        nbytes = int(_lib.lib().pp_oks_mma_table_bytes(U, H, W)) if device.type == "cuda" else 0
        if nbytes > 0:
            u_taps = torch.from_numpy(np.ascontiguousarray(uniq[:, 1:])).to(device)
            u_radius = torch.from_numpy(uniq[:, 0].astype(np.int32)).to(device)
            tables = torch.empty(nbytes // 2, dtype=torch.float16, device=device)
            with torch.cuda.device(device):
                rc = _lib.lib().pp_oks_mma_table_build(_lib.ptr(u_taps), _lib.ptr(u_radius), U, H, W, _lib.ptr(tables),
                                                       _lib.stream_ptr(device))
            _lib.check(rc, "pp_oks_mma_table_build")
            self.mma_tables = tables
            self.mma_index = torch.from_numpy(index.reshape(-1).astype(np.int32)).to(device)

    def descriptor(self):
        """The ``pp_oks_table`` struct for this table."""
        from . import _lib
        H, W = self.shape
        return _lib.OksTable(self.radius.data_ptr(), self.taps.data_ptr(), self.kernel2d.data_ptr(), self.order.data_ptr(),
                             self.mma_tables.data_ptr() if self.mma_tables is not None else None,
                             self.mma_index.data_ptr() if self.mma_index is not None else None, H, W)


def gaussian_taps(ksize: int) -> np.ndarray:
    """float32 taps of ``cv2.getGaussianKernel(ksize, 0)``: sigma = 0.3*((ksize-1)*0.5-1)+0.8,
    exp(-x^2 / (2 sigma^2)) normalised in double, rounded to float32."""
    if ksize % 2 != 1 or not 3 <= ksize <= PP_MAX_BLUR_KSIZE:
        raise ValueError(f"blur kernel size must be odd and in [3, {PP_MAX_BLUR_KSIZE}], got {ksize}")
    if ksize <= 7:
        # OpenCV uses fixed small tables for ksize <= 7 when sigma <= 0
        fixed = {3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
                 7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
        return np.asarray(fixed[ksize], dtype=np.float32)
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (t / t.sum()).astype(np.float32)
