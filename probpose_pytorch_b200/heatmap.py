"""Drop-in for the hot functions of the reference's ``probpose/heatmap.py``.

Same names, argument order and return structure as ``get_heatmap_maximum``
(heatmap.py:13-52) and ``get_heatmap_expected_value`` (heatmap.py:291-395);
the work runs in the sm_100a kernels of ``csrc/pp_decode.cu``.

Extensions over the reference (which it cannot do at all):
  * a ``torch.Tensor`` on a CUDA device is accepted and then the results are
    returned as CUDA tensors without a host synchronisation;
  * ``(B, K, H, W)`` input with B > 1 works for the expected-OKS decoder (the
    reference raises at heatmap.py:364) and means "each sample decoded on its
    own, results stacked".
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._tables import OksKernelTable

_TABLE_CACHE: dict = {}


def _oks_table(sigmas, K: int, H: int, W: int, device: torch.device) -> OksKernelTable:
    sig = np.ascontiguousarray(np.asarray(sigmas))
    key = (sig.dtype.str, sig.tobytes(), K, H, W, str(device))
    tab = _TABLE_CACHE.get(key)
    if tab is None:
        if len(_TABLE_CACHE) > 64:
            _TABLE_CACHE.clear()
        tab = _TABLE_CACHE[key] = OksKernelTable(sig, K, H, W, device)
    return tab


def _to_device(heatmaps):
    """NumPy -> CUDA tensor (float32), or pass a CUDA tensor through."""
    _lib.require_cuda()
    if isinstance(heatmaps, np.ndarray):
        if heatmaps.dtype != np.float32:
            heatmaps = heatmaps.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(heatmaps)).cuda()
        return t, True
    if not isinstance(heatmaps, torch.Tensor):
        raise AssertionError("heatmaps should be numpy.ndarray or a CUDA torch.Tensor")
    if not heatmaps.is_cuda:
        return heatmaps.detach().cuda(), False
    return heatmaps.detach(), False


def heatmap_maximum_device(heatmaps: torch.Tensor):
    """``get_heatmap_maximum`` on a CUDA tensor (..., H, W): returns ``locs (..., 2)``
    float32, ``vals (...)`` float32 and the flat ``argmax (...)`` int32, all on the device."""
    assert heatmaps.is_cuda and heatmaps.ndim >= 2
    hm = heatmaps.contiguous()
    lead = hm.shape[:-2]
    H, W = hm.shape[-2:]
    N = int(np.prod(lead)) if lead else 1
    dev = hm.device
    locs = torch.empty(lead + (2,), dtype=torch.float32, device=dev)
    vals = torch.empty(lead, dtype=torch.float32, device=dev)
    arg = torch.empty(lead, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_heatmap_maximum(_lib.ptr(hm), _lib.dtype_code(hm.dtype), N, H, W, _lib.ptr(locs),
                                           _lib.ptr(vals), _lib.ptr(arg), _lib.stream_ptr(dev))
    _lib.check(rc, "pp_heatmap_maximum")
    return locs, vals, arg


def get_heatmap_maximum(heatmaps):
    """Get maximum response location and value from heatmaps (heatmap.py:13-52).

    Args:
        heatmaps: (K, H, W) or (B, K, H, W); ``np.ndarray`` (results are NumPy,
            as in the reference) or a CUDA tensor (results stay on the device).

    Returns:
        ``locs`` (K, 2) / (B, K, 2) float32 ``[x, y]`` with -1 where the maximum
        is <= 0, and ``vals`` (K,) / (B, K).
    """
    assert isinstance(heatmaps, (np.ndarray, torch.Tensor)), "heatmaps should be numpy.ndarray"
    assert heatmaps.ndim == 3 or heatmaps.ndim == 4, f"Invalid shape {tuple(heatmaps.shape)}"
    hm, was_numpy = _to_device(heatmaps)
    locs, vals, _ = heatmap_maximum_device(hm)
    if hm.dtype != torch.float32:
        vals = vals.to(hm.dtype)
    if was_numpy:
        return locs.cpu().numpy(), vals.cpu().numpy()
    return locs, vals


def expected_value_device(heatmaps: torch.Tensor, sigmas, *, input_size=None, return_heatmap: bool = False,
                          temperature: float | None = None):
    """Expected-OKS decode of a CUDA tensor (B, K, H, W).

    Returns a dict of device tensors: ``locs`` (B,K,2) float32 heatmap px,
    ``vals`` (B,K) float32, ``argmax`` (B,K) int32 and, when ``input_size`` is
    given, ``keypoints`` (B,K,2) float64 in input-image space; ``conv`` when
    ``return_heatmap``.  ``temperature`` fuses the head tail
    ``clamp(x / temperature, 0, 1)`` (head.py:526-532) into the load.
    """
    assert heatmaps.is_cuda and heatmaps.ndim == 4
    hm = heatmaps.contiguous()
    B, K, H, W = hm.shape
    dev = hm.device
    tab = _oks_table(sigmas, K, H, W, dev)
    out = {
        "locs": torch.empty((B, K, 2), dtype=torch.float32, device=dev),
        "vals": torch.empty((B, K), dtype=torch.float32, device=dev),
        "argmax": torch.empty((B, K), dtype=torch.int32, device=dev),
    }
    kp = None
    if input_size is not None:
        kp = out["keypoints"] = torch.empty((B, K, 2), dtype=torch.float64, device=dev)
    p = _lib.DecodeParams(B, K, H, W, _lib.dtype_code(hm.dtype), int(temperature is not None),
                          float(temperature or 1.0),
                          float(input_size[0]) if input_size is not None else 0.0,
                          float(input_size[1]) if input_size is not None else 0.0)
    conv = None
    # maps too large for the shared-memory kernel go through the exact full-map path, which needs
    # a (B, K, H, W) float32 work buffer; the library says when.
    if return_heatmap or _lib.lib().pp_decode_expected_workspace_floats(p) > 0:
        conv = torch.empty((B, K, H, W), dtype=torch.float32, device=dev)
        if return_heatmap:
            out["conv"] = conv
    t = tab.descriptor()
    # work-queue counters + the hand-over list of the tensor-core kernel (one int32 per heatmap)
    scratch = torch.empty((int(_lib.lib().pp_decode_expected_scratch_bytes_for(p)) + 3) // 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_decode_expected(p, t, _lib.ptr(hm), _lib.ptr(out["locs"]), _lib.ptr(out["vals"]),
                                           _lib.ptr(out["argmax"]), _lib.ptr(kp), _lib.ptr(conv), _lib.ptr(scratch),
                                           scratch.numel() * 4, _lib.stream_ptr(dev))
    _lib.check(rc, "pp_decode_expected")
    out["_scratch"] = scratch   # word 2: number of heatmaps the tensor-core kernel handed on (tools / tests read it)
    return out


def get_heatmap_expected_value(heatmaps, sigmas, parzen_size: float = 0.1, return_heatmap: bool = False,
                               backend: str = "scipy"):
    """OKS-kernel convolution + argmax + sub-pixel refinement (heatmap.py:291-395).

    ``backend`` is accepted for signature compatibility; both reference
    back-ends ("scipy", "torch") compute the same reflect-mode convolution and
    so does the CUDA kernel that runs here.

    Returns ``locs`` (K, 2) / (B, K, 2) float32, ``vals`` (K,) / (B, K) read from
    the unconvolved map at the integer maximum and, with ``return_heatmap``,
    the convolved maps.
    """
    assert isinstance(heatmaps, (np.ndarray, torch.Tensor)), "heatmaps should be numpy.ndarray"
    assert heatmaps.ndim == 3 or heatmaps.ndim == 4, f"Invalid shape {tuple(heatmaps.shape)}"
    assert parzen_size >= 0.0 and parzen_size <= 1.0, f"Invalid parzen_size {parzen_size}"
    hm, was_numpy = _to_device(heatmaps)
    if hm.ndim == 3:
        hm = hm.unsqueeze(0)
    out = expected_value_device(hm, sigmas, return_heatmap=return_heatmap)
    locs, vals = out["locs"], out["vals"]
    if hm.dtype != torch.float32:
        vals = vals.to(hm.dtype)
    conv = out.get("conv")
    if hm.shape[0] == 1:  # 3-D input, or B == 1: the reference drops the batch axis (heatmap.py:387-390)
        locs, vals = locs[0], vals[0]
        conv = conv[0] if conv is not None else None
    if was_numpy:
        locs, vals = locs.cpu().numpy(), vals.cpu().numpy()
        conv = conv.cpu().numpy() if conv is not None else None
    if return_heatmap:
        return locs, vals, conv
    return locs, vals
