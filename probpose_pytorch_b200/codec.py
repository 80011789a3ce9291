"""Drop-in for the reference's ``probpose/codec.py``: ``generate_probmaps``, ``ProbMap``,
``ArgMaxProbMap`` and ``Codec`` with the reference's call signatures, defaults,
return structures and dtypes (codec.py:11-17, 117-126, 138-144, 214, 242-279,
422-430, 443-447, 515).  The arithmetic runs in ``csrc/pp_encode.cu`` and
``csrc/pp_decode.cu``.

Beyond the reference:
  * decoders take whole batches -- ``(B, K, H, W)`` tensors are decoded sample
    by sample on the GPU and stacked (the reference only handles B == 1,
    heatmap.py:364, codec.py:339);
  * ``encode_batch`` / ``decode_device`` keep everything on the device (no
    host round trip of B*K*H*W floats, cf. dataset.py:128 and loss.py:568-569).
"""

from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._tables import encode_divisors, gaussian_taps
from .heatmap import expected_value_device

_TORCH_OF_NUMPY = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def _current_device() -> torch.device:
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def encode_device(keypoints: Tensor, visible: Tensor | None, divisors: Tensor, heatmap_size, scale_factor,
                  input_size, *, dtype: torch.dtype = torch.float32, flags: bool = True):
    """Encode (B, K, D) device keypoints (input-image space, float32 or float64)
    into ``(B, K, H, W)`` OKS probability maps on the device.

    ``divisors`` is the (K,) float64 device table of ``2 s`` values.  Returns a
    dict with ``heatmaps``, ``keypoint_weights`` (float32) and, with ``flags``,
    ``in_image`` / ``annotated`` (bool).
    """
    assert keypoints.is_cuda and keypoints.ndim == 3 and keypoints.shape[-1] >= 2
    kp = keypoints.contiguous()
    B, K, D = kp.shape
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    dev = kp.device
    vis = None
    if visible is not None:
        vis = visible.to(device=dev, dtype=torch.float32).contiguous()
        assert vis.shape == (B, K)
    out = {
        "heatmaps": torch.empty((B, K, H, W), dtype=dtype, device=dev),
        "keypoint_weights": torch.empty((B, K), dtype=torch.float32, device=dev),
    }
    inside = ann = None
    if flags:
        # torch.bool is one byte holding 0/1, which is exactly what the kernel writes
        inside = torch.empty((B, K), dtype=torch.bool, device=dev)
        ann = torch.empty((B, K), dtype=torch.bool, device=dev)
    p = _lib.EncodeParams(B, K, H, W, _lib.dtype_code(dtype), _lib.dtype_code(kp.dtype), D,
                          float(scale_factor[0]), float(scale_factor[1]), float(input_size[0]), float(input_size[1]))
    assert divisors.dtype == torch.float64 and divisors.numel() >= K and divisors.device == dev
    with torch.cuda.device(dev):
        rc = _lib.lib().pp_encode(p, _lib.ptr(kp), _lib.ptr(vis), _lib.ptr(divisors), _lib.ptr(out["heatmaps"]),
                                  _lib.ptr(out["keypoint_weights"]), _lib.ptr(inside), _lib.ptr(ann),
                                  _lib.stream_ptr(dev))
    _lib.check(rc, "pp_encode")
    if flags:
        out["in_image"] = inside
        out["annotated"] = ann
    return out


def generate_probmaps(heatmap_size, keypoints: np.ndarray, keypoints_visible: np.ndarray, sigmas: np.ndarray,
                      sigma: float = 0.55):
    """Generate OKS probability maps (codec.py:11-70).

    Args:
        heatmap_size: ``[W, H]``
        keypoints: heatmap-space coordinates (N, K, D); N must be 1 like in every
            reference call site (later instances would overwrite earlier ones).
        keypoints_visible: (N, K)
        sigmas: per-keypoint sigmas; ``sigma > 0`` overrides them (codec.py:63-64).

    Returns:
        ``heatmaps`` (K, H, W) float32 and ``keypoint_weights`` (N, K) in the
        dtype of ``keypoints_visible``.
    """
    N, K, _ = keypoints.shape
    assert N == 1, "generate_probmaps: only single-instance keypoints (N == 1) are supported"
    W, H = heatmap_size
    dev = _current_device()
    kp = np.ascontiguousarray(keypoints)
    if kp.dtype not in _TORCH_OF_NUMPY:
        kp = kp.astype(np.float64)
    div = torch.from_numpy(encode_divisors(sigmas, sigma, K, H, W)).to(dev)
    vis32 = torch.from_numpy(np.ascontiguousarray(keypoints_visible).astype(np.float32)).to(dev)
    out = encode_device(torch.from_numpy(kp).to(dev), vis32, div, (W, H), (1.0, 1.0), (W, H), flags=False)
    weights = out["keypoint_weights"].cpu().numpy().astype(keypoints_visible.dtype)
    return out["heatmaps"][0].cpu().numpy(), weights


class _ProbMapBase:
    """Shared implementation of ``ProbMap`` and ``ArgMaxProbMap``."""

    _has_heatmap_keypoints = False

    def __init__(self, input_size, heatmap_size, sigmas, sigma, radius_factor, blur_kernel_size,
                 increase_sigma_with_padding) -> None:
        self.input_size = input_size
        self.heatmap_size = heatmap_size
        self.radius_factor = radius_factor
        self.blur_kernel_size = blur_kernel_size
        self.scale_factor = ((np.array(input_size) - 1) / (np.array(heatmap_size) - 1)).astype(np.float32)
        self.increase_sigma_with_padding = increase_sigma_with_padding
        self.sigmas = sigmas
        self.sigma = sigma
        self._device_tables: dict = {}

    # -- constant tables, uploaded once per device --------------------------------------------
    def _divisors(self, K: int, dev: torch.device) -> Tensor:
        key = ("div", K, str(dev), self.sigma)
        t = self._device_tables.get(key)
        if t is None:
            W, H = self.heatmap_size
            t = torch.from_numpy(encode_divisors(self.sigmas, self.sigma, K, H, W)).to(dev)
            self._device_tables[key] = t
        return t

    def _blur_taps(self, dev: torch.device) -> Tensor:
        key = ("blur", self.blur_kernel_size, str(dev))
        t = self._device_tables.get(key)
        if t is None:
            t = torch.from_numpy(gaussian_taps(self.blur_kernel_size)).to(dev)
            self._device_tables[key] = t
        return t

    def _blur_mma_table(self, dev: torch.device, H: int, W: int):
        """Operand table of the DARK decoder's tensor-core kernel for (H, W, blur_kernel_size), built once per device;
        None for shapes / kernel sizes without such a kernel."""
        key = ("blur_mma", self.blur_kernel_size, H, W, str(dev))
        if key not in self._device_tables:
            L = _lib.lib()
            nbytes = int(L.pp_oks_mma_table_bytes(1, H, W)) if self.blur_kernel_size <= 15 else 0
            t = None
            if nbytes > 0:
                t = torch.empty(nbytes // 2, dtype=torch.float16, device=dev)
                with torch.cuda.device(dev):
                    rc = L.pp_blur_mma_table_build(_lib.ptr(self._blur_taps(dev)), int(self.blur_kernel_size), H, W, _lib.ptr(t),
                                                   _lib.stream_ptr(dev))
                _lib.check(rc, "pp_blur_mma_table_build")
            self._device_tables[key] = t
        return self._device_tables[key]

    # -- encode -------------------------------------------------------------------------------
    def encode_batch(self, keypoints, keypoints_visible=None, *, dtype: torch.dtype = torch.float32,
                     device: torch.device | None = None) -> dict:
        """Batched, device-resident encode: ``keypoints`` (B, K, D) in input-image
        space (tensor or array) -> dict of device tensors (``heatmaps`` (B,K,H,W),
        ``keypoint_weights``, ``in_image``, ``annotated``)."""
        dev = device or (keypoints.device if isinstance(keypoints, Tensor) and keypoints.is_cuda else _current_device())
        kp = torch.as_tensor(keypoints)
        if kp.dtype not in (torch.float32, torch.float64):
            kp = kp.to(torch.float64)
        kp = kp.to(dev, non_blocking=True)
        vis = None
        if keypoints_visible is not None:
            vis = torch.as_tensor(keypoints_visible).to(dev, non_blocking=True)
        return encode_device(kp, vis, self._divisors(kp.shape[1], dev), self.heatmap_size, self.scale_factor,
                             self.input_size, dtype=dtype)

    def encode(self, keypoints: np.ndarray, keypoints_visible: np.ndarray | None = None,
               id_similarity: float | None = 0.0, keypoints_visibility: np.ndarray | None = None) -> dict:
        """Encode keypoints (1, K, D), given in input-image space, into heatmaps
        (codec.py:138-212 / 443-513).  Returns the reference's dict of NumPy arrays."""
        assert keypoints.shape[0] == 1, (
            f"{self.__class__.__name__} only support single-instance keypoint encoding")
        if keypoints_visibility is None:
            keypoints_visibility = np.zeros(keypoints.shape[:2], dtype=np.float32)
        if keypoints_visible is None:
            keypoints_visible = np.ones(keypoints.shape[:2], dtype=np.float32)

        kp = np.ascontiguousarray(keypoints)
        if kp.dtype not in _TORCH_OF_NUMPY:
            kp = kp.astype(np.float64)
        vis32 = np.ascontiguousarray(keypoints_visible).astype(np.float32)
        dev_out = self.encode_batch(kp, vis32)

        # weights keep the dtype of keypoints_visible (codec.py:46); flags are boolean arrays
        weights = dev_out["keypoint_weights"].cpu().numpy().astype(np.asarray(keypoints_visible).dtype)
        encoded = dict(
            heatmaps=dev_out["heatmaps"][0].cpu().numpy(),
            keypoint_weights=weights,
            annotated=dev_out["annotated"].cpu().numpy(),
            in_image=dev_out["in_image"].cpu().numpy(),
            keypoints_scaled=keypoints,
            identification_similarity=id_similarity,
        )
        if self._has_heatmap_keypoints:
            encoded["heatmap_keypoints"] = keypoints / self.scale_factor
        return encoded

    # -- decode -------------------------------------------------------------------------------
    def decode_device(self, heatmaps: Tensor, *, temperature: float | None = None) -> dict:
        raise NotImplementedError

    def decode(self, encoded) -> tuple[np.ndarray, np.ndarray]:
        """Decode keypoint coordinates (input-image space) from heatmaps.

        ``encoded``: (K, H, W) -- as in the reference -- or a batch (B, K, H, W);
        NumPy array or torch tensor.  Returns ``keypoints`` (N, K, 2) float64 and
        ``scores`` (N, K) float32 with N = 1 for a single sample, B for a batch.
        """
        hm = _as_device_heatmaps(encoded)
        out = self.decode_device(hm)
        return out["keypoints"].cpu().numpy(), out["scores"].cpu().numpy()


def _as_device_heatmaps(encoded) -> Tensor:
    _lib.require_cuda()
    if isinstance(encoded, np.ndarray):
        arr = encoded if encoded.dtype == np.float32 else encoded.astype(np.float32)
        hm = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    elif isinstance(encoded, Tensor):
        hm = encoded.detach()
        if not hm.is_cuda:
            hm = hm.cuda()
    else:
        raise TypeError("heatmaps must be a numpy.ndarray or a torch.Tensor")
    if hm.ndim == 3:
        hm = hm.unsqueeze(0)
    if hm.ndim != 4:
        raise ValueError(f"Invalid heatmap shape {tuple(hm.shape)}")
    return hm


class ProbMap(_ProbMapBase):
    r"""Per-pixel expected-OKS heatmap codec (ProbPose, arXiv:2412.02254) with the
    expected-OKS decoder; same constructor and methods as the reference's
    ``ProbMap`` (codec.py:73-239).

    Args:
        input_size: image size ``[w, h]``
        heatmap_size: heatmap size ``[W, H]``
        sigmas: per-keypoint sigmas
        sigma: scalar variance override of the targets; ``<= 0`` uses the
            per-keypoint table.  Defaults to 2.0 (codec.py:122).
        radius_factor, increase_sigma_with_padding: stored, unused (as in the reference).
        blur_kernel_size: stored (used by the argmax codec).
    """

    _has_heatmap_keypoints = True

    def __init__(self, input_size, heatmap_size, sigmas, sigma: float = 2.0, radius_factor: float = 0.0546875,
                 blur_kernel_size: int = 11, increase_sigma_with_padding=False) -> None:
        super().__init__(input_size, heatmap_size, sigmas, sigma, radius_factor, blur_kernel_size,
                         increase_sigma_with_padding)

    def decode_device(self, heatmaps: Tensor, *, temperature: float | None = None) -> dict:
        """Expected-OKS decode of (B, K, H, W) on the device (codec.py:214-239 per sample):
        ``keypoints`` (B,K,2) float64 input space, ``scores`` (B,K) float32, ``locs``, ``argmax``."""
        out = expected_value_device(heatmaps, self.sigmas, input_size=self.input_size, temperature=temperature)
        out["scores"] = out["vals"]
        return out


class ArgMaxProbMap(_ProbMapBase):
    r"""Expected-OKS heatmap codec decoded by argmax + DARK-UDP refinement; same
    constructor and methods as the reference's ``ArgMaxProbMap`` (codec.py:377-543).
    ``sigma`` defaults to -1, i.e. per-keypoint variances (codec.py:426)."""

    def __init__(self, input_size, heatmap_size, sigmas=None, sigma: float = -1, radius_factor: float = 0.0546875,
                 blur_kernel_size: int = 11, increase_sigma_with_padding=False) -> None:
        super().__init__(input_size, heatmap_size, sigmas, sigma, radius_factor, blur_kernel_size,
                         increase_sigma_with_padding)

    def decode_device(self, heatmaps: Tensor, *, temperature: float | None = None) -> dict:
        """argmax + Gaussian modulation + DARK-UDP refinement of (B, K, H, W) on the device
        (codec.py:515-543 per sample).  Empty channels (max <= 0) keep the (-1, -1) sentinel."""
        assert heatmaps.is_cuda and heatmaps.ndim == 4
        hm = heatmaps.contiguous()
        B, K, H, W = hm.shape
        dev = hm.device
        out = {
            "peaks": torch.empty((B, K, 2), dtype=torch.float32, device=dev),
            "scores": torch.empty((B, K), dtype=torch.float32, device=dev),
            "locs": torch.empty((B, K, 2), dtype=torch.float32, device=dev),
            "keypoints": torch.empty((B, K, 2), dtype=torch.float64, device=dev),
        }
        p = _lib.DecodeParams(B, K, H, W, _lib.dtype_code(hm.dtype), int(temperature is not None),
                              float(temperature or 1.0), float(self.input_size[0]), float(self.input_size[1]))
        taps = self._blur_taps(dev)
        table = self._blur_mma_table(dev, H, W)
        # work-queue counters + the hand-over list of the tensor-core kernel (one int32 per heatmap)
        scratch = torch.empty((int(_lib.lib().pp_decode_expected_scratch_bytes_for(p)) + 3) // 4, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().pp_decode_argmax_dark(p, _lib.ptr(taps), int(self.blur_kernel_size), _lib.ptr(table), _lib.ptr(hm),
                                                  _lib.ptr(out["peaks"]), _lib.ptr(out["scores"]),
                                                  _lib.ptr(out["locs"]), _lib.ptr(out["keypoints"]),
                                                  _lib.ptr(scratch), scratch.numel() * 4, _lib.stream_ptr(dev))
        _lib.check(rc, "pp_decode_argmax_dark")
        out["_scratch"] = scratch   # word 2: number of heatmaps the tensor-core kernel handed on
        return out


class Codec:
    """Adapter between the model's 5-tuple and a probmap codec (codec.py:242-279)."""

    def __init__(self, probmap):
        self.probmap = probmap

    def decode(self, pred: tuple[Tensor, Tensor, Tensor, Tensor, Tensor]):
        """``pred = (heatmaps (B,K,H,W), probabilities, visibilities, oks, errors (B,K,1,1))``
        -> ``((keypoints, scores), probabilities, visibilities, oks, errors)`` as NumPy, the four
        scalar heads reshaped to (B, 1, K) and the error divided by the heatmap diagonal
        (codec.py:249-263).  Only the small decoded records cross to the host."""
        heatmaps, probabilities, visibilities, oks, errors = pred
        B, C, H, W = heatmaps.shape
        preds = self.probmap.decode(heatmaps)

        def host(t):
            a = t.detach().cpu().numpy() if isinstance(t, Tensor) else np.asarray(t)
            return a.reshape((B, 1, C))

        probabilities, visibilities, oks, errors = host(probabilities), host(visibilities), host(oks), host(errors)
        errors = errors / np.sqrt(H ** 2 + W ** 2)
        return preds, probabilities, visibilities, oks, errors

    def decode_device(self, pred, *, temperature: float | None = None, mailbox=None, slot: int = 0,
                      loss: Tensor | None = None) -> Tensor:
        """Fully device-resident decode: returns one packed (B, K, 7) float64 tensor of records
        ``(x, y, score, probability, visibility, oks, error / diagonal)`` -- the unit that is
        gathered across GPUs.  The records are packed by one kernel (``pp_pack_records``); with a
        :class:`~probpose_pytorch_b200.distributed.PeerMailbox` the same kernel also stores them into ``slot`` of every
        rank's mailbox over NVLink; ``loss`` (a device scalar) then completes the publication (``mailbox.commit``) --
        without it the caller commits later, once the step's loss exists."""
        heatmaps, probabilities, visibilities, oks, errors = pred
        out = self.probmap.decode_device(heatmaps if heatmaps.ndim == 4 else heatmaps.unsqueeze(0),
                                         temperature=temperature)
        return self.pack_records(out, pred, mailbox=mailbox, slot=slot, loss=loss)

    def pack_records(self, decoded: dict, pred, *, mailbox=None, slot: int = 0, loss: Tensor | None = None) -> Tensor:
        """The record-packing half of :meth:`decode_device` on an already decoded batch (``probmap.decode_device``
        output); lets a caller publish the records together with a loss that is computed later in the step."""
        heatmaps, probabilities, visibilities, oks, errors = pred
        B, C, H, W = heatmaps.shape if heatmaps.ndim == 4 else (1,) + tuple(heatmaps.shape)
        dev = heatmaps.device
        n = B * C
        heads = [h.reshape(-1).to(torch.float32).contiguous() for h in (probabilities, visibilities, oks, errors)]
        for h in heads:
            assert h.numel() == n, "scalar heads must hold one value per keypoint"
        kp = decoded["keypoints"].reshape(n, 2)
        sc = decoded["scores"].reshape(n)
        assert kp.dtype == torch.float64 and sc.dtype == torch.float32 and kp.is_contiguous() and sc.is_contiguous()
        rec = torch.empty((B, C, 7), dtype=torch.float64, device=dev)
        inv_diag = float(np.float32(1.0) / np.float32(np.sqrt(H ** 2 + W ** 2)))   # what torch's f32 / scalar multiplies by
        mb = None
        if mailbox is not None:
            assert mailbox.n_records == n, f"mailbox built for {mailbox.n_records} records, got {n}"
            mb = mailbox.descriptor(slot)
        with torch.cuda.device(dev):
            rc = _lib.lib().pp_pack_records(n, _lib.ptr(kp), _lib.ptr(sc), *[_lib.ptr(h) for h in heads], inv_diag,
                                            _lib.ptr(rec), mb, _lib.stream_ptr(dev))
        _lib.check(rc, "pp_pack_records")
        if mailbox is not None and loss is not None:
            mailbox.commit(slot, loss)
        return rec

    def decode_heatmap(self, heatmaps):
        return self.probmap.decode(heatmaps)

    def encode(self, keypoints: np.ndarray, keypoints_visible: np.ndarray | None = None,
               id_similarity: float | None = 0.0) -> dict:
        return self.probmap.encode(keypoints=keypoints, keypoints_visible=keypoints_visible,
                                   id_similarity=id_similarity)
