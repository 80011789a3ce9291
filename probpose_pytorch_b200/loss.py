"""Drop-in for ``OKSHeatmapLoss`` of the reference's ``probpose/loss.py`` (loss.py:18-191).

Same constructor, same ``forward`` signature, same three reduction modes and
mask semantics; forward and backward run in the fused sm_100a kernel of
``csrc/pp_loss.cu`` (no autograd graph of ~20 elementwise/conv kernels).

``forward_mean`` is the fast path for what ``ProbPoseLoss`` does with this
module (``per_pixel=True`` followed by ``.mean()``, loss.py:428-431): one kernel
reads ``output`` and ``target`` once, reduces the loss and writes
``d loss / d output`` in the same pass.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

from . import _lib

_OKS_TYPES = {"minus": 0, "plus": 1, "both": 2}


def _scratch(p: _lib.LossParams, dev: torch.device) -> Tensor:
    n = int(_lib.lib().pp_oks_loss_scratch_bytes(p))
    return torch.empty((n + 7) // 8, dtype=torch.float64, device=dev)


def _check_loss_out(loss_out, output):
    if loss_out is None:
        return None
    if not (loss_out.dtype == torch.float32 and loss_out.numel() == 1 and loss_out.device == output.device and loss_out.is_contiguous()):
        raise ValueError("loss_out must be a (1,) float32 tensor on the device of `output`")
    return loss_out.view(1)


class _Prepared:
    """Validated, contiguous inputs + the parameter block shared by forward and backward."""

    def __init__(self, module: "OKSHeatmapLoss", output: Tensor, target: Tensor, target_weights, mask, mode: int):
        _lib.require_cuda()
        if not output.is_cuda:
            raise RuntimeError("OKSHeatmapLoss (B200) needs CUDA tensors; there is no CPU fallback")
        if output.ndim != 4 or output.shape != target.shape:
            raise ValueError(f"output {tuple(output.shape)} and target {tuple(target.shape)} must be equal (B,K,H,W)")
        B, K, H, W = output.shape
        dt = output.dtype
        dev = output.device
        self.shape, self.dtype, self.device = (B, K, H, W), dt, dev
        self.output = output.detach().contiguous()
        self.target = target.detach().to(device=dev, dtype=dt).contiguous()

        self.kp_weights = self.pix_weights = self.mask = None
        sb = sk = 0
        if mask is not None:  # loss.py:155-162
            assert (mask.ndim == target.ndim and all(
                d_m == d_t or d_m == 1 for d_m, d_t in zip(mask.shape, target.shape))), (
                f"mask and target have mismatched shapes {mask.shape} v.s.{target.shape}")
            m = mask.detach().to(device=dev, dtype=dt)
            if m.shape[2:] != (H, W) or m.shape[0] != B:
                m = m.expand(B, m.shape[1], H, W)
            m = m.contiguous()
            self.mask = m
            sb = m.shape[1] * H * W
            sk = H * W if m.shape[1] == K else 0
        if target_weights is not None:  # loss.py:165-178
            assert (target_weights.ndim in (2, 4)
                    and target_weights.shape == target.shape[:target_weights.ndim]), (
                "target_weights and target have mismatched shapes "
                f"{target_weights.shape} v.s. {target.shape}")
            if target_weights.ndim == 2:
                self.kp_weights = target_weights.detach().to(device=dev, dtype=torch.float32).contiguous()
            else:
                self.pix_weights = target_weights.detach().to(device=dev, dtype=dt).contiguous()
        self.params = _lib.LossParams(B, K, H, W, _lib.dtype_code(dt), mode, _OKS_TYPES[module.oks_type],
                                      int(module.skip_empty_channel), float(module.smoothing_weight),
                                      float(module.gaussian_weight), float(module.loss_weight), sb, sk)
        self.scratch = _scratch(self.params, dev)

    publish = None   # pp_mailbox descriptor: the loss' finalize kernel also publishes it (multi-GPU exchange)
    scalar_out = None   # optional caller-owned (1,) float32 tensor that receives the scalar loss

    def forward(self, *, want_grad: bool, grad_scale: float = 1.0):
        B, K, H, W = self.shape
        dev, mode = self.device, self.params.mode
        loss_map = loss_kpt = peak = grad = None
        scalar = self.scalar_out if self.scalar_out is not None else torch.empty(1, dtype=torch.float32, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        if mode == _lib.PP_LOSS_PER_PIXEL:
            loss_map = torch.empty(self.shape, dtype=self.dtype, device=dev)
        elif mode == _lib.PP_LOSS_PER_KEYPOINT:
            loss_kpt = torch.empty((B, K), dtype=torch.float32, device=dev)
            peak = torch.empty((B, K), dtype=torch.int32, device=dev)
        if want_grad:
            grad = torch.empty(self.shape, dtype=self.dtype, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().pp_oks_loss_forward(
                self.params, _lib.ptr(self.output), _lib.ptr(self.target), _lib.ptr(self.kp_weights),
                _lib.ptr(self.pix_weights), _lib.ptr(self.mask), _lib.ptr(loss_map), _lib.ptr(loss_kpt),
                _lib.ptr(scalar), _lib.ptr(peak), _lib.ptr(grad), float(grad_scale), _lib.ptr(flag),
                _lib.ptr(self.scratch), self.scratch.numel() * 8, self.publish, _lib.stream_ptr(dev))
        _lib.check(rc, "pp_oks_loss_forward")
        return dict(loss_map=loss_map, loss_kpt=loss_kpt, scalar=scalar, peak=peak, grad=grad, flag=flag)

    def backward(self, upstream: Tensor, kind: int, peak) -> Tensor:
        dev = self.device
        grad = torch.empty(self.shape, dtype=self.dtype, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().pp_oks_loss_backward(
                self.params, _lib.ptr(self.output), _lib.ptr(self.target), _lib.ptr(self.kp_weights),
                _lib.ptr(self.pix_weights), _lib.ptr(self.mask), _lib.ptr(upstream), kind, _lib.ptr(peak),
                _lib.ptr(grad), _lib.ptr(self.scratch), self.scratch.numel() * 8, _lib.stream_ptr(dev))
        _lib.check(rc, "pp_oks_loss_backward")
        return grad


class _PreparedEncoded:
    """Inputs of the encode-inside-loss pass (``pp_oks_loss_forward_encoded``): the target is never materialised,
    the kernel forms it from the keypoints.  Mirrors :class:`_Prepared` for the autograd bridge."""

    def __init__(self, module: "OKSHeatmapLoss", output: Tensor, probmap, keypoints, keypoints_visible, target_weights):
        _lib.require_cuda()
        if not output.is_cuda or output.ndim != 4:
            raise RuntimeError("OKSHeatmapLoss (B200) needs (B, K, H, W) CUDA tensors; there is no CPU fallback")
        B, K, H, W = output.shape
        dev, dt = output.device, output.dtype
        if (W, H) != tuple(int(v) for v in probmap.heatmap_size):
            raise ValueError(f"output is {W}x{H}, the codec encodes {tuple(probmap.heatmap_size)}")
        self.shape, self.dtype, self.device = (B, K, H, W), dt, dev
        self.output = output.detach().contiguous()
        kp = torch.as_tensor(keypoints)
        if kp.dtype not in (torch.float32, torch.float64):
            kp = kp.to(torch.float64)
        self.keypoints = kp.to(dev).reshape(B, K, -1).contiguous()
        self.visible = None
        if keypoints_visible is not None:
            self.visible = torch.as_tensor(keypoints_visible).to(device=dev, dtype=torch.float32).reshape(B, K).contiguous()
        self.kp_weights = None
        if target_weights is not None:
            assert target_weights.shape == (B, K), "the fused encode + loss pass takes per-keypoint weights (B, K)"
            self.kp_weights = target_weights.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.divisors = probmap._divisors(K, dev)
        self.params = _lib.LossParams(B, K, H, W, _lib.dtype_code(dt), _lib.PP_LOSS_PIXEL_MEAN, _OKS_TYPES[module.oks_type],
                                      int(module.skip_empty_channel), float(module.smoothing_weight),
                                      float(module.gaussian_weight), float(module.loss_weight), 0, 0)
        sf = probmap.scale_factor
        self.enc_params = _lib.EncodeParams(B, K, H, W, _lib.dtype_code(dt), _lib.dtype_code(self.keypoints.dtype),
                                            self.keypoints.shape[-1], float(sf[0]), float(sf[1]),
                                            float(probmap.input_size[0]), float(probmap.input_size[1]))
        self.scratch = _scratch(self.params, dev)
        self.encoded = {"keypoint_weights": torch.empty((B, K), dtype=torch.float32, device=dev),
                        "in_image": torch.empty((B, K), dtype=torch.bool, device=dev),
                        "annotated": torch.empty((B, K), dtype=torch.bool, device=dev)}
        self.last_flag = torch.zeros(1, dtype=torch.int32, device=dev)   # an encoded target is in [0, 1] by construction

    publish = None
    scalar_out = None

    def forward(self, *, want_grad: bool, grad_scale: float = 1.0):
        dev = self.device
        scalar = self.scalar_out if self.scalar_out is not None else torch.empty(1, dtype=torch.float32, device=dev)
        grad = torch.empty(self.shape, dtype=self.dtype, device=dev) if want_grad else None
        e = self.encoded
        with torch.cuda.device(dev):
            rc = _lib.lib().pp_oks_loss_forward_encoded(
                self.params, self.enc_params, _lib.ptr(self.output), _lib.ptr(self.keypoints), _lib.ptr(self.visible),
                _lib.ptr(self.divisors), _lib.ptr(self.kp_weights), _lib.ptr(scalar), _lib.ptr(grad), float(grad_scale),
                _lib.ptr(e["keypoint_weights"]), _lib.ptr(e["in_image"]), _lib.ptr(e["annotated"]),
                _lib.ptr(self.scratch), self.scratch.numel() * 8, self.publish, _lib.stream_ptr(dev))
        _lib.check(rc, "pp_oks_loss_forward_encoded")
        return dict(loss_map=None, loss_kpt=None, scalar=scalar, peak=None, grad=grad, flag=self.last_flag)

    def backward(self, upstream: Tensor, kind: int, peak) -> Tensor:
        # only reached when the fused gradient was already consumed (a second backward): run the pass again
        grad = self.forward(want_grad=True)["grad"]
        with torch.cuda.device(self.device):
            rc = _lib.lib().pp_scale_inplace(_lib.ptr(grad), _lib.dtype_code(grad.dtype), grad.numel(), _lib.ptr(upstream),
                                             _lib.stream_ptr(self.device))
        _lib.check(rc, "pp_scale_inplace")
        return grad


def _first_element(g: Tensor) -> Tensor:
    """(1,) float32 copy of g[0, ..., 0] without materialising an expanded tensor."""
    return g.detach()[(0,) * g.ndim].reshape(1).to(torch.float32).contiguous()


_UNIT_UPSTREAM: dict = {}


def unit_upstream(device, dtype=torch.float32) -> Tensor:
    """A cached 0-dim tensor holding 1: ``loss.backward(gradient=unit_upstream(loss.device, loss.dtype))`` is
    ``loss.backward()`` without autograd's ``ones_like`` fill and -- because the fused loss recognises this tensor by its
    address -- without the launch that rescales the already-written gradient by the upstream value (two ~2-3 us kernels
    at the tail of a 100 us step).  Do not write to it."""
    key = (torch.device(device), dtype)
    t = _UNIT_UPSTREAM.get(key)
    if t is None:
        t = _UNIT_UPSTREAM[key] = torch.ones((), dtype=dtype, device=device)
    return t


def _is_unit_upstream(g: Tensor) -> bool:
    t = _UNIT_UPSTREAM.get((g.device, g.dtype))
    return t is not None and g.numel() == 1 and g.data_ptr() == t.data_ptr()


def _is_broadcast_scalar(g: Tensor) -> bool:
    return g.numel() == 1 or all(s == 0 for s, n in zip(g.stride(), g.shape) if n > 1)


class _OKSLossFunction(torch.autograd.Function):
    """autograd bridge: forward in one kernel, backward in one kernel."""

    @staticmethod
    def forward(ctx, output: Tensor, prep: _Prepared, default_mean: bool, fused: bool):
        mode = prep.params.mode
        want_fused_grad = fused and mode == _lib.PP_LOSS_PIXEL_MEAN and output.requires_grad
        res = prep.forward(want_grad=want_fused_grad)
        ctx.prep, ctx.mode, ctx.default_mean = prep, mode, default_mean
        ctx.peak = res["peak"]
        ctx.stashed = res["grad"]
        ctx.flag = res["flag"]
        prep.last_flag = res["flag"]
        if mode == _lib.PP_LOSS_PER_PIXEL:
            return res["loss_map"]
        if mode == _lib.PP_LOSS_PER_KEYPOINT and not default_mean:
            return res["loss_kpt"].to(prep.dtype)
        return res["scalar"].reshape(()).to(prep.dtype)

    @staticmethod
    def backward(ctx, g: Tensor):
        prep, mode = ctx.prep, ctx.mode
        B, K, H, W = prep.shape
        dev = prep.device
        if mode == _lib.PP_LOSS_PIXEL_MEAN:
            up = None if (ctx.stashed is not None and _is_unit_upstream(g)) else g.detach().to(torch.float32).reshape(1).contiguous()
            if ctx.stashed is not None and _is_unit_upstream(g):   # fused gradient, upstream known to be 1: nothing to do
                grad, ctx.stashed = ctx.stashed, None
            elif ctx.stashed is not None:  # fused gradient: rescale in place (no-op launch when g == 1)
                grad, ctx.stashed = ctx.stashed, None
                with torch.cuda.device(dev):
                    rc = _lib.lib().pp_scale_inplace(_lib.ptr(grad), _lib.dtype_code(grad.dtype), grad.numel(),
                                                     _lib.ptr(up), _lib.stream_ptr(dev))
                _lib.check(rc, "pp_scale_inplace")
            else:
                grad = prep.backward(up, _lib.PP_UPSTREAM_SCALAR, None)
        elif mode == _lib.PP_LOSS_PER_PIXEL:
            if _is_broadcast_scalar(g):  # e.g. the gradient of .mean() / .sum(): no (B,K,H,W) upstream in HBM
                up = _first_element(g)
                grad = prep.backward(up, _lib.PP_UPSTREAM_SCALAR, None)
            else:
                grad = prep.backward(g.detach().to(prep.dtype).contiguous(), _lib.PP_UPSTREAM_FULL, None)
        else:
            if ctx.default_mean:  # d mean / d loss_kpt = 1 / (B K)
                up = (g.detach().to(torch.float32).reshape(1) / (B * K)).contiguous()
                grad = prep.backward(up, _lib.PP_UPSTREAM_SCALAR, ctx.peak)
            elif _is_broadcast_scalar(g):
                up = _first_element(g)
                grad = prep.backward(up, _lib.PP_UPSTREAM_SCALAR, ctx.peak)
            else:
                grad = prep.backward(g.detach().to(torch.float32).contiguous(), _lib.PP_UPSTREAM_FULL, ctx.peak)
        return grad, None, None, None


class OKSHeatmapLoss(nn.Module):
    """Loss that maximises the expected OKS (ProbPose, arXiv:2412.02254); drop-in for the
    reference module (loss.py:18-191).

    Args:
        use_target_weight: kept for compatibility; as in the reference the weights apply
            whenever they are passed (loss.py:165).
        skip_empty_channel: channels whose target is all-zero do not contribute.
        smoothing_weight: weight of the Sobel smoothness term.
        gaussian_weight: weight of the MSE term.
        loss_weight: global factor.
        oks_type: ``"minus"`` (``out * (1 - tgt)``), ``"plus"`` or ``"both"``.
        check_target: replicate the reference's ``assert 0 <= target <= 1`` (loss.py:85-86).
            The range test itself is fused into the loss kernel; reading its flag costs one
            4-byte device-to-host copy and a stream synchronisation.  Set to ``False`` to
            keep the call asynchronous.
    """

    def __init__(self, use_target_weight: bool = False, skip_empty_channel: bool = False,
                 smoothing_weight: float = 0.2, gaussian_weight: float = 0.0, loss_weight: float = 1.,
                 oks_type: str = "minus", check_target: bool = True):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.skip_empty_channel = skip_empty_channel
        self.loss_weight = loss_weight
        self.smoothing_weight = smoothing_weight
        self.gaussian_weight = gaussian_weight
        self.oks_type = oks_type.lower()
        self.check_target = check_target
        assert self.oks_type in ["minus", "plus", "both"]

    def _run(self, output, target, target_weights, mask, mode, default_mean, fused, publish=None, loss_out=None):
        prep = _Prepared(self, output, target, target_weights, mask, mode)
        prep.publish = publish
        prep.scalar_out = _check_loss_out(loss_out, output)
        loss = _OKSLossFunction.apply(output, prep, default_mean, fused)
        if self.check_target:
            assert int(prep.last_flag.item()) == 0, "target should be normalized"
        return loss

    def forward(self, output: Tensor, target: Tensor, target_weights: Tensor | None = None,
                mask: Tensor | None = None, per_pixel: bool = False, per_keypoint: bool = False) -> Tensor:
        """Forward (loss.py:55-143).

        Args:
            output, target: heatmaps ``[B, K, H, W]`` (float32 or bfloat16, on the GPU).
            target_weights: ``[B, K]`` or ``[B, K, H, W]``.
            mask: ``[B, K, H, W]`` or ``[B, 1, H, W]``.
            per_pixel: return the un-reduced ``[B, K, H, W]`` loss.
            per_keypoint: return ``[B, K]``; otherwise the scalar mean of that.
        """
        if per_pixel:
            return self._run(output, target, target_weights, mask, _lib.PP_LOSS_PER_PIXEL, False, False)
        return self._run(output, target, target_weights, mask, _lib.PP_LOSS_PER_KEYPOINT, not per_keypoint, False)

    def forward_mean(self, output: Tensor, target: Tensor, target_weights: Tensor | None = None,
                     mask: Tensor | None = None, publish=None, loss_out: Tensor | None = None) -> Tensor:
        """``forward(..., per_pixel=True).mean()`` in a single fused kernel (forward + backward).  ``publish``: a
        ``PeerMailbox.descriptor(slot)`` -- the kernel that finishes the loss also stores it into that mailbox slot on
        every GPU (the loss party of the multi-GPU exchange; call ``mailbox.loss_enqueued(slot)`` afterwards).
        ``loss_out``: a caller-owned (1,) float32 CUDA tensor that receives the scalar (e.g. ``PeerMailbox.loss_slot(s)``,
        for a publication one step late); the returned loss is a view of it."""
        return self._run(output, target, target_weights, mask, _lib.PP_LOSS_PIXEL_MEAN, False, True, publish, loss_out)

    def forward_mean_encoded(self, output: Tensor, probmap, keypoints, keypoints_visible=None,
                             target_weights: Tensor | None = None, return_encoded: bool = False, publish=None,
                             loss_out: Tensor | None = None):
        """``forward_mean(output, probmap.encode_batch(keypoints, keypoints_visible)["heatmaps"], weights)`` without
        the target: ONE pass reads ``output``, forms the target of every pixel from the keypoint's separable factors
        (generate_probmaps, codec.py:56-66), reduces the loss and writes ``d loss / d output`` -- 2 H W e bytes per
        heatmap instead of 1 (encode) + 3 (loss).

        ``probmap`` is this package's ``ProbMap`` / ``ArgMaxProbMap`` (or a ``Codec`` holding one); ``keypoints``
        (B, K, D) in input-image space.  ``target_weights`` (B, K) defaults to the encoder's ``keypoint_weights``.
        With ``return_encoded`` the encoder's (B, K) outputs (``keypoint_weights``, ``in_image``, ``annotated``) are
        returned as a second value."""
        probmap = getattr(probmap, "probmap", probmap)
        prep = _PreparedEncoded(self, output, probmap, keypoints, keypoints_visible, target_weights)
        prep.publish = publish
        prep.scalar_out = _check_loss_out(loss_out, output)
        loss = _OKSLossFunction.apply(output, prep, False, True)
        return (loss, prep.encoded) if return_encoded else loss
