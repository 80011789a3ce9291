"""ctypes binding of ``csrc/libprobpose_b200.so`` (the C ABI declared in ``include/probpose_b200.h``).

There is no CPU fallback: if the library is missing, or a call fails, a
``RuntimeError`` is raised.  PyTorch is used for device memory and streams only.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libprobpose_b200.so"

PP_ABI_VERSION = 2
PP_F32, PP_BF16, PP_F64 = 0, 1, 2
PP_MAX_OKS_RADIUS = 9
PP_OKS_TAPS = 2 * PP_MAX_OKS_RADIUS + 1
PP_MAX_BLUR_KSIZE = 31
PP_LOSS_PIXEL_MEAN, PP_LOSS_PER_PIXEL, PP_LOSS_PER_KEYPOINT = 0, 1, 2
PP_UPSTREAM_SCALAR, PP_UPSTREAM_FULL = 0, 1

#: every symbol include/probpose_b200.h declares
EXPORTS = (
    "pp_version", "pp_source_hash", "pp_last_error_string", "pp_device_info", "pp_encode", "pp_decode_expected",
    "pp_oks_mma_table_bytes", "pp_oks_mma_table_build", "pp_decode_expected_scratch_bytes_for",
    "pp_decode_expected_workspace_floats", "pp_decode_expected_scratch_bytes", "pp_decode_expected_last_kernel",
    "pp_heatmap_maximum", "pp_decode_argmax_dark", "pp_blur_mma_table_build", "pp_decode_argmax_dark_last_kernel", "pp_heatmap_tail", "pp_heatmap_tail_backward",
    "pp_sparsemax_tail", "pp_sparsemax_tail_backward",
    "pp_oks_loss_scratch_bytes",
    "pp_oks_loss_forward", "pp_oks_loss_forward_encoded", "pp_oks_loss_backward", "pp_scale_inplace", "pp_pose_targets",
    "pp_pck_accuracy", "pp_binary_accuracy", "pp_masked_mae",
    "pp_mailbox_block_bytes", "pp_mailbox_bytes", "pp_mailbox_state_words", "pp_pack_records", "pp_mailbox_commit",
    "pp_mailbox_wait", "pp_mailbox_ack", "pp_mailbox_consume", "pp_mailbox_commit_deferred",
)


class EncodeParams(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("heatmap_dtype", C.c_int32), ("keypoint_dtype", C.c_int32), ("keypoint_dim", C.c_int32),
                ("scale_x", C.c_float), ("scale_y", C.c_float), ("input_w", C.c_float), ("input_h", C.c_float)]


class OksTable(C.Structure):
    _fields_ = [("radius", C.c_void_p), ("taps_f32", C.c_void_p), ("kernel2d", C.c_void_p), ("order", C.c_void_p),
                ("mma_tables", C.c_void_p), ("mma_index", C.c_void_p), ("mma_H", C.c_int32), ("mma_W", C.c_int32)]


class DecodeParams(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("heatmap_dtype", C.c_int32), ("apply_tail", C.c_int32), ("temperature", C.c_float),
                ("input_w", C.c_double), ("input_h", C.c_double)]


class LossParams(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("dtype", C.c_int32), ("mode", C.c_int32), ("oks_type", C.c_int32),
                ("skip_empty_channel", C.c_int32),
                ("smoothing_weight", C.c_double), ("gaussian_weight", C.c_double), ("loss_weight", C.c_double),
                ("mask_stride_b", C.c_int64), ("mask_stride_k", C.c_int64)]


class Mailbox(C.Structure):
    _fields_ = [("peer_bufs", C.c_void_p), ("state", C.c_void_p), ("world", C.c_int32), ("rank", C.c_int32),
                ("slots", C.c_int32), ("slot", C.c_int32), ("block_bytes", C.c_int64),
                ("flow_control", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    # The library embeds a hash of the sources it was compiled from.  A missing or stale binary is rebuilt in-tree
    # (nvcc is on every box this runs on; concurrent ranks are serialised by a file lock inside build()); if that is
    # impossible the import fails -- an edited kernel never runs against an old binary, and there is no CPU fallback.
    from .build import build, built_hash, source_hash
    have = built_hash()
    if have is None or have not in (source_hash(False), source_hash(True)):
        try:
            build()
        except Exception as e:
            raise RuntimeError(
                f"{LIB_PATH} is {'missing' if have is None else 'stale (built from other sources)'} and could not be "
                f"built ({e}): run `python -m probpose_pytorch_b200.build`.  There is no CPU fallback.") from e
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.pp_version.restype = C.c_int
    L.pp_last_error_string.restype = C.c_char_p
    L.pp_source_hash.restype = C.c_char_p
    L.pp_oks_mma_table_bytes.argtypes = [i32, i32, i32]
    L.pp_oks_mma_table_bytes.restype = i64
    L.pp_oks_mma_table_build.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.pp_decode_expected_scratch_bytes_for.argtypes = [C.POINTER(DecodeParams)]
    L.pp_decode_expected_scratch_bytes_for.restype = i64
    L.pp_device_info.argtypes = [vp, vp, vp, vp]
    L.pp_encode.argtypes = [C.POINTER(EncodeParams), vp, vp, vp, vp, vp, vp, vp, vp]
    L.pp_decode_expected.argtypes = [C.POINTER(DecodeParams), C.POINTER(OksTable), vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.pp_decode_expected_last_kernel.argtypes = []
    L.pp_decode_expected_last_kernel.restype = C.c_int
    L.pp_decode_expected_scratch_bytes.argtypes = []
    L.pp_decode_expected_scratch_bytes.restype = i64
    L.pp_decode_expected_workspace_floats.argtypes = [C.POINTER(DecodeParams)]
    L.pp_decode_expected_workspace_floats.restype = i64
    L.pp_heatmap_maximum.argtypes = [vp, i32, i64, i32, i32, vp, vp, vp, vp]
    L.pp_decode_argmax_dark.argtypes = [C.POINTER(DecodeParams), vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.pp_blur_mma_table_build.argtypes = [vp, i32, i32, i32, vp, vp]
    L.pp_decode_argmax_dark_last_kernel.argtypes = []
    L.pp_heatmap_tail.argtypes = [vp, vp, i32, i64, f32, vp]
    L.pp_heatmap_tail_backward.argtypes = [vp, vp, vp, i32, i64, f32, vp]
    L.pp_pck_accuracy.argtypes = [vp, vp, vp, vp, i32, i32, i32, C.c_double, vp, vp, vp, vp, vp]
    L.pp_binary_accuracy.argtypes = [vp, vp, vp, i64, vp, i32, vp, vp, vp]
    L.pp_masked_mae.argtypes = [vp, vp, vp, i64, vp, vp]
    L.pp_sparsemax_tail.argtypes = [vp, vp, vp, i32, i64, i64, f32, f32, vp]
    L.pp_sparsemax_tail_backward.argtypes = [vp, vp, vp, vp, i32, i64, i64, f32, f32, vp]
    L.pp_oks_loss_scratch_bytes.argtypes = [C.POINTER(LossParams)]
    L.pp_oks_loss_scratch_bytes.restype = i64
    L.pp_oks_loss_forward.argtypes = [C.POINTER(LossParams), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, i64,
                                      C.POINTER(Mailbox), vp]
    L.pp_oks_loss_forward_encoded.argtypes = [C.POINTER(LossParams), C.POINTER(EncodeParams), vp, vp, vp, vp, vp, vp, vp, f32,
                                              vp, vp, vp, vp, i64, C.POINTER(Mailbox), vp]
    L.pp_oks_loss_backward.argtypes = [C.POINTER(LossParams), vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i64, vp]
    L.pp_scale_inplace.argtypes = [vp, i32, i64, vp, vp]
    L.pp_pose_targets.argtypes = [vp, vp, vp, vp, i32, i32, C.c_double, C.c_double, vp, vp, vp, vp]
    L.pp_mailbox_block_bytes.argtypes = [i64]
    L.pp_mailbox_block_bytes.restype = i64
    L.pp_mailbox_bytes.argtypes = [i64, i32, i32]
    L.pp_mailbox_bytes.restype = i64
    L.pp_mailbox_state_words.argtypes = [i32]
    L.pp_mailbox_state_words.restype = i64
    L.pp_mailbox_ack.argtypes = [C.POINTER(Mailbox), C.c_uint32, vp]
    L.pp_mailbox_commit_deferred.argtypes = [C.POINTER(Mailbox), i64, vp, vp]
    L.pp_mailbox_consume.argtypes = [C.POINTER(Mailbox), i64, vp, vp, i64, vp, vp]
    L.pp_pack_records.argtypes = [i64, vp, vp, vp, vp, vp, vp, f32, vp, C.POINTER(Mailbox), vp]
    L.pp_mailbox_commit.argtypes = [C.POINTER(Mailbox), i64, vp, vp]
    L.pp_mailbox_wait.argtypes = [vp, i32, i32, i64, C.c_uint32, i64, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("pp_version", "pp_source_hash", "pp_last_error_string", "pp_oks_loss_scratch_bytes",
                        "pp_oks_mma_table_bytes", "pp_decode_expected_scratch_bytes_for",
                        "pp_decode_expected_workspace_floats", "pp_decode_expected_scratch_bytes",
                        "pp_mailbox_block_bytes", "pp_mailbox_bytes", "pp_mailbox_state_words"):
            fn.restype = C.c_int
    if L.pp_version() != PP_ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH}: ABI version {L.pp_version()} != {PP_ABI_VERSION}; rebuild the extension")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pp_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("probpose_pytorch_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return PP_F32
    if dt == torch.bfloat16:
        return PP_BF16
    if dt == torch.float64:
        return PP_F64
    raise TypeError(f"unsupported dtype {dt}: the heatmap path computes on float32 or bfloat16 maps")


def ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
