"""probpose_pytorch_b200 -- B200-native heatmap hot path of ProbPose.

Drop-in, same call signatures as zir-vision/ProbPose_pytorch for:
  * ``codec``   : ``generate_probmaps``, ``ProbMap``, ``ArgMaxProbMap``, ``Codec``
  * ``heatmap`` : ``get_heatmap_maximum``, ``get_heatmap_expected_value``
  * ``loss``    : ``OKSHeatmapLoss``; ``probpose_loss``: ``patch_probpose_loss`` (device-side members for the
    reference's own ``ProbPoseLoss`` instance), ``ground_truth_from_keypoints``
  * ``metrics`` : ``pose_pck_accuracy``, ``keypoint_pck_accuracy``
  * ``head``    : ``heatmap_tail`` (tail of ``ProbMapHead.forward_heatmap``), ``Sparsemax``

Everything computes in hand-written sm_100a CUDA behind the C ABI of
``include/probpose_b200.h`` (``csrc/libprobpose_b200.so``); there is no CPU
fallback -- calls raise when the library or a GPU is missing.
"""

from .codec import ArgMaxProbMap, Codec, ProbMap, generate_probmaps  # noqa: F401
from .head import HeatmapTail, Sparsemax, heatmap_tail, patch_probmap_head  # noqa: F401
from .heatmap import get_heatmap_expected_value, get_heatmap_maximum  # noqa: F401
from .loss import OKSHeatmapLoss, unit_upstream  # noqa: F401
from .probpose_loss import FusedOKSHeatmapLoss, ground_truth_from_keypoints, patch_probpose_loss  # noqa: F401

__all__ = ["ArgMaxProbMap", "Codec", "ProbMap", "generate_probmaps", "heatmap_tail", "HeatmapTail",
           "patch_probmap_head", "Sparsemax",
           "get_heatmap_expected_value", "get_heatmap_maximum", "OKSHeatmapLoss",
           "FusedOKSHeatmapLoss", "patch_probpose_loss", "ground_truth_from_keypoints", "unit_upstream"]
