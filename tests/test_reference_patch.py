"""Build container only (needs /root/reference): the patch functions applied to the REFERENCE's own objects.

``patch_probpose_loss`` and ``patch_probmap_head`` exist so that a user keeps the reference's ``ProbPoseLoss`` /
``ProbMapHead`` and only the hot-path members change.  The GPU box has no reference, so the numerics of the patched
members are tested there on a stand-in with the same member names (oracle.ProbPoseLossLayout, pinned against the
reference by tests/test_oracle_golden.py); here the real objects are patched and the contract between the
reference's ``forward`` and the patched members is checked on its source."""

import inspect
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "probpose" / "loss.py").exists(), reason="the reference is only present in the build container")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, str(REF))
    try:
        from probpose import codec, head, loss
    finally:
        sys.path.remove(str(REF))
    return dict(codec=codec, head=head, loss=loss)


def test_patch_reference_probpose_loss_instance(ref):
    import probpose_pytorch_b200 as pp
    from probpose_pytorch_b200 import synth
    from probpose_pytorch_b200.probpose_loss import FusedOKSHeatmapLoss, LazyPerPixelLoss
    wl = synth.WORKLOADS[3]
    rc, rl = ref["codec"], ref["loss"]
    mod = rl.ProbPoseLoss(rc.Codec(rc.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)), freeze_error=False)
    before = mod.keypoint_loss_module
    scalar_modules = [mod.probability_loss_module, mod.visibility_loss_module, mod.oks_loss_module, mod.error_loss_module]
    assert pp.patch_probpose_loss(mod) is mod
    new = mod.keypoint_loss_module
    assert isinstance(new, FusedOKSHeatmapLoss) and new in list(mod.modules())
    for attr in ("use_target_weight", "skip_empty_channel", "smoothing_weight", "gaussian_weight", "loss_weight", "oks_type"):
        assert getattr(new, attr) == getattr(before, attr), attr
    # the scalar-head losses and forward stay the reference's own
    assert [mod.probability_loss_module, mod.visibility_loss_module, mod.oks_loss_module, mod.error_loss_module] == scalar_modules
    assert type(mod).forward is rl.ProbPoseLoss.forward
    for name in ("_oks_from_heatmaps", "_error_from_heatmaps", "get_pose_accuracy", "get_binary_accuracy", "get_mae"):
        assert name in vars(mod), f"{name} not replaced on the instance"
    dc = mod.device_codec.probmap
    assert tuple(dc.input_size) == tuple(wl.input_size) and tuple(dc.heatmap_size) == tuple(wl.heatmap_size)
    assert np.array_equal(dc.sigmas, wl.sigmas) and dc.sigma == mod.codec.probmap.sigma
    # the contract the patched members rely on, read off the reference's forward:
    src = inspect.getsource(rl.ProbPoseLoss.forward)
    assert "per_pixel=True" in src and "heatmap_loss_pxl.mean()" in src            # -> LazyPerPixelLoss.mean()
    assert "torch.from_numpy(gt_errs)" in src                                      # -> _error_from_heatmaps returns NumPy
    assert "self._oks_from_heatmaps(" in src and "heatmap_size=(W, H)" in src
    for name in ("get_pose_accuracy", "get_binary_accuracy", "get_mae"):
        assert f"self.{name}(" in src
    # the lazy result: .mean() is the fused kernel's, everything else materialises (CUDA needed to run either)
    lazy = new(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), torch.ones(1, 1), per_pixel=True)
    assert isinstance(lazy, LazyPerPixelLoss) and tuple(lazy.shape) == (1, 1, 4, 4)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            lazy.mean()


def test_patch_reference_probmap_head_instance(ref):
    import probpose_pytorch_b200 as pp
    torch.manual_seed(0)
    head = ref["head"].ProbMapHead(32, 5, [(4, 3), (2, 2), (2, 2)], (16,), (4,))
    orig = type(head).forward_heatmap
    assert pp.patch_probmap_head(head) is head
    assert "forward_heatmap" in vars(head) and type(head).forward_heatmap is orig
    # forward() reaches the heatmaps through self.forward_heatmap, so the patched tail is what the 5-tuple carries
    assert "self.forward_heatmap(" in inspect.getsource(type(head).forward)
    src = inspect.getsource(orig)
    assert "x / self.temperature" in src and "torch.clamp(x, 0, 1)" in src
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            head.forward_heatmap(torch.randn(1, 32, 4, 3))
