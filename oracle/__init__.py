"""CPU oracle for the ProbPose heatmap hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy / torch-CPU) of the reference
algorithms on the hot path named by BASELINE.json (encode, both decoders,
OKS heatmap loss forward + backward).  It exists to *check* the CUDA product
in ``probpose_pytorch_b200``; it is never the thing that is shipped or measured.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``probpose_pytorch_b200/`` imports it, and the product path has no CPU
fallback (it raises when the CUDA library is missing).

Parity pinning ("Oracle"):
  * every function cites the reference file:line it restates;
  * ``oracle/make_golden.py`` imports the *reference itself* from
    ``/root/reference`` in the build container, runs it on seeded inputs and
    commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
    checks this restatement against those fixtures on every run (bit-exact for
    encode / convolution / argmax, <=1e-6 for the floating-point tails);
  * the two known answers held by the reference's own tests
    (``tests/test_loss.py``: target peak 0.9970669150352478, zero-prediction
    loss 0.0; ``tests/test_heatmap.py``: conv backend equivalence at rtol 1e-5)
    are replayed in ``tests/test_oracle_golden.py``.

Third-party arithmetic that the reference calls but does not contain:
  * ``scipy.ndimage.convolve(mode='reflect')`` (scipy unpinned by the
    reference; 1.18.1 here)  -> restated in ``conv_reflect_f64`` (double
    accumulation in row-major tap order, float32 store);
  * ``cv2.GaussianBlur((11,11),0)`` (opencv-python 4.11.0.86 pinned; 4.13.0
    here) -> restated in ``blur_zero_pad_f32`` (separable float32 taps of
    ``getGaussianKernel(11, 0)``); agreement with cv2 is ~2e-7 relative, which
    is the precision at which cv2 itself is reproducible;
  * ``numpy.linalg.pinv`` on 2x2 symmetric matrices -> closed form;
  * ``sparsemax.Sparsemax`` (sparsemax==0.1.9 pinned, not installed, no reference test) -> published
    algorithm restated in ``sparsemax_oracle`` -- PARITY UNPINNED, see that module's header.
"""

from .codec_oracle import (  # noqa: F401
    COCO17_SIGMAS,
    oks_variance_table,
    generate_probmaps,
    encode,
    oks_kernels_2d,
    conv_reflect_f64,
    heatmap_expected_value,
    subpixel_quadratic,
    heatmap_maximum,
    gaussian_taps_f32,
    blur_zero_pad_f32,
    gaussian_blur_modulate,
    dark_udp_refine,
    decode_expected,
    decode_argmax_dark,
    head_tail,
)
from .loss_oracle import oks_heatmap_loss, oks_heatmap_loss_grad_closed_form  # noqa: F401
from . import metrics_oracle  # noqa: F401
from .sparsemax_oracle import head_tail_sparsemax, sparsemax, sparsemax_f64  # noqa: F401
from .targets_oracle import error_from_heatmaps, oks_from_heatmaps  # noqa: F401
from .probpose_loss_oracle import ProbPoseLossLayout, training_losses  # noqa: F401
