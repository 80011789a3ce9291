"""CPU: the oracle restatement against outputs of the reference itself (tests/golden/,
produced by oracle/make_golden.py) and against the reference's own known answers."""

import hashlib
import json

import numpy as np
import pytest
import torch

import oracle as oc
from probpose_pytorch_b200 import synth


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def meta(golden_dir):
    return json.loads((golden_dir / "hashes.json").read_text())


def _encode_batch(kind, wl, kps, vis):
    outs = [oc.encode(kind, wl.input_size, wl.heatmap_size, wl.sigmas, kps[b:b + 1], vis[b:b + 1])
            for b in range(kps.shape[0])]
    return (np.stack([o["heatmaps"] for o in outs]),
            np.concatenate([o["keypoint_weights"] for o in outs]),
            np.concatenate([o["in_image"] for o in outs]),
            np.concatenate([o["annotated"] for o in outs]))


@pytest.mark.parametrize("kind", ["probmap", "argmax"])
def test_encode_bit_exact_small(golden_dir, kind):
    g = np.load(golden_dir / "encode_small.npz")
    wl = synth.WORKLOADS[3]
    hm, w, inside, ann = _encode_batch(kind, wl, g["keypoints"], g["visible"])
    assert hm.dtype == np.float32
    assert np.array_equal(hm, g[f"{kind}_heatmaps"])
    assert np.array_equal(w, g[f"{kind}_weights"])
    assert np.array_equal(inside, g[f"{kind}_in_image"])
    assert np.array_equal(ann, g[f"{kind}_annotated"])
    assert (~inside).any(), "fixture must exercise out-of-image keypoints"


def test_encode_hashes_larger_shapes(meta):
    for key, want in meta["hashes"].items():
        _, cfg, b, dt, kind = key.split("/")
        wl = synth.WORKLOADS[int(cfg[1:])]
        kps, vis, _ = synth.make_keypoints(wl, batch=int(b[1:]), dtype=np.dtype(dt).type)
        hm, w, _, _ = _encode_batch(kind, wl, kps, vis)
        assert _sha(hm) == want["heatmaps_sha256"], key
        assert _sha(w.astype(np.float32)) == want["weights_sha256"], key


def test_reference_known_answers(meta):
    """tests/test_loss.py of the reference: target peak and zero-prediction loss."""
    known = meta["known"]["tests/test_loss.py"]
    enc = oc.encode("argmax", (768, 768), (192, 192), np.array([0.1] * 20),
                    np.array([[[96.0, 96.0]]]), np.array([[1.0]]))
    hm = enc["heatmaps"][None]
    assert list(hm.shape) == known["heatmap_shape"]
    assert float(hm.max()) == known["expected_target_max"] == known["target_max"]
    assert _sha(hm) == known["heatmap_sha256"]
    loss = oc.oks_heatmap_loss(torch.zeros(hm.shape), torch.from_numpy(hm), torch.tensor([[1.0]]),
                               smoothing_weight=0.05, oks_type="minus")
    assert float(loss) == known["expected_loss"] == 0.0


def test_conv_restatement_equals_scipy():
    """tests/test_heatmap.py of the reference pins the reflect-mode convolution at rtol 1e-5;
    the double-accumulation restatement is bit-identical to scipy on a seeded version of it."""
    rng = np.random.default_rng(5)
    hms = rng.random((4, 40, 36), dtype=np.float32)
    sig = rng.random(4).astype(np.float32)
    _, _, a = oc.heatmap_expected_value(hms, sig, return_heatmap=True, conv="numpy")
    _, _, b = oc.heatmap_expected_value(hms, sig, return_heatmap=True, conv="scipy")
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-8)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["blob", "uniform", "clean"])
def test_expected_decoder_bit_exact(golden_dir, name):
    g = np.load(golden_dir / "decode.npz")
    wl = synth.WORKLOADS[3]
    arr = g[name]
    for b in range(arr.shape[0]):
        locs, vals, conv = oc.heatmap_expected_value(arr[b].copy(), wl.sigmas, return_heatmap=True)
        assert np.array_equal(locs, g[f"{name}_locs"][b])
        assert np.array_equal(vals, g[f"{name}_vals"][b])
        if b == 0:
            assert np.array_equal(conv, g[f"{name}_conv0"])
            assert np.array_equal(conv.reshape(conv.shape[0], -1).argmax(1), g[f"{name}_argmax0"])
        kp, sc = oc.decode_expected(arr[b], wl.input_size, wl.heatmap_size, wl.sigmas)
        assert kp.dtype == np.float64 and np.array_equal(kp[0], g[f"{name}_keypoints"][b])
        assert np.array_equal(sc[0], g[f"{name}_scores"][b])


@pytest.mark.parametrize("name", ["blob", "clean"])
def test_dark_decoder(golden_dir, name):
    g = np.load(golden_dir / "decode.npz")
    wl = synth.WORKLOADS[3]
    arr = g[name]
    for b in range(arr.shape[0]):
        peaks, scores = oc.heatmap_maximum(arr[b])
        assert np.array_equal(peaks, g[f"{name}_peaks"][b])
        # with the library blur the restatement is the reference, bit for bit
        kp, sc = oc.decode_argmax_dark(arr[b], wl.input_size, wl.heatmap_size, backend="cv2")
        assert np.array_equal(kp[0], g[f"{name}_dark_keypoints"][b])
        assert np.array_equal(sc[0], g[f"{name}_dark_scores"][b])
        # restated blur: cv2 itself is reproducible to ~2e-7 only, and DARK amplifies that by the
        # conditioning of the 2x2 Hessian; on noise-free OKS-shaped maps it stays below 1e-5.
        kp2, _ = oc.decode_argmax_dark(arr[b], wl.input_size, wl.heatmap_size, backend="numpy")
        rel = np.abs(kp2[0] - g[f"{name}_dark_keypoints"][b]) / np.maximum(np.abs(g[f"{name}_dark_keypoints"][b]), 1.0)
        assert rel.max() < (2e-5 if name == "clean" else 5e-3), rel.max()


def test_gaussian_taps_match_cv2():
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(cv2.getGaussianKernel(11, 0, cv2.CV_32F).ravel(), oc.gaussian_taps_f32(11))


def test_loss_values_and_grads(golden_dir):
    g = np.load(golden_dir / "loss.npz")
    out, tgt = torch.from_numpy(g["output"]), torch.from_numpy(g["target"])
    tw, mask = torch.from_numpy(g["target_weights"]), torch.from_numpy(g["mask"])
    variants = {
        "train": dict(smoothing_weight=0.05, oks_type="minus"),
        "both": dict(smoothing_weight=0.2, gaussian_weight=0.1, oks_type="both", loss_weight=2.0),
        "plus_skip": dict(oks_type="plus", skip_empty_channel=True),
    }
    for vname, kw in variants.items():
        for wname, w, m in (("w", tw, None), ("wm", tw, mask), ("none", None, None)):
            for mname, mode in (("pixel", dict(per_pixel=True)), ("kpt", dict(per_keypoint=True)), ("mean", dict())):
                o = out.clone().requires_grad_(True)
                l = oc.oks_heatmap_loss(o, tgt, w, m, **mode, **kw)
                (l.mean() if mname == "pixel" else l.sum()).backward()
                key = f"{vname}/{wname}/{mname}"
                assert np.array_equal(l.detach().numpy(), g[key + "/value"]), key
                assert np.array_equal(o.grad.numpy(), g[key + "/grad"]), key


def test_closed_form_gradient_matches_autograd(golden_dir):
    g = np.load(golden_dir / "loss.npz")
    kw = dict(smoothing_weight=0.05, oks_type="minus")
    cf = oc.oks_heatmap_loss_grad_closed_form(g["output"], g["target"], g["target_weights"][:, :, None, None], **kw)
    ref = g["train/w/pixel/grad"]
    assert np.abs(cf - ref).max() <= 1e-5 * np.abs(ref).max()


def test_head_tail():
    x = np.random.default_rng(0).normal(0, 0.4, (2, 3, 8, 6)).astype(np.float32)
    want = torch.clamp(torch.from_numpy(x) / 0.5, 0, 1).numpy()
    assert np.array_equal(oc.head_tail(x, 0.5), want)


def test_pose_targets(golden_dir):
    """ProbPoseLoss._oks_from_heatmaps / _error_from_heatmaps (inputs: the decode fixture's maps)."""
    g = np.load(golden_dir / "decode.npz")
    t = np.load(golden_dir / "targets.npz")
    wl = synth.WORKLOADS[3]
    oks, w = oc.oks_from_heatmaps(g["clean"], g["blob"], t["weight"], wl.sigmas, wl.input_size, wl.heatmap_size,
                                  area_size=wl.heatmap_size, backend="cv2")
    assert oks.dtype == np.float32 and np.array_equal(oks, t["oks"]) and np.array_equal(w, t["oks_weights"])
    assert w[2] == 0 and not oks[2].any()
    err = oc.error_from_heatmaps(g["clean"], g["blob"], wl.input_size, wl.heatmap_size, backend="cv2")
    assert np.array_equal(err, t["error"])


def test_sparsemax_oracle_is_the_simplex_projection():
    """The Sparsemax restatement (parity unpinned: the package is absent) against the projection's defining
    properties and a float64 evaluation; autograd of the custom backward against finite differences."""
    import torch
    rng = np.random.default_rng(5)
    for scale in (0.05, 1.0, 10.0):
        z = (rng.standard_normal((4, 700)) * scale).astype(np.float32)
        p = oc.sparsemax(torch.from_numpy(z)).numpy()
        assert (p >= 0).all()
        np.testing.assert_allclose(p.sum(-1), 1.0, atol=2e-5)
        np.testing.assert_allclose(p, oc.sparsemax_f64(z), atol=1e-6)
        # KKT: on the support p = z - tau with one tau per row; off the support z <= tau
        for r in range(z.shape[0]):
            s = p[r] > 0
            tau = (z[r][s] - p[r][s])
            assert np.ptp(tau) < 1e-5 and (z[r][~s] <= tau.mean() + 1e-6).all()
    z = torch.from_numpy(rng.standard_normal((3, 9))).double().requires_grad_(True)
    assert torch.autograd.gradcheck(oc.sparsemax, (z,), eps=1e-7, atol=1e-5)
    y = oc.head_tail_sparsemax(torch.from_numpy(rng.standard_normal((2, 3, 4, 5)).astype(np.float32)), 0.5, 2.0)
    assert y.shape == (2, 3, 4, 5) and float(y.max()) <= 1.0 and float(y.min()) >= 0.0


def test_metrics_oracle_matches_reference_outputs(golden_dir):
    """PCK / binary accuracy / MAE restatements against the reference's outputs (tests/golden/metrics.npz)."""
    mo = oc.metrics_oracle
    m = np.load(golden_dir / "metrics.npz")
    g = np.load(golden_dir / "decode.npz")
    for thr in (0.05, 0.2):
        acc, avg, cnt = mo.pose_pck_accuracy(g["blob"], g["clean"], m["mask"], thr=thr)
        assert np.array_equal(acc, m[f"pose/{thr}/acc"]) and avg == m[f"pose/{thr}/avg"] and cnt == m[f"pose/{thr}/cnt"]
    acc, avg, cnt = mo.pose_pck_accuracy(g["blob"], g["clean"], m["mask"], thr=0.1, normalize=m["norm64"])
    assert np.array_equal(acc, m["pose/norm64/acc"]) and avg == m["pose/norm64/avg"] and cnt == m["pose/norm64/cnt"]
    acc, avg, cnt = mo.keypoint_pck_accuracy(m["pred"], m["gt"], m["mask"], 0.05, m["norm32"])
    assert np.array_equal(acc, m["kpt/acc"]) and avg == m["kpt/avg"] and cnt == m["kpt/cnt"]
    acc, avg, cnt = mo.keypoint_pck_accuracy(m["pred"], m["gt"], np.zeros_like(m["mask"]), 0.05, m["norm32"])
    assert np.array_equal(acc, m["kpt_none/acc"]) and avg == 0.0 and cnt == 0
    a, t = mo.binary_accuracy(m["scalar_dt"], m["scalar_gt"], m["mask"])
    assert a == m["binary/acc"] and t == m["binary/thr"]
    assert mo.masked_mae(m["scalar_dt"], m["scalar_gt"], m["mask"]) == m["mae"]


@pytest.mark.parametrize("name,freeze,kwargs", [("frozen", True, {}), ("live", False, {}),
                                                ("zeros", False, dict(learn_heatmaps_from_zeros=True)),
                                                ("weights", True, dict(keypoint_weights=True))])
def test_training_losses_restatement_matches_reference_forward(golden_dir, name, freeze, kwargs):
    """oracle.training_losses + oracle.ProbPoseLossLayout (the stand-in the GPU tests patch) against the reference's
    own ProbPoseLoss.forward run (tests/golden/probpose_loss.npz): losses, accuracies and gradients."""
    g = np.load(golden_dir / "probpose_loss.npz")
    wl = synth.WORKLOADS[3]

    class PM:
        input_size, heatmap_size, sigmas = wl.input_size, wl.heatmap_size, wl.sigmas

    mod = oc.ProbPoseLossLayout(PM(), freeze_error=freeze)
    gt = {k.split("/", 1)[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith("gt/")}
    pred = [torch.from_numpy(g[k]).clone().requires_grad_(True) for k in ("dt_heatmaps", "dt_probs", "dt_vis", "dt_oks", "dt_errs")]
    if kwargs.get("keypoint_weights"):
        kwargs = dict(keypoint_weights=torch.from_numpy(g["keypoint_weights"]))
    np.random.seed(99)
    losses, acc = oc.training_losses(mod, gt, tuple(pred), compute_acc=True, **kwargs)
    sum(losses.values()).backward()
    for k, v in losses.items():
        np.testing.assert_allclose(v.item(), float(g[f"{name}/loss/{k}"]), rtol=1e-6, atol=1e-9, err_msg=k)
    for k, v in acc.items():
        np.testing.assert_allclose(float(v), float(g[f"{name}/acc/{k}"]), rtol=1e-6, atol=1e-9, err_msg=k)
    for n, p_ in zip(("heatmaps", "probs", "vis", "oks", "errs"), pred):
        want = g[f"{name}/grad/{n}"]
        np.testing.assert_allclose(p_.grad.numpy(), want, rtol=1e-5, atol=1e-6 * np.abs(want).max(), err_msg=n)


def test_head_golden_is_the_reference_tail(golden_dir):
    """tests/golden/head.npz (reference ProbMapHead outputs): the tail restatement reproduces them bit for bit."""
    g = np.load(golden_dir / "head.npz")
    assert np.array_equal(oc.head_tail(g["pre_tail"], float(g["temperature"])), g["heatmaps"])
    frac = (g["heatmaps"] == 0).mean(), (g["heatmaps"] == 1).mean()
    assert 0.2 < frac[0] < 0.6 and 0.05 < frac[1] < 0.4      # all three regimes of the clamp are populated
