"""GPU parity: the CUDA path (through the Python shims -> ctypes -> C ABI) against the oracle on the
same seeded inputs and against the committed reference outputs (tests/golden).

Tolerances are the ones BASELINE.json's north_star states: argmax indices bit-exact; encoded targets,
decoded coordinates, scores and loss values within 1e-5 relative in float32 and 1e-2 in bfloat16.
"""

import numpy as np
import pytest
import torch

import oracle as oc
from probpose_pytorch_b200 import synth

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5
RTOL16 = 1e-2


@pytest.fixture(scope="module")
def pp():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import probpose_pytorch_b200 as mod
    from probpose_pytorch_b200 import _lib
    _lib.lib()  # fail loudly if the extension is missing
    torch.cuda.set_device(0)
    return mod



@pytest.fixture(params=["auto", "mma-grid2", "nomma", "team1", "team2", "team1-fulltaps", "team1-grid3", "cta", "cta-grid5"])
def expected_kernel(request, monkeypatch):
    """Which kernel runs the expected-OKS decode: the library's rule ("auto": the tensor-core kernel pp_decode_mma.cuh
    for the 64x48 / 96x72 shapes it is built for, else the size rule of the general kernels = "nomma"), the
    team-per-heatmap kernel (pp_decode_warp.cuh) with one or two warps per heatmap, or the CTA-per-heatmap kernel
    (pp_decode_fast.cuh) forced.  "-gridN" caps the launch at N CTAs, so that every warp / CTA decodes MANY heatmaps:
    the work-queue pull, the single-slot TMA refill and the mbarrier phase flips of the persistent loops run against
    the oracle."""
    mode = request.param
    if "-grid" in mode:
        monkeypatch.setenv("PP_DECODE_GRID", mode.split("-grid")[1])
        mode = mode.split("-grid")[0]
    if mode == "nomma":
        monkeypatch.setenv("PP_DECODE_MMA", "0")
    elif mode == "cta":
        monkeypatch.setenv("PP_DECODE_WARP", "0")
    elif mode.startswith("team"):
        monkeypatch.setenv("PP_DECODE_WARP", "1")
        monkeypatch.setenv("PP_DECODE_TEAM", mode[4])
        if mode.endswith("fulltaps"):   # prefilter with every tap instead of the truncated wide kernels
            monkeypatch.setenv("PP_DECODE_FULLTAPS", "1")
    return request.param


def _oracle_encode(kind, wl, kps, vis, sigma=None):
    outs = [oc.encode(kind, wl.input_size, wl.heatmap_size, wl.sigmas, kps[b:b + 1], vis[b:b + 1], sigma=sigma)
            for b in range(kps.shape[0])]
    return {k: (np.stack([o[k] for o in outs]) if k == "heatmaps" else np.concatenate([o[k] for o in outs]))
            for k in ("heatmaps", "keypoint_weights", "in_image", "annotated")}


def _assert_maps_close(got, want, rtol):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    # relative on normal numbers; float32 subnormals (< 1.2e-38) carry fewer bits, so give them an absolute floor
    err = np.abs(got - want)
    assert (err <= rtol * np.abs(want) + 1e-37).all(), f"max rel err {np.max(err / np.maximum(np.abs(want), 1e-30))}"


# --------------------------------------------------------------------------- encode
@pytest.mark.parametrize("kind", ["probmap", "argmax"])
def test_encode_matches_golden_reference(pp, golden_dir, kind):
    g = np.load(golden_dir / "encode_small.npz")
    wl = synth.WORKLOADS[3]
    cls = pp.ProbMap if kind == "probmap" else pp.ArgMaxProbMap
    codec = cls(wl.input_size, wl.heatmap_size, wl.sigmas)
    for b in range(g["keypoints"].shape[0]):
        enc = codec.encode(g["keypoints"][b:b + 1], g["visible"][b:b + 1])
        assert enc["heatmaps"].dtype == np.float32 and enc["heatmaps"].shape == g[f"{kind}_heatmaps"][b].shape
        _assert_maps_close(enc["heatmaps"], g[f"{kind}_heatmaps"][b], RTOL32)
        mism = np.count_nonzero(enc["heatmaps"] != g[f"{kind}_heatmaps"][b])
        assert mism <= 2, f"{mism} float32 elements differ from the reference bit pattern"
        assert np.array_equal(enc["keypoint_weights"], g[f"{kind}_weights"][b:b + 1])
        assert np.array_equal(enc["in_image"], g[f"{kind}_in_image"][b:b + 1])
        assert np.array_equal(enc["annotated"], g[f"{kind}_annotated"][b:b + 1])
        if kind == "probmap":
            assert np.array_equal(enc["heatmap_keypoints"], g["keypoints"][b:b + 1] / codec.scale_factor)


@pytest.mark.parametrize("cid,batch", [(1, 8), (3, 8), (4, 4), (5, 2)])
@pytest.mark.parametrize("kpdtype", [np.float32, np.float64])
def test_encode_batch_matches_oracle(pp, cid, batch, kpdtype):
    wl = synth.WORKLOADS[cid]
    kps, vis, _ = synth.make_keypoints(wl, batch=batch, dtype=kpdtype)
    for kind, cls in (("probmap", pp.ProbMap), ("argmax", pp.ArgMaxProbMap)):
        codec = cls(wl.input_size, wl.heatmap_size, wl.sigmas)
        out = codec.encode_batch(kps, vis)
        want = _oracle_encode(kind, wl, kps, vis)
        _assert_maps_close(out["heatmaps"].cpu().numpy(), want["heatmaps"], RTOL32)
        assert np.array_equal(out["keypoint_weights"].cpu().numpy(), want["keypoint_weights"].astype(np.float32))
        assert np.array_equal(out["in_image"].cpu().numpy(), want["in_image"])
        assert np.array_equal(out["annotated"].cpu().numpy(), want["annotated"])


def test_encode_edge_cases(pp):
    wl = synth.WORKLOADS[1]
    K = wl.num_keypoints
    codec = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    kps = np.zeros((1, K, 2), dtype=np.float64)
    kps[0, 0] = (-500.0, -500.0)      # far outside: map underflows, weight 0
    kps[0, 1] = (191.0, 255.0)        # last pixel
    kps[0, 2] = (192.0, 10.0)         # just outside in x
    kps[0, 3] = (95.5, 127.5)         # half-grid
    vis = np.ones((1, K), dtype=np.float32)
    vis[0, 4] = 0.0                   # unlabelled: zero channel, weight copied
    vis[0, 5] = 0.4
    got = codec.encode(kps, vis)
    want = oc.encode("argmax", wl.input_size, wl.heatmap_size, wl.sigmas, kps, vis)
    _assert_maps_close(got["heatmaps"], want["heatmaps"], RTOL32)
    for key in ("keypoint_weights", "in_image", "annotated"):
        assert np.array_equal(got[key], want[key]), key
    assert got["keypoint_weights"][0, 0] == 0 and not got["heatmaps"][0].any()
    assert not got["heatmaps"][4].any() and got["keypoint_weights"][0, 5] == np.float32(0.4)
    # bool visibility keeps a bool weight array (dataset.py:124)
    got_b = codec.encode(kps, vis > 0.5)
    assert got_b["keypoint_weights"].dtype == bool
    # default visible = ones; generate_probmaps with heatmap-space keypoints and the scalar sigma override
    hm, w = pp.generate_probmaps((48, 64), (kps / codec.scale_factor), vis, wl.sigmas, sigma=0.55)
    hm_ref, w_ref = oc.generate_probmaps((48, 64), (kps / codec.scale_factor), vis, wl.sigmas, sigma=0.55)
    _assert_maps_close(hm, hm_ref, RTOL32)
    assert np.array_equal(w, w_ref)
    with pytest.raises(AssertionError):
        codec.encode(np.zeros((2, K, 2)))


def test_encode_odd_width_and_bf16(pp):
    wl = synth.Workload(9, "odd", 3, 5, (101, 75), (27, 19), True)
    kps, vis, _ = synth.make_keypoints(wl, seed=5)
    codec = pp.ProbMap(wl.input_size, wl.heatmap_size, np.full(5, 0.05), sigma=-1)
    out = codec.encode_batch(kps, vis)
    want = np.stack([oc.encode("probmap", wl.input_size, wl.heatmap_size, np.full(5, 0.05), kps[b:b + 1], vis[b:b + 1],
                               sigma=-1)["heatmaps"] for b in range(3)])
    _assert_maps_close(out["heatmaps"].cpu().numpy(), want, RTOL32)
    wl = synth.WORKLOADS[3]
    kps, vis, _ = synth.make_keypoints(wl, batch=4)
    codec = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    out16 = codec.encode_batch(kps, vis, dtype=torch.bfloat16)["heatmaps"]
    want = torch.from_numpy(_oracle_encode("argmax", wl, kps, vis)["heatmaps"]).bfloat16()
    assert out16.dtype == torch.bfloat16
    assert torch.equal(out16.cpu(), want) or (out16.cpu().float() - want.float()).abs().max() <= RTOL16


# --------------------------------------------------------------------------- argmax
def test_heatmap_maximum_exact(pp):
    rng = np.random.default_rng(11)
    hm = rng.random((3, 17, 64, 48), dtype=np.float32)
    hm[0, 0] = 0.0                       # empty channel -> (-1, -1)
    hm[0, 1] = -1.0
    hm[1, 2, 10, 7] = hm[1, 2, 50, 3] = 2.0   # tie -> lowest flat index
    for arr in (hm, hm[0]):
        locs, vals = pp.get_heatmap_maximum(arr)
        locs_ref, vals_ref = oc.heatmap_maximum(arr)
        assert locs.dtype == np.float32 and np.array_equal(locs, locs_ref) and np.array_equal(vals, vals_ref)
    odd = rng.random((2, 3, 19, 27), dtype=np.float32)
    assert np.array_equal(pp.get_heatmap_maximum(odd)[0], oc.heatmap_maximum(odd)[0])
    t = torch.from_numpy(hm).cuda()
    locs_t, vals_t = pp.get_heatmap_maximum(t)
    assert locs_t.is_cuda and np.array_equal(locs_t.cpu().numpy(), oc.heatmap_maximum(hm)[0])
    with pytest.raises(AssertionError):
        pp.get_heatmap_maximum(np.zeros((4, 4), dtype=np.float32))


# --------------------------------------------------------------------------- expected-OKS decoder
@pytest.mark.parametrize("name", ["blob", "uniform", "clean"])
def test_expected_decoder_matches_golden_reference(pp, golden_dir, name, expected_kernel):
    g = np.load(golden_dir / "decode.npz")
    wl = synth.WORKLOADS[3]
    arr = g[name]
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    # batched device call: argmax bit-exact, coordinates/scores within 1e-5
    dev = pm.decode_device(torch.from_numpy(arr).cuda())
    locs = dev["locs"].cpu().numpy()
    np.testing.assert_allclose(locs, g[f"{name}_locs"], rtol=RTOL32, atol=1e-5)
    assert np.array_equal(dev["vals"].cpu().numpy(), g[f"{name}_vals"])
    np.testing.assert_allclose(dev["keypoints"].cpu().numpy(), g[f"{name}_keypoints"], rtol=RTOL32, atol=1e-5)
    assert np.array_equal(dev["argmax"][0].cpu().numpy(), g[f"{name}_argmax0"])
    # reference-shaped single-sample calls
    for b in range(arr.shape[0]):
        l, v = pp.get_heatmap_expected_value(arr[b], wl.sigmas)
        assert l.shape == (17, 2) and l.dtype == np.float32 and v.shape == (17,)
        np.testing.assert_allclose(l, g[f"{name}_locs"][b], rtol=RTOL32, atol=1e-5)
        kp, sc = pm.decode(arr[b])
        assert kp.shape == (1, 17, 2) and kp.dtype == np.float64 and sc.shape == (1, 17)
        np.testing.assert_allclose(kp[0], g[f"{name}_keypoints"][b], rtol=RTOL32, atol=1e-5)
        assert np.array_equal(sc[0], g[f"{name}_scores"][b])
    # return_heatmap: the convolved map itself, float32, bit for bit (double accumulation in scipy's order)
    l, v, conv = pp.get_heatmap_expected_value(arr[0], wl.sigmas, return_heatmap=True)
    assert np.array_equal(conv, g[f"{name}_conv0"])
    np.testing.assert_allclose(l, g[f"{name}_locs"][0], rtol=RTOL32, atol=1e-5)


@pytest.mark.parametrize("cid,batch", [(2, 6), (4, 3), (5, 1)])
def test_expected_decoder_argmax_exact_vs_oracle(pp, cid, batch, expected_kernel):
    wl = synth.WORKLOADS[cid]
    kps, vis, _ = synth.make_keypoints(wl, batch=batch, seed=200 + cid)
    tgt = _oracle_encode("argmax", wl, synth.jitter_keypoints(wl, kps, 201), vis)["heatmaps"]
    pred = synth.blob_predictions_numpy(tgt, synth.blob_params(tgt.shape[:2], 202), 203)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = pm.decode_device(torch.from_numpy(pred).cuda())
    W = wl.heatmap_size[0]
    for b in range(batch):
        locs, vals, conv = oc.heatmap_expected_value(pred[b], wl.sigmas, return_heatmap=True, conv="scipy")
        am = conv.reshape(conv.shape[0], -1).argmax(1)
        assert np.array_equal(dev["argmax"][b].cpu().numpy(), am), f"sample {b}"
        np.testing.assert_allclose(dev["locs"][b].cpu().numpy(), locs, rtol=RTOL32, atol=1e-5)
        assert np.array_equal(dev["vals"][b].cpu().numpy(), vals)


def test_expected_decoder_ties_plateaus_and_constants(pp, expected_kernel):
    """On-grid / half-grid targets (exact ties), saturated plateaus, all-zero and constant maps."""
    wl = synth.WORKLOADS[1]
    K, (W, H) = wl.num_keypoints, wl.heatmap_size
    rng = np.random.default_rng(31)
    grid = (rng.integers(0, [W - 1, H - 1], size=(4, K, 2)) + rng.choice([0.0, 0.5], size=(4, K, 2)))
    kps = (grid * ((np.array(wl.input_size) - 1) / (np.array(wl.heatmap_size) - 1))).astype(np.float64)
    vis = np.ones((4, K), dtype=np.float32)
    maps = _oracle_encode("argmax", wl, kps, vis)["heatmaps"]
    maps[1] = np.clip(maps[1] * 3.0, 0, 1)        # plateaus of exactly 1.0
    maps[2, 0] = 0.0                               # all-zero channel -> (0, 0), no -1 sentinel (B-5)
    maps[2, 1] = 0.25                              # constant channel
    maps[3] = np.round(maps[3] * 4) / 4            # heavy quantisation -> many exact ties
    maps[0, 2] = 0.0
    maps[0, 2, 20:40, 10:30] = 0.5               # a plateau wider than the kernel: hundreds of exactly tied maxima
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = pm.decode_device(torch.from_numpy(maps).cuda())
    for b in range(4):
        locs, vals, conv = oc.heatmap_expected_value(maps[b], wl.sigmas, return_heatmap=True, conv="scipy")
        assert np.array_equal(dev["argmax"][b].cpu().numpy(), conv.reshape(K, -1).argmax(1)), f"sample {b}"
        np.testing.assert_allclose(dev["locs"][b].cpu().numpy(), locs, rtol=RTOL32, atol=1e-5)
        assert np.array_equal(dev["vals"][b].cpu().numpy(), vals)
    assert dev["locs"][2, 0].tolist() == [0.0, 0.0]


def test_expected_decoder_reference_test_shape_and_generic_path(pp):
    """Seeded version of the reference's tests/test_heatmap.py (20 x 256 x 256 uniform maps, random
    sigmas): maps too large for shared memory take the exact full-map path."""
    rng = np.random.default_rng(5)
    hms = rng.random((3, 256, 256), dtype=np.float32)
    sig = rng.random(3).astype(np.float32)
    l, v, conv = pp.get_heatmap_expected_value(hms, sig, return_heatmap=True, backend="torch")
    l_ref, v_ref, conv_ref = oc.heatmap_expected_value(hms, sig, return_heatmap=True, conv="scipy")
    np.testing.assert_allclose(conv, conv_ref, rtol=1e-5, atol=1e-8)   # the reference test's own tolerance
    assert np.array_equal(conv, conv_ref)
    np.testing.assert_allclose(l, l_ref, rtol=RTOL32, atol=1e-5)
    assert np.array_equal(v, v_ref)
    l2, v2 = pp.get_heatmap_expected_value(hms, sig)                   # no conv requested: work space is internal
    assert np.array_equal(l2, l) and np.array_equal(v2, v)
    # odd sizes
    odd = rng.random((5, 19, 27), dtype=np.float32)
    l, v = pp.get_heatmap_expected_value(odd, np.full(5, 0.07))
    l_ref, v_ref = oc.heatmap_expected_value(odd, np.full(5, 0.07), conv="scipy")
    np.testing.assert_allclose(l, l_ref, rtol=RTOL32, atol=1e-5)
    assert np.array_equal(v, v_ref)


def test_decoder_fused_head_tail(pp, expected_kernel):
    wl = synth.WORKLOADS[2]
    rng = np.random.default_rng(41)
    logits = rng.normal(0.1, 0.3, size=(4, 17, 64, 48)).astype(np.float32)
    x = torch.from_numpy(logits).cuda()
    tail = pp.heatmap_tail(x, 0.5)
    assert np.array_equal(tail.cpu().numpy(), oc.head_tail(logits, 0.5))
    assert torch.equal(tail, torch.clamp(x / 0.5, 0, 1))
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    a = pm.decode_device(x, temperature=0.5)
    b = pm.decode_device(tail)
    for key in ("locs", "vals", "argmax", "keypoints"):
        assert torch.equal(a[key], b[key]), key
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    a, b = am.decode_device(x, temperature=0.5), am.decode_device(tail)
    for key in ("peaks", "scores", "locs", "keypoints"):
        assert torch.equal(a[key], b[key]), key


# --------------------------------------------------------------------------- DARK decoder
@pytest.fixture(params=["mma", "mma-grid2", "cta"])
def dark_kernel(request, monkeypatch):
    """Which kernel runs the argmax + DARK-UDP decode: the tensor-core kernel (pp_decode_mma.cuh, kDark; default for the
    64x48 / 96x72 shapes; "-grid2": two CTAs, many heatmaps per warp) or the CTA-per-heatmap kernel (pp_dark_fast.cuh)."""
    if request.param == "cta":
        monkeypatch.setenv("PP_DARK_MMA", "0")
    elif request.param.endswith("grid2"):
        monkeypatch.setenv("PP_DECODE_GRID", "2")
    return request.param


def _dark_tolerance(hm, peaks, wl, base_rtol):
    """Per-keypoint tolerance (input px).  cv2 / numpy float32 blurs agree to ~2e-7 relative and the
    float32 log to ~1 ulp; DARK multiplies that by the inverse Hessian of the log-map at the peak."""
    K, H, W = hm.shape
    work = oc.gaussian_blur_modulate(hm.copy(), 11)
    np.clip(work, 1e-3, 50.0, work)
    lg = np.log(work.astype(np.float64))
    pad = np.pad(lg, ((0, 0), (1, 1), (1, 1)), mode="edge")
    tol = np.zeros((K, 2))
    scale = np.array(wl.input_size, dtype=np.float64) / (np.array(wl.heatmap_size) - 1)
    for k in range(K):
        x, y = int(peaks[k, 0]) + 1, int(peaks[k, 1]) + 1
        if x <= 0:
            continue
        c = pad[k, y, x]
        dxx = pad[k, y, x + 1] - 2 * c + pad[k, y, x - 1]
        dyy = pad[k, y + 1, x] - 2 * c + pad[k, y - 1, x]
        dxy = 0.5 * (pad[k, y + 1, x + 1] - pad[k, y, x + 1] - pad[k, y + 1, x] + 2 * c - pad[k, y, x - 1] - pad[k, y - 1, x]
                     + pad[k, y - 1, x - 1])
        hess = np.array([[dxx, dxy], [dxy, dyy]]) + np.finfo(np.float32).eps * np.eye(2)
        inv = np.linalg.pinv(hess)
        g = np.array([0.5 * (pad[k, y, x + 1] - pad[k, y, x - 1]), 0.5 * (pad[k, y + 1, x] - pad[k, y - 1, x])])
        shift = np.abs(inv @ g).max()
        delta = 2e-6  # absolute noise of the float32 log-values (|log| <= 6.9, a few ulp)
        amp = np.abs(inv).sum(axis=1).max()
        tol[k] = amp * delta * (1.0 + 4.0 * shift)
    return tol * scale


@pytest.mark.parametrize("name", ["clean", "blob"])
def test_dark_decoder_matches_golden_reference(pp, golden_dir, name, dark_kernel):
    from probpose_pytorch_b200 import _lib
    g = np.load(golden_dir / "decode.npz")
    wl = synth.WORKLOADS[3]
    arr = g[name]
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = am.decode_device(torch.from_numpy(arr).cuda())
    assert _lib.lib().pp_decode_argmax_dark_last_kernel() == (1 if dark_kernel == "cta" else 5)
    assert np.array_equal(dev["peaks"].cpu().numpy(), g[f"{name}_peaks"])           # argmax bit-exact
    assert np.array_equal(dev["scores"].cpu().numpy(), g[f"{name}_dark_scores"])
    got = dev["keypoints"].cpu().numpy()
    want = g[f"{name}_dark_keypoints"]
    n_strict = n_total = 0
    for b in range(arr.shape[0]):
        live = g[f"{name}_peaks"][b, :, 0] >= 0
        # DARK is only meaningful (and only reproducible: cv2 vs any other float32 blur differ by 1e-2 px there)
        # on blob-shaped maps; channels that hold nothing but the U(0, 0.02) noise floor are checked for
        # argmax / score parity above and for finiteness here ("parity inputs must be blob-shaped")
        blobby = live & (arr[b].reshape(arr.shape[1], -1).max(axis=1) >= 0.1)
        assert np.isfinite(got[b][live]).all()
        tol = _dark_tolerance(arr[b], g[f"{name}_peaks"][b], wl, RTOL32)
        err = np.abs(got[b] - want[b])
        bound = RTOL32 * np.maximum(np.abs(want[b]), 1.0) + tol
        assert (err[blobby] <= bound[blobby]).all(), (b, np.max(err[blobby] / bound[blobby]))
        strict = err[blobby] <= RTOL32 * np.maximum(np.abs(want[b][blobby]), 1.0)
        n_strict += strict.sum(); n_total += strict.size
        # empty channels keep the sentinel (documented deviation from the reference's out-of-range reads)
        assert (dev["locs"][b].cpu().numpy()[~live] == -1).all()
        kp, sc = am.decode(arr[b])      # reference-shaped call
        assert kp.shape == (1, 17, 2) and kp.dtype == np.float64
        assert np.array_equal(kp[0], got[b]) and np.array_equal(sc[0], g[f"{name}_dark_scores"][b])
    if name == "clean":      # noise-free OKS-shaped maps: plain 1e-5 everywhere
        assert n_strict == n_total, f"{n_total - n_strict} of {n_total} coordinates outside 1e-5"
    else:
        assert n_strict >= 0.9 * n_total


@pytest.mark.parametrize("cid,batch", [(4, 2), (5, 1)])
def test_dark_decoder_vs_oracle_other_shapes(pp, cid, batch, dark_kernel):
    wl = synth.WORKLOADS[cid]
    kps, vis, _ = synth.make_keypoints(wl, batch=batch, seed=300 + cid)
    inside = (kps[..., 0] >= 8) & (kps[..., 0] < wl.input_size[0] - 8) & (kps[..., 1] >= 8) & (kps[..., 1] < wl.input_size[1] - 8)
    maps = _oracle_encode("argmax", wl, kps, np.ones_like(vis))["heatmaps"]
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = am.decode_device(torch.from_numpy(maps).cuda())
    for b in range(batch):
        kp, sc = oc.decode_argmax_dark(maps[b], wl.input_size, wl.heatmap_size, backend="cv2")
        peaks, _ = oc.heatmap_maximum(maps[b])
        assert np.array_equal(dev["peaks"][b].cpu().numpy(), peaks)
        assert np.array_equal(dev["scores"][b].cpu().numpy(), sc[0])
        ok = inside[b] & (peaks[:, 0] >= 0)
        got = dev["keypoints"][b].cpu().numpy()
        tol = _dark_tolerance(maps[b], peaks, wl, RTOL32)
        bound = RTOL32 * np.maximum(np.abs(kp[0]), 1.0) + tol
        assert (np.abs(got - kp[0])[ok] <= bound[ok]).all()


# --------------------------------------------------------------------------- loss
_VARIANTS = {
    "train": dict(smoothing_weight=0.05, oks_type="minus"),
    "both": dict(smoothing_weight=0.2, gaussian_weight=0.1, oks_type="both", loss_weight=2.0),
    "plus_skip": dict(oks_type="plus", skip_empty_channel=True),
}


def _close(got, want, rtol, scale=None):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max() if scale is None else scale
    err = np.abs(got - want)
    assert (err <= rtol * np.abs(want) + rtol * 0.05 * scale + 1e-30).all(), \
        f"max err {err.max():.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("vname", list(_VARIANTS))
def test_loss_matches_golden_reference(pp, golden_dir, vname):
    g = np.load(golden_dir / "loss.npz")
    out = torch.from_numpy(g["output"]).cuda()
    tgt = torch.from_numpy(g["target"]).cuda()
    tw = torch.from_numpy(g["target_weights"]).cuda()
    mask = torch.from_numpy(g["mask"]).cuda()
    mod = pp.OKSHeatmapLoss(use_target_weight=True, **_VARIANTS[vname])
    for wname, w, m in (("w", tw, None), ("wm", tw, mask), ("none", None, None)):
        for mname, mode in (("pixel", dict(per_pixel=True)), ("kpt", dict(per_keypoint=True)), ("mean", dict())):
            o = out.clone().requires_grad_(True)
            l = mod(o, tgt, w, m, **mode)
            (l.mean() if mname == "pixel" else l.sum()).backward()
            key = f"{vname}/{wname}/{mname}"
            want_v, want_g = g[key + "/value"], g[key + "/grad"]
            assert tuple(l.shape) == want_v.shape, key
            _close(l.detach().cpu().numpy(), want_v, RTOL32)
            _close(o.grad.cpu().numpy(), want_g, RTOL32)
        # fused mean == per_pixel.mean(), values and gradients
        o = out.clone().requires_grad_(True)
        l = mod.forward_mean(o, tgt, w, m)
        (l * 3.0).backward()
        key = f"{vname}/{wname}/pixel"
        assert abs(l.item() - g[key + "/value"].mean(dtype=np.float64)) <= RTOL32 * abs(g[key + "/value"].mean(dtype=np.float64)) + 1e-12
        _close(o.grad.cpu().numpy(), 3.0 * g[key + "/grad"], RTOL32)


def test_loss_full_upstream_and_pixel_weights(pp):
    torch.manual_seed(3)
    B, K, H, W = 3, 4, 40, 36
    out, tgt = torch.rand(B, K, H, W), torch.rand(B, K, H, W)
    twp = torch.rand(B, K, H, W)
    up_px, up_k = torch.rand(B, K, H, W), torch.rand(B, K)
    kw = dict(smoothing_weight=0.1, gaussian_weight=0.2, oks_type="both")
    mod = pp.OKSHeatmapLoss(**kw)
    for mode, up in ((dict(per_pixel=True), up_px), (dict(per_keypoint=True), up_k)):
        o_ref = out.clone().requires_grad_(True)
        l_ref = oc.oks_heatmap_loss(o_ref, tgt, twp, None, **mode, **kw)
        (l_ref * up).sum().backward()
        o = out.cuda().requires_grad_(True)
        l = mod(o, tgt.cuda(), twp.cuda(), None, **mode)
        (l * up.cuda()).sum().backward()
        _close(l.detach().cpu().numpy(), l_ref.detach().numpy(), RTOL32)
        _close(o.grad.cpu().numpy(), o_ref.grad.numpy(), RTOL32)


def test_loss_known_answer_and_asserts(pp):
    """tests/test_loss.py of the reference: 192x192 target, zero prediction -> loss 0.0; target range assert."""
    codec = pp.ArgMaxProbMap((768, 768), (192, 192), np.array([0.1] * 20))
    enc = codec.encode(np.array([[[96.0, 96.0]]]), np.array([[1.0]]), np.array([[1.0]]))
    hm = enc["heatmaps"][None]
    assert hm.shape == (1, 1, 192, 192)
    assert abs(float(hm.max()) - 0.9970669150352478) <= 1e-5 * 0.9970669150352478
    mod = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus")
    loss = mod(torch.zeros(hm.shape).cuda(), torch.from_numpy(hm).cuda(), torch.tensor([[1.0]]).cuda())
    assert float(loss) == 0.0
    with pytest.raises(AssertionError, match="normalized"):
        mod(torch.zeros(1, 1, 8, 8).cuda(), torch.full((1, 1, 8, 8), 1.5).cuda())
    with pytest.raises(AssertionError):
        mod(torch.zeros(1, 2, 8, 8).cuda(), torch.zeros(1, 2, 8, 8).cuda(), torch.ones(1, 3).cuda())


def test_loss_closed_form_gradient_c4_shape(pp):
    """Full-size shape (96x72): gradient of the fused kernel vs the float64 closed form; linearity in
    the upstream gradient; determinism."""
    wl = synth.WORKLOADS[4]
    B = 4
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=17)
    tgt = _oracle_encode("argmax", wl, kps, vis)["heatmaps"]
    pred = synth.blob_predictions_numpy(_oracle_encode("argmax", wl, synth.jitter_keypoints(wl, kps, 18), vis)["heatmaps"],
                                        synth.blob_params(tgt.shape[:2], 19), 20)
    w = (np.random.default_rng(21).random((B, 17)) < 0.8).astype(np.float32)
    kw = dict(smoothing_weight=0.05, oks_type="minus")
    mod = pp.OKSHeatmapLoss(use_target_weight=True, **kw)
    o = torch.from_numpy(pred).cuda().requires_grad_(True)
    l = mod.forward_mean(o, torch.from_numpy(tgt).cuda(), torch.from_numpy(w).cuda())
    l.backward()
    want = oc.oks_heatmap_loss_grad_closed_form(pred, tgt, w[:, :, None, None], **kw)
    _close(o.grad.cpu().numpy(), want, RTOL32)
    l_ref = oc.oks_heatmap_loss(torch.from_numpy(pred), torch.from_numpy(tgt), torch.from_numpy(w), per_pixel=True, **kw).mean()
    assert abs(l.item() - l_ref.item()) <= RTOL32 * abs(l_ref.item())
    o2 = torch.from_numpy(pred).cuda().requires_grad_(True)
    l2 = mod.forward_mean(o2, torch.from_numpy(tgt).cuda(), torch.from_numpy(w).cuda())
    (l2 * 0.25).backward()
    assert l2.item() == l.item()                                         # deterministic reduction
    _close(o2.grad.cpu().numpy() * 4.0, o.grad.cpu().numpy(), 1e-6)      # linear in the upstream gradient


def test_loss_bf16(pp):
    """The reference cannot run the loss in bfloat16 (Sobel kernels are float32); the
    bf16 oracle is the float32 reference on bf16-rounded inputs, tolerance 1e-2."""
    torch.manual_seed(9)
    B, K, H, W = 2, 5, 64, 48
    out = torch.rand(B, K, H, W).bfloat16()
    tgt = torch.rand(B, K, H, W).bfloat16()
    tw = (torch.rand(B, K) < 0.8).float()
    kw = dict(smoothing_weight=0.05, oks_type="minus")
    o_ref = out.float().requires_grad_(True)
    l_ref = oc.oks_heatmap_loss(o_ref, tgt.float(), tw, per_pixel=True, **kw).mean()
    l_ref.backward()
    mod = pp.OKSHeatmapLoss(use_target_weight=True, **kw)
    o = out.cuda().requires_grad_(True)
    l = mod.forward_mean(o, tgt.cuda(), tw.cuda())
    l.backward()
    assert o.grad.dtype == torch.bfloat16
    assert abs(l.item() - l_ref.item()) <= RTOL16 * abs(l_ref.item())
    _close(o.grad.float().cpu().numpy(), o_ref.grad.numpy(), RTOL16)


def test_decoders_bf16(pp, expected_kernel):
    wl = synth.WORKLOADS[3]
    kps, vis, _ = synth.make_keypoints(wl, batch=3, seed=55)
    maps = _oracle_encode("argmax", wl, kps, vis)["heatmaps"]
    pred = synth.blob_predictions_numpy(maps, synth.blob_params(maps.shape[:2], 56), 57)
    p16 = torch.from_numpy(pred).bfloat16()
    rounded = p16.float().numpy()
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = pm.decode_device(p16.cuda())
    for b in range(3):
        locs, vals, conv = oc.heatmap_expected_value(rounded[b], wl.sigmas, return_heatmap=True, conv="scipy")
        assert np.array_equal(dev["argmax"][b].cpu().numpy(), conv.reshape(17, -1).argmax(1))   # ties are common in bf16
        np.testing.assert_allclose(dev["locs"][b].cpu().numpy(), locs, rtol=RTOL16, atol=1e-2)
        assert np.array_equal(dev["vals"][b].cpu().numpy(), vals)
    peaks, scores = oc.heatmap_maximum(rounded)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    d = am.decode_device(p16.cuda())
    assert np.array_equal(d["peaks"].cpu().numpy(), peaks) and np.array_equal(d["scores"].cpu().numpy(), scores)


# --------------------------------------------------------------------------- round trips at full size
@pytest.mark.parametrize("cid", [2, 4, 5])
def test_round_trip_full_size(pp, cid):
    """encode -> decode at BASELINE.json's full batch sizes: for in-image, labelled keypoints the
    expected-OKS argmax is the pixel nearest to the encoded keypoint (size-independent property)."""
    wl = synth.WORKLOADS[cid]
    kps, vis, _ = synth.make_keypoints(wl)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    enc = am.encode_batch(kps, vis)
    hm = enc["heatmaps"]
    assert tuple(hm.shape) == (wl.batch, wl.num_keypoints, wl.heatmap_size[1], wl.heatmap_size[0])
    dec = pm.decode_device(hm)
    dark = am.decode_device(hm)
    hm_kp = kps / am.scale_factor
    W, H = wl.heatmap_size
    live = (vis >= 0.5) & (hm_kp[..., 0] >= 1) & (hm_kp[..., 0] <= W - 2) & (hm_kp[..., 1] >= 1) & (hm_kp[..., 1] <= H - 2)
    frac = hm_kp - np.floor(hm_kp)
    unambiguous = live & (np.abs(frac - 0.5) > 0.02).all(-1)
    nearest = np.floor(hm_kp + 0.5)
    peaks = dark["peaks"].cpu().numpy()
    assert np.array_equal(peaks[unambiguous], nearest[unambiguous].astype(np.float32))
    am_idx = dec["argmax"].cpu().numpy()
    got_xy = np.stack([am_idx % W, am_idx // W], -1)
    # away from the borders (the reflect-mode convolution pulls border peaks inwards) the expected-OKS
    # argmax is the nearest pixel too
    deep = unambiguous & (hm_kp[..., 0] >= 10) & (hm_kp[..., 0] <= W - 11) & (hm_kp[..., 1] >= 10) & (hm_kp[..., 1] <= H - 11)
    assert (np.abs(got_xy[deep] - nearest[deep]) <= 1).all()
    assert (np.abs(dark["locs"].cpu().numpy()[deep] - hm_kp[deep]) < 0.1).all()  # DARK recovers the sub-pixel position
    dead = vis < 0.5
    assert (peaks[dead] == -1).all() and (dark["scores"].cpu().numpy()[dead] == 0).all()
    # loss of the target against itself: closed form check of the fused kernel on the full batch
    mod = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)
    o = hm.clone().requires_grad_(True)
    l = mod.forward_mean(o, hm, enc["keypoint_weights"])
    l.backward()
    assert torch.isfinite(l) and torch.isfinite(o.grad).all()
    per = mod(hm[:2], hm[:2], enc["keypoint_weights"][:2], per_pixel=True)
    assert abs(float(per.mean()) - float(mod.forward_mean(hm[:2], hm[:2], enc["keypoint_weights"][:2]))) <= 1e-6 * abs(float(per.mean())) + 1e-12


def test_empty_batch_and_errors(pp):
    wl = synth.WORKLOADS[1]
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    out = pm.decode_device(torch.zeros((0, 17, 64, 48), device="cuda"))
    assert out["locs"].shape == (0, 17, 2)
    enc = pm.encode_batch(np.zeros((0, 17, 2), dtype=np.float32))
    assert enc["heatmaps"].shape == (0, 17, 64, 48)
    with pytest.raises(TypeError):
        pm.decode_device(torch.zeros((1, 17, 64, 48), device="cuda", dtype=torch.float16))
    with pytest.raises(IndexError):
        pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas[:5]).decode_device(torch.zeros((1, 17, 64, 48), device="cuda"))


def test_codec_decode_tuple(pp):
    """Codec.decode of the model 5-tuple (codec.py:249-263) on a batch."""
    wl = synth.WORKLOADS[2]
    B, K = 5, 17
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=71)
    maps = _oracle_encode("argmax", wl, kps, vis)["heatmaps"]
    rng = np.random.default_rng(72)
    heads = [torch.from_numpy(rng.random((B, K, 1, 1), dtype=np.float32)).cuda() for _ in range(4)]
    codec = pp.Codec(pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas))
    (kp, sc), prob, visb, oks, err = codec.decode((torch.from_numpy(maps).cuda(), *heads))
    assert kp.shape == (B, K, 2) and sc.shape == (B, K)
    for a, h in ((prob, heads[0]), (visb, heads[1]), (oks, heads[2])):
        assert a.shape == (B, 1, K) and np.array_equal(a, h.cpu().numpy().reshape(B, 1, K))
    np.testing.assert_allclose(err, heads[3].cpu().numpy().reshape(B, 1, K) / np.sqrt(64 ** 2 + 48 ** 2), rtol=1e-7)
    for b in range(B):
        k_ref, s_ref = oc.decode_expected(maps[b], wl.input_size, wl.heatmap_size, wl.sigmas, conv="scipy")
        np.testing.assert_allclose(kp[b], k_ref[0], rtol=RTOL32, atol=1e-5)
        assert np.array_equal(sc[b], s_ref[0])
    rec = codec.decode_device((torch.from_numpy(maps).cuda(), *heads))
    assert rec.shape == (B, K, 7) and np.array_equal(rec[..., :2].cpu().numpy(), kp)
    hk, hs = codec.decode_heatmap(torch.from_numpy(maps[0]).cuda())
    assert hk.shape == (1, K, 2) and np.array_equal(hk[0], kp[0])


def test_head_tail_autograd_and_patch(pp):
    """The fused tail against torch.clamp(x / t, 0, 1) incl. autograd (head.py:526-532), and the patch helper
    on a stand-in module with the reference head's attribute names."""
    torch.manual_seed(4)
    x = (torch.randn(3, 5, 16, 12, device="cuda") * 0.4)
    x[0, 0, 0, :4] = torch.tensor([0.0, 0.5, -0.0, 0.25], device="cuda")   # boundary values: gradient passes at 0 and 1
    for dtype, tol in ((torch.float32, 0.0), (torch.bfloat16, 0.0)):
        a = x.to(dtype).clone().requires_grad_(True)
        b = x.to(dtype).clone().requires_grad_(True)
        up = torch.rand_like(a)
        ya = pp.heatmap_tail(a, 0.5)
        yb = torch.clamp(b / 0.5, 0, 1)
        assert torch.equal(ya, yb)
        (ya * up).sum().backward()
        (yb * up).sum().backward()
        assert torch.equal(a.grad, b.grad)
    t3 = torch.clamp(x / 0.3, 0, 1)
    # torch's CUDA kernel multiplies by the reciprocal of a scalar divisor; the kernel divides (IEEE), 1 ulp apart
    torch.testing.assert_close(pp.heatmap_tail(x, 0.3), t3, rtol=1e-6, atol=1e-7)

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.deconv_layers = torch.nn.Identity()
            self.conv_layers = torch.nn.Conv2d(5, 5, 1)
            self.final_layer = torch.nn.Identity()
            self.temperature = 0.5
            self.normalize = None

        def forward_heatmap(self, x):
            x = self.final_layer(self.conv_layers(self.deconv_layers(x)))
            B, C, H, W = x.shape
            x = x.reshape((B, C, H * W)) / self.temperature
            return torch.clamp(x, 0, 1).reshape((B, C, H, W))

    m = Stub().cuda()
    want = m.forward_heatmap(x)
    got = pp.patch_probmap_head(m).forward_heatmap(x)
    assert torch.equal(got, want)
    got.sum().backward()
    assert m.conv_layers.weight.grad is not None


def test_pose_targets_from_heatmaps(pp, golden_dir, dark_kernel):
    """Device-side ProbPoseLoss._oks_from_heatmaps / _error_from_heatmaps against the reference's outputs."""
    from probpose_pytorch_b200.pose_targets import error_from_heatmaps, oks_from_heatmaps
    g = np.load(golden_dir / "decode.npz")
    t = np.load(golden_dir / "targets.npz")
    wl = synth.WORKLOADS[3]
    codec = pp.Codec(pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas))
    gt, dt = torch.from_numpy(g["clean"]).cuda(), torch.from_numpy(g["blob"]).cuda()
    oks, w = oks_from_heatmaps(codec, gt, dt, torch.from_numpy(t["weight"]).cuda(), heatmap_size=wl.heatmap_size)
    assert oks.dtype == torch.float32 and tuple(oks.shape) == t["oks"].shape
    assert np.array_equal(w.cpu().numpy(), t["oks_weights"])
    # The targets are functions of two DARK decodes.  On blob-shaped channels (the only ones DARK is defined for, see
    # test_dark_decoder_matches_golden_reference) they must agree with the reference to 1e-5 plus what the decoder's own
    # conditioning allows: the coordinate tolerance of _dark_tolerance (cv2's float32 blur is reproducible to ~2e-7,
    # amplified by the inverse Hessian of the log-map) propagated through oks = exp(-(dx^2 + dy^2) / (2 var area)).
    got, want = oks.cpu().numpy(), t["oks"]
    err = error_from_heatmaps(codec, gt, dt).cpu().numpy()
    var = (np.asarray(wl.sigmas) * 2) ** 2
    area = wl.heatmap_size[0] * wl.heatmap_size[1] * 0.53
    n_blob = 0
    for b in range(got.shape[0]):
        live = g["blob_peaks"][b, :, 0] >= 0
        blobby = live & (g["blob"][b].reshape(17, -1).max(axis=1) >= 0.1) & (g["clean"][b].reshape(17, -1).max(axis=1) >= 0.1)
        tol = (_dark_tolerance(g["blob"][b], g["blob_peaks"][b], wl, RTOL32) + _dark_tolerance(g["clean"][b], g["clean_peaks"][b], wl, RTOL32)
               + RTOL32 * np.maximum(np.abs(g["blob_dark_keypoints"][b]), 1.0))           # (K, 2) input px
        d = np.abs(g["blob_dark_keypoints"][b] - g["clean_dark_keypoints"][b])           # (K, 2)
        w_b = t["weight"][b].reshape(-1) > 0
        d_oks = want[b] * ((d * tol).sum(1) + 0.5 * (tol ** 2).sum(1)) / (var * area)
        bound = RTOL32 * np.abs(want[b]) + 1e-7 + d_oks
        sel = blobby & w_b
        assert (np.abs(got[b] - want[b])[sel] <= bound[sel]).all(), (b, np.max((np.abs(got[b] - want[b]) / bound)[sel]))
        e_bound = RTOL32 * np.maximum(t["error"][b], 1.0) + np.sqrt((tol ** 2).sum(1))
        assert (np.abs(err[b] - t["error"][b])[blobby] <= e_bound[blobby]).all(), (b, np.max((np.abs(err[b] - t["error"][b]) / e_bound)[blobby]))
        assert np.isfinite(got[b]).all() and np.isfinite(err[b][live]).all()
        n_blob += int(sel.sum())
    assert n_blob >= 12                                     # a third of the fixture's channels are weighted blobs
    assert (got[t["weight"] == 0] == 0).all()
    # the same through the oracle restatement (numpy blur) on the identical inputs, tighter on clean maps
    o2, w2 = oc.oks_from_heatmaps(g["clean"], g["clean"], t["weight"], wl.sigmas, wl.input_size, wl.heatmap_size,
                                  area_size=wl.heatmap_size)
    o3, w3 = oks_from_heatmaps(codec, gt, gt, torch.from_numpy(t["weight"]).cuda(), heatmap_size=wl.heatmap_size)
    np.testing.assert_allclose(o3.cpu().numpy(), o2, rtol=RTOL32, atol=1e-6)
    assert np.array_equal(w3.cpu().numpy(), w2)


# --------------------------------------------------------------------------- shape / content sweep
def _sweep_maps(rng, K, H, W):
    """Heatmaps that stress the pruned kernels: blobs hugging borders and corners, two competing blobs,
    saturated plateaus, pure noise, a single hot pixel, negative values."""
    ys, xs = np.mgrid[0:H, 0:W]
    maps = np.zeros((K, H, W), dtype=np.float32)
    for k in range(K):
        kind = k % 7
        cx, cy = rng.uniform(-1, W), rng.uniform(-1, H)
        if kind == 0:    # blob near a border / corner
            cx, cy = rng.choice([0.3, W - 1.2, rng.uniform(0, W - 1)]), rng.choice([0.4, H - 1.3, rng.uniform(0, H - 1)])
        s = rng.uniform(0.6, 3.0)
        blob = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * s)).astype(np.float32)
        if kind in (0, 1):
            maps[k] = blob * rng.uniform(0.3, 1.0)
        elif kind == 2:  # two blobs of similar height, far apart
            cx2, cy2 = W - 1 - cx, H - 1 - cy
            maps[k] = 0.8 * blob + 0.79 * np.exp(-((xs - cx2) ** 2 + (ys - cy2) ** 2) / (2 * s)).astype(np.float32)
        elif kind == 3:  # saturated plateau
            maps[k] = np.clip(blob * 4.0, 0, 1)
        elif kind == 4:  # noise floor only
            maps[k] = rng.random((H, W), dtype=np.float32) * 0.02
        elif kind == 5:  # one hot pixel on a constant background
            maps[k] = 0.125
            maps[k, rng.integers(0, H), rng.integers(0, W)] = 0.75
        else:            # blob with negative surroundings
            maps[k] = blob - 0.25
        if kind in (0, 1, 2, 6):
            maps[k] += rng.random((H, W), dtype=np.float32) * 0.01
    return maps


@pytest.mark.parametrize("H,W", [(12, 12), (16, 20), (24, 20), (33, 28), (40, 36), (64, 48), (80, 64), (96, 72), (128, 96)])
def test_decoders_shape_sweep(pp, H, W, expected_kernel):
    rng = np.random.default_rng(H * 1000 + W)
    K = 14
    sigmas = rng.uniform(0.02, 0.12, size=K)          # radii 2 .. 9
    maps = np.stack([_sweep_maps(rng, K, H, W) for _ in range(2)])
    pm = pp.ProbMap((W * 4, H * 4), (W, H), sigmas)
    am = pp.ArgMaxProbMap((W * 4, H * 4), (W, H), sigmas)
    t = torch.from_numpy(maps).cuda()
    dev = pm.decode_device(t)
    dark = am.decode_device(t)
    # the same maps through the fallback kernels (a mis-aligned view rules the TMA kernels out)
    buf = torch.empty(maps.size + 1, device="cuda")
    mis = buf[1:].view(maps.shape)
    mis.copy_(t)
    dev_v1 = pm.decode_device(mis)
    dark_v1 = am.decode_device(mis)
    for b in range(maps.shape[0]):
        locs, vals, conv = oc.heatmap_expected_value(maps[b], sigmas, return_heatmap=True, conv="scipy")
        am_ref = conv.reshape(K, -1).argmax(1)
        for name, d in (("tma", dev), ("fallback", dev_v1)):
            assert np.array_equal(d["argmax"][b].cpu().numpy(), am_ref), (name, H, W, b)
            np.testing.assert_allclose(d["locs"][b].cpu().numpy(), locs, rtol=RTOL32, atol=1e-5, err_msg=name)
            assert np.array_equal(d["vals"][b].cpu().numpy(), vals), name
        peaks, scores = oc.heatmap_maximum(maps[b])
        for name, d in (("tma", dark), ("fallback", dark_v1)):
            assert np.array_equal(d["peaks"][b].cpu().numpy(), peaks), (name, H, W, b)
            assert np.array_equal(d["scores"][b].cpu().numpy(), scores), name
        # the two DARK kernels blur with different operation orders on flat maps; on blob-shaped channels
        # (kinds 0-3, 6) they must agree closely
        blobby = np.array([k % 7 in (0, 1, 3) for k in range(K)]) & (peaks[:, 0] >= 0)
        a, c = dark["locs"][b].cpu().numpy(), dark_v1["locs"][b].cpu().numpy()
        assert np.abs(a - c)[blobby].max() <= 2e-3, (H, W, b, np.abs(a - c)[blobby].max())
        assert np.isfinite(a).all()


def test_expected_decoder_small_and_odd_shapes(pp):
    """Shapes the TMA kernels do not take (odd widths, tiny maps) still decode exactly."""
    rng = np.random.default_rng(77)
    for H, W in ((7, 9), (10, 10), (19, 27), (21, 30)):
        K = 6
        sigmas = rng.uniform(0.02, 0.06, size=K)
        maps = _sweep_maps(rng, K, H, W)
        r_max = int(np.ceil(3 * oc.oks_variance_table(sigmas, H, W)).max())
        if r_max >= min(H, W):      # scipy reflects repeatedly there; not a shape the codec is used with
            continue
        l, v = pp.get_heatmap_expected_value(maps, sigmas)
        l_ref, v_ref, conv = oc.heatmap_expected_value(maps, sigmas, return_heatmap=True, conv="scipy")
        np.testing.assert_allclose(l, l_ref, rtol=RTOL32, atol=1e-5)
        assert np.array_equal(v, v_ref)


@pytest.mark.parametrize("H,W", [(4, 4), (5, 8), (16, 12), (33, 28), (64, 48), (96, 72), (192, 192)])
def test_loss_fast_path_shape_sweep(pp, H, W):
    """Fused mean-mode kernel (TMA fast path) against the oracle and against the general kernel, across shapes
    incl. tiny maps, a strip count that is not a multiple of anything, and a map that needs single-stage TMA."""
    torch.manual_seed(H * 100 + W)
    B, K = (2, 3) if H * W > 20000 else (3, 5)
    out, tgt = torch.rand(B, K, H, W), torch.rand(B, K, H, W)
    tw = (torch.rand(B, K) < 0.7).float()
    for kw in (dict(smoothing_weight=0.05, oks_type="minus"),
               dict(smoothing_weight=0.2, gaussian_weight=0.15, oks_type="both", loss_weight=1.7)):
        o_ref = out.clone().requires_grad_(True)
        l_ref = oc.oks_heatmap_loss(o_ref, tgt, tw, per_pixel=True, **kw).mean()
        l_ref.backward()
        mod = pp.OKSHeatmapLoss(use_target_weight=True, **kw)
        o = out.cuda().requires_grad_(True)
        l = mod.forward_mean(o, tgt.cuda(), tw.cuda())
        (l * 2.0).backward()
        assert abs(l.item() - l_ref.item()) <= RTOL32 * abs(l_ref.item()) + 1e-9
        _close(o.grad.cpu().numpy(), 2.0 * o_ref.grad.numpy(), RTOL32)
        # general kernel, same mode through per_pixel + mean
        o2 = out.cuda().requires_grad_(True)
        l2 = mod(o2, tgt.cuda(), tw.cuda(), per_pixel=True).mean()     # 192 x 192: several row bands
        l2.backward()
        assert abs(l2.item() - l_ref.item()) <= RTOL32 * abs(l_ref.item()) + 1e-9
        _close(o2.grad.cpu().numpy(), o_ref.grad.numpy(), RTOL32)
    # forward only (no grad requested) and backward through the non-fused route
    with torch.no_grad():
        l3 = pp.OKSHeatmapLoss(smoothing_weight=0.05)(out.cuda(), tgt.cuda(), tw.cuda(), per_pixel=False)
    l3_ref = oc.oks_heatmap_loss(out, tgt, tw, smoothing_weight=0.05)
    assert abs(l3.item() - l3_ref.item()) <= RTOL32 * abs(l3_ref.item()) + 1e-9


@pytest.mark.parametrize("H,W", [(8, 8), (17, 23), (64, 48), (100, 60), (192, 192), (300, 8)])
def test_encode_shape_sweep(pp, H, W):
    rng = np.random.default_rng(H + 7 * W)
    K, B = 9, 3
    wl = synth.Workload(8, "sweep", B, K, (W * 4, H * 4), (W, H), True)
    kps, vis, _ = synth.make_keypoints(wl, seed=H * W)
    sig = rng.uniform(0.02, 0.12, size=K)
    for kind, cls, sigma in (("probmap", pp.ProbMap, 2.0), ("argmax", pp.ArgMaxProbMap, -1), ("argmax", pp.ArgMaxProbMap, 0.8)):
        codec = cls(wl.input_size, wl.heatmap_size, sig, sigma=sigma)
        got = codec.encode_batch(kps, vis)
        want = np.stack([oc.encode(kind, wl.input_size, wl.heatmap_size, sig, kps[b:b + 1], vis[b:b + 1], sigma=sigma)["heatmaps"]
                         for b in range(B)])
        _assert_maps_close(got["heatmaps"].cpu().numpy(), want, RTOL32)


@pytest.mark.parametrize("band", [0, 1, 3, 7])
def test_loss_general_kernel_row_bands(pp, band, monkeypatch):
    """The general loss kernel walks a heatmap in row bands when three planes do not fit shared memory.
    PP_LOSS_BAND forces short bands on a small map (band seams at every 1 / 3 / 7 rows); band=0 is a
    256 x 200 map that needs bands on its own.  All three modes, with pixel weights and a mask."""
    if band:
        monkeypatch.setenv("PP_LOSS_BAND", str(band))
        B, K, H, W = 2, 3, 23, 19
    else:
        B, K, H, W = 1, 2, 256, 200
    torch.manual_seed(band + 11)
    out, tgt = torch.rand(B, K, H, W), torch.rand(B, K, H, W)
    tgt[0, 1] = 0.0                                   # an empty channel for skip_empty_channel
    twp = torch.rand(B, K, H, W)
    mask = (torch.rand(B, 1, H, W) < 0.8).float()
    up_px, up_k = torch.rand(B, K, H, W), torch.rand(B, K)
    kw = dict(smoothing_weight=0.15, gaussian_weight=0.1, oks_type="both", loss_weight=0.7, skip_empty_channel=True)
    mod = pp.OKSHeatmapLoss(**kw)
    for mode, up in ((dict(per_pixel=True), up_px), (dict(per_keypoint=True), up_k), (dict(), torch.tensor(1.3))):
        o_ref = out.clone().requires_grad_(True)
        l_ref = oc.oks_heatmap_loss(o_ref, tgt, twp, mask, **mode, **kw)
        (l_ref * up).sum().backward()
        o = out.cuda().requires_grad_(True)
        l = mod(o, tgt.cuda(), twp.cuda(), mask.cuda(), **mode)
        (l * up.cuda()).sum().backward()
        _close(l.detach().cpu().numpy(), l_ref.detach().numpy(), RTOL32)
        _close(o.grad.cpu().numpy(), o_ref.grad.numpy(), RTOL32)


def test_dark_decoder_large_maps_generic_path(pp):
    """192 x 192 maps (the shape of the reference's tests/test_loss.py) do not fit shared memory: the
    global-memory DARK path must agree with the oracle (and bit-for-bit with itself on repeated runs)."""
    H = W = 192
    codec = pp.ArgMaxProbMap((768, 768), (W, H), np.array([0.1] * 4))
    kps = np.array([[[96.0, 96.0], [400.5, 300.25], [700.0, 20.0], [-40.0, 900.0]]])
    vis = np.array([[1.0, 1.0, 1.0, 1.0]])
    maps = codec.encode(kps, vis)["heatmaps"]
    assert maps.shape == (4, H, W)
    kp, sc = codec.decode(maps)
    kp_ref, sc_ref = oc.decode_argmax_dark(maps, (768, 768), (W, H), backend="cv2")
    assert np.array_equal(sc, sc_ref)
    live = sc_ref[0] > 0.1
    np.testing.assert_allclose(kp[0][live], kp_ref[0][live], rtol=RTOL32, atol=1e-4)
    kp2, _ = codec.decode(maps)
    assert np.array_equal(kp, kp2)
    # expected-OKS decoder on the same large maps (exact full-map path)
    pm = pp.ProbMap((768, 768), (W, H), np.array([0.1] * 4))
    k2, s2 = pm.decode(maps)
    k2_ref, s2_ref = oc.decode_expected(maps, (768, 768), (W, H), np.array([0.1] * 4), conv="scipy")
    np.testing.assert_allclose(k2, k2_ref, rtol=RTOL32, atol=1e-5)
    assert np.array_equal(s2, s2_ref)


def test_host_pipeline_orders_results_and_reuses_buffers(pp):
    """host_io.pipelined_steps: results come back in order, one per batch, although input buffers are reused
    and copies overlap the kernels; the step is the real encode -> decode round trip."""
    from probpose_pytorch_b200.host_io import pipelined_steps
    wl = synth.WORKLOADS[1]
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    batches, want = [], []
    for i in range(7):
        kps, vis, _ = synth.make_keypoints(wl, batch=4, seed=300 + i)
        batches.append((torch.from_numpy(kps).pin_memory(), torch.from_numpy(vis).pin_memory()))
        enc = am.encode_batch(torch.from_numpy(kps).cuda(), torch.from_numpy(vis).cuda())
        want.append(am.decode_device(enc["heatmaps"])["keypoints"].cpu())

    def step(kps, vis):
        enc = am.encode_batch(kps, vis)
        return am.decode_device(enc["heatmaps"])["keypoints"], enc["keypoint_weights"].sum()

    got = list(pipelined_steps(iter(batches), step, torch.device("cuda:0")))
    assert len(got) == len(batches)
    for (rec, wsum), w, (kps, vis) in zip(got, want, batches):
        assert torch.equal(rec, w)
    assert list(pipelined_steps(iter([]), step, torch.device("cuda:0"))) == []
    with pytest.raises(ValueError):
        list(pipelined_steps(iter([(torch.zeros(1, 17, 2), torch.zeros(1, 17))]), step, torch.device("cuda:0")))


# ---- Sparsemax-normalised head tail (oracle = published algorithm, parity unpinned) --------
# The projection's outputs are O(1 / support); the float32 sort + cumsum of the published algorithm carries
# ~1e-7 of absolute noise in tau, so values are compared at atol 1e-6 (of a [0, 1] range) + rtol 1e-5, and
# against the float64 evaluation of the same projection at the same bar.
SPARSE_ATOL = 1e-6


@pytest.mark.parametrize("shape,scale", [((3, 5, 64, 48), 3.0), ((2, 17, 64, 48), 0.05), ((2, 3, 7, 5), 1.0),
                                         ((1, 2, 96, 72), 20.0), ((1, 2, 100, 100), 1.0), ((1, 2, 250, 240), 2.0)])
def test_sparsemax_tail_forward_backward(pp, shape, scale):
    """clamp(sparsemax(x / T) * normalize, 0, 1) and its gradient: dense supports (small logits), sparse
    supports (large logits), a map larger than shared memory (250 x 240: logits re-read from L2)."""
    torch.manual_seed(shape[-1] + int(scale * 10))
    x = torch.randn(shape) * scale
    x[0, 0].view(-1)[:6] = x[0, 0].max() + 0.3           # exact ties at the maximum
    up = torch.rand(shape)
    for T, normalize in ((0.5, 1.0), (0.5, 2.5), (1.0, 0.5)):
        a = x.clone().requires_grad_(True)
        ya = oc.head_tail_sparsemax(a, T, normalize)
        (ya * up).sum().backward()
        b = x.cuda().requires_grad_(True)
        yb = pp.heatmap_tail(b, T, normalize=normalize)
        (yb * up.cuda()).sum().backward()
        torch.testing.assert_close(yb.detach().cpu(), ya.detach(), rtol=RTOL32, atol=SPARSE_ATOL)
        p64 = oc.sparsemax_f64((x / T).reshape(*shape[:2], -1).numpy()).reshape(shape)
        np.testing.assert_allclose(yb.detach().cpu().numpy(), np.clip(p64 * normalize, 0, 1), rtol=RTOL32, atol=SPARSE_ATOL)
        # the support (and with it the routing of the gradient) may differ on pixels whose output is ~1e-7
        ga, gb = a.grad.numpy(), b.grad.cpu().numpy()
        agree = np.abs(ga - gb) <= 1e-4 * np.abs(ga) + 1e-5 * np.abs(ga).max()
        assert agree.mean() >= 0.999, agree.mean()
        if normalize == 1.0:
            s = yb.detach().reshape(*shape[:2], -1).sum(-1)
            torch.testing.assert_close(s, torch.ones_like(s), rtol=0, atol=1e-4)
    # inference call: no aux, no graph
    with torch.no_grad():
        assert torch.equal(pp.heatmap_tail(x.cuda(), 0.5, normalize=1.0), pp.heatmap_tail(x.cuda().requires_grad_(True), 0.5, normalize=1.0))


def test_sparsemax_module_bf16_and_patched_head(pp):
    torch.manual_seed(9)
    x = torch.randn(4, 6, 300)
    y = pp.Sparsemax(dim=-1)(x.cuda())
    torch.testing.assert_close(y.cpu(), oc.sparsemax(x), rtol=RTOL32, atol=SPARSE_ATOL)
    y1 = pp.Sparsemax(dim=1)(x.cuda())                    # any axis, as the package allows
    torch.testing.assert_close(y1.cpu(), oc.sparsemax(x.movedim(1, -1)).movedim(-1, 1), rtol=RTOL32, atol=SPARSE_ATOL)
    # bf16: oracle = float32 algorithm on the bf16-rounded logits, result rounded to bf16
    xb = (torch.randn(2, 3, 16, 12) * 2).bfloat16()
    want = oc.head_tail_sparsemax((xb.float() / 0.5).bfloat16().float(), 1.0, 1.0)
    got = pp.heatmap_tail(xb.cuda(), 0.5, normalize=1.0)
    assert got.dtype == torch.bfloat16
    torch.testing.assert_close(got.float().cpu(), want, rtol=RTOL16, atol=1e-3)

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.deconv_layers = torch.nn.Identity()
            self.conv_layers = torch.nn.Conv2d(5, 5, 1)
            self.final_layer = torch.nn.Identity()
            self.temperature = 0.5
            self.normalize = 1.0
            self.normalize_layer = lambda z: oc.sparsemax(z.cpu()).to(z.device)   # stands in for the package

        def forward_heatmap(self, x):
            x = self.final_layer(self.conv_layers(self.deconv_layers(x)))
            B, C, H, W = x.shape
            x = self.normalize_layer(x.reshape(B, C, H * W) / self.temperature) * self.normalize
            return torch.clamp(x, 0, 1).reshape(B, C, H, W)

    m = Stub().cuda()
    xin = torch.randn(2, 5, 16, 12, device="cuda") * 3
    want = m.forward_heatmap(xin)
    got = pp.patch_probmap_head(m).forward_heatmap(xin)
    torch.testing.assert_close(got, want, rtol=RTOL32, atol=SPARSE_ATOL)
    got.square().sum().backward()
    assert m.conv_layers.weight.grad is not None and torch.isfinite(m.conv_layers.weight.grad).all()


# ---- validation metrics ------------------------------------------------------------------
def test_pck_metrics_match_reference_outputs(pp, golden_dir):
    """pose_pck_accuracy / keypoint_pck_accuracy / get_binary_accuracy / get_mae against the reference's own
    outputs (tests/golden/metrics.npz): per-keypoint accuracies and counts bit-exact, the average to 1e-12."""
    from probpose_pytorch_b200 import metrics
    m = np.load(golden_dir / "metrics.npz")
    g = np.load(golden_dir / "decode.npz")
    for thr in (0.05, 0.2):
        acc, avg, cnt = metrics.pose_pck_accuracy(g["blob"], g["clean"], m["mask"], thr=thr)
        assert np.array_equal(acc, m[f"pose/{thr}/acc"]) and cnt == m[f"pose/{thr}/cnt"]
        assert abs(avg - m[f"pose/{thr}/avg"]) <= 1e-12
    norm = m["norm64"].copy()
    acc, avg, cnt = metrics.pose_pck_accuracy(torch.from_numpy(g["blob"]).cuda(), torch.from_numpy(g["clean"]).cuda(),
                                              m["mask"], thr=0.1, normalize=norm)
    assert np.array_equal(acc, m["pose/norm64/acc"]) and cnt == m["pose/norm64/cnt"] and abs(avg - m["pose/norm64/avg"]) <= 1e-12
    assert np.array_equal(norm, m["norm64"])                       # the caller's array is left alone
    acc, avg, cnt = metrics.keypoint_pck_accuracy(m["pred"], m["gt"], m["mask"], 0.05, m["norm32"])
    assert np.array_equal(acc, m["kpt/acc"]) and cnt == m["kpt/cnt"] and abs(avg - m["kpt/avg"]) <= 1e-12
    acc, avg, cnt = metrics.keypoint_pck_accuracy(m["pred"], m["gt"], np.zeros_like(m["mask"]), 0.05, m["norm32"])
    assert (acc == -1).all() and avg == 0.0 and cnt == 0
    # distances against the oracle, both arithmetic types
    for norm in (m["norm32"], m["norm64"]):
        d = metrics.keypoint_pck_accuracy_device(m["pred"], m["gt"], m["mask"], 0.05, norm, return_distances=True)[3]
        assert np.array_equal(d.cpu().numpy(), oc.metrics_oracle.calc_distances(m["pred"], m["gt"], m["mask"], norm))
    with pytest.raises(ValueError):
        metrics.pose_pck_accuracy(g["blob"], g["clean"], m["mask"], method="median")
    a, t = metrics.get_binary_accuracy(torch.from_numpy(m["scalar_dt"]).cuda(), torch.from_numpy(m["scalar_gt"]).cuda(),
                                       torch.from_numpy(m["mask"]).cuda())
    assert a.item() == m["binary/acc"] and t.item() == m["binary/thr"]
    mae = metrics.get_mae(torch.from_numpy(m["scalar_dt"]).cuda(), torch.from_numpy(m["scalar_gt"]).cuda(),
                          torch.from_numpy(m["mask"]).cuda())
    assert abs(mae.item() - float(m["mae"])) <= 1e-6 * float(m["mae"])
    pa = metrics.get_pose_accuracy(torch.from_numpy(g["blob"]).cuda(), torch.from_numpy(g["clean"]).cuda(), m["mask"])
    assert pa.is_cuda and abs(pa.item() - m["pose/0.05/avg"]) <= 1e-12


def test_pck_metrics_wholebody_batch_against_oracle(pp):
    """K = 133, N = 64: more keypoints than warps, more instances than lanes; random masks and factors."""
    from probpose_pytorch_b200 import metrics
    rng = np.random.default_rng(12)
    N, K = 64, 133
    pred = rng.uniform(0, 48, size=(N, K, 2)).astype(np.float32)
    gt = (pred + rng.normal(0, 3.0, size=pred.shape)).astype(np.float32)
    mask = rng.random((N, K)) < 0.7
    for norm in (np.tile(np.array([[64, 48]]), (N, 1)), rng.uniform(-5, 60, size=(N, 2)), rng.uniform(1, 60, size=(N, 2)).astype(np.float32)):
        for thr in (0.05, 0.5):
            acc, avg, cnt = metrics.keypoint_pck_accuracy(pred, gt, mask, thr, norm)
            acc_o, avg_o, cnt_o = oc.metrics_oracle.keypoint_pck_accuracy(pred, gt, mask, thr, norm)
            assert np.array_equal(acc, acc_o) and cnt == cnt_o and abs(avg - avg_o) <= 1e-12
    dt = rng.random((N, K)).astype(np.float32)
    gb = (rng.random((N, K)) < 0.3).astype(np.float32)
    a, t = metrics.get_binary_accuracy(dt, gb, mask)
    a_o, t_o = oc.metrics_oracle.binary_accuracy(dt, gb, mask)
    assert a.item() == a_o and t.item() == t_o


@pytest.mark.parametrize("name,freeze,kwargs", [("frozen", True, {}), ("live", False, {}),
                                                ("zeros", False, dict(learn_heatmaps_from_zeros=True)),
                                                ("weights", True, dict(keypoint_weights=True))])
def test_probpose_loss_matches_reference_forward_and_backward(pp, golden_dir, name, freeze, kwargs):
    """The training-step caller with the hot-path members patched (patch_probpose_loss): five losses, five accuracies
    and the gradients of the losses' sum against the reference's own run (tests/golden/probpose_loss.npz).  The
    reference's class is not importable on the GPU box: the patched object is the oracle's stand-in with the same
    member names, driven by the oracle's restatement of the reference's forward (pinned on the CPU by
    tests/test_oracle_golden.py; the real instance is patched in tests/test_reference_patch.py)."""
    g = np.load(golden_dir / "probpose_loss.npz")
    wl = synth.WORKLOADS[3]
    codec = pp.Codec(pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas))
    mod = pp.patch_probpose_loss(oc.ProbPoseLossLayout(codec, freeze_error=freeze))
    assert isinstance(mod.keypoint_loss_module, pp.FusedOKSHeatmapLoss)
    gt = {k.split("/", 1)[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith("gt/")}     # host tensors, as a DataLoader yields
    pred = [torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ("dt_heatmaps", "dt_probs", "dt_vis", "dt_oks", "dt_errs")]
    if kwargs.get("keypoint_weights"):
        kwargs = dict(keypoint_weights=torch.from_numpy(g["keypoint_weights"]).cuda())
    np.random.seed(99)
    losses, acc = oc.training_losses(mod, gt, tuple(pred), compute_acc=True, **kwargs)
    sum(losses.values()).backward()
    for k, v in losses.items():
        want = float(g[f"{name}/loss/{k}"])
        assert abs(v.item() - want) <= RTOL32 * abs(want) + 1e-8, (k, v.item(), want)
    for k, v in acc.items():
        want = float(g[f"{name}/acc/{k}"])
        assert abs(float(v) - want) <= 1e-5 * abs(want) + 1e-8, (k, float(v), want)
    for n, p_ in zip(("heatmaps", "probs", "vis", "oks", "errs"), pred):
        want = g[f"{name}/grad/{n}"]
        if n == "heatmaps":
            _close(p_.grad.cpu().numpy(), want, RTOL32)
        else:   # the OKS / error targets come out of a DARK decode (1e-6 apart): 1e-5 of the gradient's scale
            np.testing.assert_allclose(p_.grad.cpu().numpy(), want, rtol=RTOL32, atol=RTOL32 * np.abs(want).max())
    # without accuracies the call returns the dictionary alone
    assert set(oc.training_losses(mod, gt, tuple(p_.detach() for p_ in pred))) == {"kpt", "probability", "visibility", "oks", "error"}


def test_lazy_per_pixel_loss_is_the_fused_mean_or_the_real_map(pp):
    """What the patched keypoint_loss_module returns for per_pixel=True: .mean() is the fused forward+backward kernel,
    anything else materialises the reference's (B, K, H, W) map."""
    torch.manual_seed(5)
    out, tgt, w = torch.rand(3, 5, 64, 48), torch.rand(3, 5, 64, 48), (torch.rand(3, 5) < 0.7).float()
    o_ref = out.clone().requires_grad_(True)
    ref_map = oc.oks_heatmap_loss(o_ref, tgt, w, per_pixel=True, smoothing_weight=0.05, oks_type="minus")
    ref_map.mean().backward()
    mod = pp.FusedOKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus")
    o = out.cuda().requires_grad_(True)
    lazy = mod(o, tgt.cuda(), w.cuda(), per_pixel=True)
    m = lazy.mean()
    assert lazy._tensor is None                      # nothing (B, K, H, W)-sized was written
    m.backward()
    assert abs(m.item() - ref_map.mean().item()) <= RTOL32 * abs(ref_map.mean().item())
    _close(o.grad.cpu().numpy(), o_ref.grad.numpy(), RTOL32)
    lazy2 = mod(o.detach(), tgt.cuda(), w.cuda(), per_pixel=True)
    _close((lazy2 * 2.0).cpu().numpy(), 2.0 * ref_map.detach().numpy(), RTOL32)
    _close(torch.sum(lazy2, dim=(2, 3)).cpu().numpy(), ref_map.detach().sum(dim=(2, 3)).numpy(), RTOL32)
    assert lazy2[0, 0].shape == (64, 48) and lazy2._tensor is not None


def test_ground_truth_from_keypoints_front_end(pp):
    """A loader that ships keypoints (section 8 f-3): ground_truth_from_keypoints builds the reference-layout GT dict
    on the device; the losses equal the ones from the dict built by per-sample ``encode`` on the host."""
    wl = synth.WORKLOADS[3]
    B = 5
    kps, vis, visibility = synth.make_keypoints(wl, batch=B, seed=21)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    mod = pp.patch_probpose_loss(oc.ProbPoseLossLayout(pp.Codec(am), freeze_error=False))
    enc = [am.encode(kps[b:b + 1], vis[b:b + 1] > 0.5) for b in range(B)]
    gt_ref = dict(heatmaps=torch.from_numpy(np.stack([e["heatmaps"] for e in enc])),
                  in_image=torch.from_numpy(np.concatenate([e["in_image"] for e in enc])),
                  keypoints_visible=torch.from_numpy(vis > 0.5), keypoints_visibility=torch.from_numpy(visibility > 0.5))
    gt_kp = pp.ground_truth_from_keypoints(mod, kps, vis > 0.5, visibility > 0.5)
    assert gt_kp["heatmaps"].is_cuda and gt_kp["heatmaps"].shape == (B, 17, 64, 48)
    jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=22)).cuda()
    hm = (am.encode_batch(jit, None)["heatmaps"] * 0.7 + 0.002).clamp(0, 1)
    torch.manual_seed(2)
    heads = [torch.rand(B, 17, 1, 1, device="cuda") * 0.9 + 0.05 for _ in range(4)]
    a = oc.training_losses(mod, gt_ref, (hm, *heads))
    b = oc.training_losses(mod, gt_kp, (hm, *heads))
    for k in a:
        assert abs(a[k].item() - b[k].item()) <= 1e-6 * abs(a[k].item()) + 1e-9, k


def test_head_tail_matches_reference_head_outputs(pp, golden_dir):
    """tests/golden/head.npz: activations entering the tail of the REFERENCE's ProbMapHead.forward_heatmap (captured
    with a hook on its final_layer) and what the reference returned (head.py:513-534).  The fused tail -- alone, through
    patch_probmap_head on a module with the reference's member names, and fused into both decoders' load -- must give
    the reference's heatmaps bit for bit."""
    g = np.load(golden_dir / "head.npz")
    pre = torch.from_numpy(g["pre_tail"]).cuda()
    want = torch.from_numpy(g["heatmaps"]).cuda()
    t = float(g["temperature"])
    assert np.array_equal(g["heatmaps"], g["forward_heatmaps"])
    assert torch.equal(pp.heatmap_tail(pre, t), want)

    class Carrier(torch.nn.Module):     # the reference head's member names; the layer stacks are identities here
        def __init__(self):
            super().__init__()
            self.deconv_layers = self.conv_layers = self.final_layer = torch.nn.Identity()
            self.temperature, self.normalize = t, None

    head = pp.patch_probmap_head(Carrier())
    assert torch.equal(head.forward_heatmap(pre), want)
    wl = synth.WORKLOADS[2]
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    a, b = pm.decode_device(pre, temperature=t), pm.decode_device(want)
    for key in ("locs", "vals", "argmax", "keypoints"):
        assert torch.equal(a[key], b[key]), key
    for bi in range(want.shape[0]):
        l, v, conv = oc.heatmap_expected_value(g["heatmaps"][bi], wl.sigmas, return_heatmap=True, conv="scipy")
        assert np.array_equal(a["argmax"][bi].cpu().numpy(), conv.reshape(17, -1).argmax(1))
        assert np.array_equal(a["vals"][bi].cpu().numpy(), v)


# --------------------------------------------------------------------------- records + peer mailbox
def test_pack_records_and_single_process_mailbox(pp):
    """pp_pack_records = the record tail of Codec.decode (codec.py:249-263); with a mailbox the same kernel publishes
    records + loss (one process plays every peer here; bench.py checks the multi-GPU exchange against an all-gather)."""
    from probpose_pytorch_b200.distributed import PeerMailbox
    wl = synth.WORKLOADS[2]
    B, K = 5, wl.num_keypoints
    W, H = wl.heatmap_size
    rng = np.random.default_rng(91)
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=92)
    maps = _oracle_encode("argmax", wl, kps, vis)["heatmaps"]
    pred = torch.from_numpy(synth.blob_predictions_numpy(maps, synth.blob_params(maps.shape[:2], 93), 94)).cuda()
    heads = [torch.from_numpy(rng.random((B, K, 1, 1), dtype=np.float32)).cuda() for _ in range(4)]
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    codec = pp.Codec(pm)
    rec = codec.decode_device((pred, *heads))
    dec = pm.decode_device(pred)
    assert rec.shape == (B, K, 7) and rec.dtype == torch.float64
    assert torch.equal(rec[..., 0:2], dec["keypoints"]) and torch.equal(rec[..., 2], dec["scores"].double())
    for c, h in zip((3, 4, 5), heads):
        assert torch.equal(rec[..., c], h.reshape(B, K).double())
    want_err = (heads[3].reshape(B, K) / float(np.sqrt(H ** 2 + W ** 2))).double()       # torch's float32 / scalar
    assert torch.equal(rec[..., 6], want_err)
    # mailbox: two slots, published twice each, every publication consumed (flow control) before the slot comes round
    mb = PeerMailbox(B, K, 2, pred.device)
    for rnd in range(2):
        for slot in range(2):
            scale = 1.0 + slot + 10 * rnd
            hs = [h * scale for h in heads]
            loss = torch.tensor(0.25 * scale, device="cuda")
            r = codec.decode_device((pred, *hs), mailbox=mb, slot=slot, loss=loss)
            got, got_loss = mb.read(slot)
            assert torch.equal(got.view(r.shape), r)
            assert got_loss.shape == (1,) and abs(got_loss.item() - 0.25 * scale) < 1e-6
    # the loss party may come first, on another stream, and from the loss' own finalize kernel
    side = torch.cuda.Stream()
    loss_fn = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)
    tgt = torch.from_numpy(maps).cuda()
    w = torch.from_numpy(vis).cuda()
    want_loss = loss_fn.forward_mean(pred, tgt, w)
    for order in ("loss-first", "records-first"):
        if order == "loss-first":
            got_l = loss_fn.forward_mean(pred, tgt, w, publish=mb.descriptor(0))
            mb.loss_enqueued(0)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            r = codec.decode_device((pred, *heads), mailbox=mb, slot=0)
        torch.cuda.current_stream().wait_stream(side)
        if order == "records-first":
            got_l = loss_fn.forward_mean(pred, tgt, w, publish=mb.descriptor(0))
            mb.loss_enqueued(0)
        got, got_loss = mb.read(0)
        assert torch.equal(got.view(r.shape), r) and torch.equal(got_l, want_loss)
        assert got_loss.item() == float(want_loss)
    # captured in a CUDA graph: the sequence numbers advance on the device, the host is told after each replay
    g = torch.cuda.CUDAGraph()
    static_heads = [h.clone() for h in heads]
    loss = torch.tensor(3.0, device="cuda")
    codec.decode_device((pred, *static_heads), mailbox=mb, slot=1, loss=loss)           # warm-up outside the graph
    mb.skip(1)
    with torch.cuda.graph(g):
        r = codec.decode_device((pred, *static_heads), mailbox=mb, slot=1, loss=loss)
    for k in range(3):
        static_heads[0].fill_(0.5 + k)
        g.replay()
        mb.published(1)
        got, got_loss = mb.read(1)
        assert torch.equal(got.view(r.shape), r) and float(got[0, 0, 3]) == 0.5 + k and got_loss.item() == 3.0


def test_mailbox_guards_readers_against_overwrites(pp, monkeypatch):
    """A block never changes under a reader.  With flow control a producer waits (bounded) for every consumer's
    acknowledgement before it rewrites a slot, and reports the consumer that never acknowledged; without it a reader
    that comes too late finds a newer sequence number and gets an error -- never silently mixed steps."""
    from probpose_pytorch_b200.distributed import PeerMailbox
    wl = synth.WORKLOADS[2]
    B, K = 2, wl.num_keypoints
    rng = np.random.default_rng(5)
    pred = torch.from_numpy(rng.random((B, K, 64, 48), dtype=np.float32)).cuda()
    heads = [torch.from_numpy(rng.random((B, K, 1, 1), dtype=np.float32)).cuda() for _ in range(4)]
    codec = pp.Codec(pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas))
    one = torch.tensor(1.0, device="cuda")
    # no flow control: two publications, the reader asks for the first one too late
    mb = PeerMailbox(B, K, 1, pred.device, flow_control=False)
    codec.decode_device((pred, *heads), mailbox=mb, slot=0, loss=one)
    r2 = codec.decode_device((pred, *[h * 2 for h in heads]), mailbox=mb, slot=0, loss=one)
    got, _ = mb.read(0)                                   # the latest publication is fine
    assert torch.equal(got.view(r2.shape), r2)
    mb._published[0] = 1                                  # a consumer whose book-keeping is one step behind
    with pytest.raises(RuntimeError, match="overwritten"):
        mb.read(0)
    # flow control: the second publication of an unconsumed slot runs into the acknowledgement time-out
    monkeypatch.setenv("PP_MAILBOX_ACK_TIMEOUT_US", "2000")
    mb = PeerMailbox(B, K, 1, pred.device)
    codec.decode_device((pred, *heads), mailbox=mb, slot=0, loss=one)
    codec.decode_device((pred, *heads), mailbox=mb, slot=0, loss=one)
    with pytest.raises(RuntimeError, match="did not acknowledge"):
        mb.check_async()
    mb = PeerMailbox(B, K, 1, pred.device, flow_control=False)
    mb._published[0] = 1      # the host believes a step was published; no kernel ever raised the flag
    with pytest.raises(RuntimeError, match="did not publish in time"):
        mb.read(0, timeout_us=2000)


@pytest.mark.parametrize("publish", ["in-step", "deferred"])
def test_mailbox_consumer_inside_the_step_graph(pp, publish):
    """The consumer side keeps its sequence numbers on the device (pp_mailbox_consume), so one CUDA graph per phase can
    publish slot g AND consume slot g - 2: every replay finds the records + loss of two steps ago in the consumer's
    private copy, nothing before anything was published, and the producers never run into the acknowledgement time-out
    (bench.py runs its multi-GPU steps exactly like this).  "deferred": the loss of a step is written to the mailbox's
    loss slot and published at the start of the NEXT step (pp_mailbox_commit_deferred), a no-op when nothing waits."""
    from probpose_pytorch_b200.distributed import PeerMailbox
    wl = synth.WORKLOADS[2]
    B, K, S = 3, wl.num_keypoints, 4
    rng = np.random.default_rng(17)
    pred = torch.from_numpy(rng.random((B, K, 64, 48), dtype=np.float32)).cuda()
    heads = [torch.from_numpy(rng.random((B, K, 1, 1), dtype=np.float32)).cuda() for _ in range(4)]
    codec = pp.Codec(pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas))
    mb = PeerMailbox(B, K, S, pred.device)
    loss = torch.zeros((), device="cuda")
    marker = heads[0]

    def phase(g):
        if publish == "deferred":
            mb.commit_deferred((g - 1) % S)                       # last step's loss; nothing on the very first step
            got = mb.read_async((g - 2) % S)
            r = codec.decode_device((pred, *heads), mailbox=mb, slot=g)      # records party only
            mb.loss_slot(g).copy_(loss.reshape(1))                            # "the loss kernel" of this step
            return r, got
        r = codec.decode_device((pred, *heads), mailbox=mb, slot=g, loss=loss)
        return r, mb.read_async((g - 2) % S)

    # with the deferred publication a slot becomes visible one step later, still before its consumer two steps on
    step = 0
    for g in range(S):                                   # eager cycle first (also the warm-up before capture)
        marker.fill_(step); loss.fill_(step)
        _, (rec, ls) = phase(g)
        if step >= 2:
            assert float(rec[0, 0, 3]) == step - 2 and float(ls[0]) == step - 2
        else:
            assert float(rec.abs().sum()) == 0.0          # nothing published into that slot yet: nothing consumed
        step += 1
    graphs = []
    for g in range(S):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = phase(g)
        graphs.append((gr, out))
    for _ in range(3 * S + 1):
        g = step % S
        marker.fill_(step); loss.fill_(step)
        graphs[g][0].replay()
        mb.published(g)
        r, (rec, ls) = graphs[g][1]
        assert float(r[0, 0, 3]) == step
        assert float(rec[0, 0, 3]) == step - 2 and float(ls[0]) == step - 2
        step += 1
    if publish == "deferred":
        mb.commit_deferred((step - 1) % S)                # flush the last step's loss
        mb.commit_deferred((step - 1) % S)                # ... and a second call finds nothing waiting
    mb.check_async()                                      # no consumer or producer time-out along the way
    for back in (2, 1):                                   # the two publications nobody has consumed yet
        rec, ls = mb.read((step - back) % S)
        assert float(rec[0, 0, 3]) == step - back and float(ls[0]) == step - back


def test_expected_decoder_kernel_selection(pp, monkeypatch):
    """pp_decode_expected_last_kernel reports the kernel that ran: the tensor-core kernel for the shapes it is built
    for; without it small batches take the CTA-per-heatmap kernel, batches with at least two heatmaps per resident warp
    the warp-per-heatmap kernel, odd shapes the generic one; all of them agree bit for bit."""
    from probpose_pytorch_b200 import _lib
    wl = synth.WORKLOADS[2]
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    last = _lib.lib().pp_decode_expected_last_kernel
    small = torch.rand(4, 17, 64, 48, device="cuda")
    m = pm.decode_device(small)
    assert last() == 5
    monkeypatch.setenv("PP_DECODE_MMA", "0")
    a = pm.decode_device(small)
    assert last() == 1
    monkeypatch.setenv("PP_DECODE_WARP", "1")
    b = pm.decode_device(small)
    assert last() == 2
    monkeypatch.delenv("PP_DECODE_WARP")
    for key in ("locs", "vals", "argmax", "keypoints"):
        assert torch.equal(a[key], b[key]) and torch.equal(a[key], m[key]), key
    big = torch.rand(256, 17, 64, 48, device="cuda") * 0.02
    big[:, :, 30, 20] = 1.0
    c = pm.decode_device(big)
    assert last() == 2
    monkeypatch.setenv("PP_DECODE_WARP", "0")
    d = pm.decode_device(big)
    assert last() == 1
    monkeypatch.delenv("PP_DECODE_WARP")
    monkeypatch.delenv("PP_DECODE_MMA")
    e = pm.decode_device(big)
    assert last() == 5
    for key in ("locs", "vals", "argmax", "keypoints"):
        assert torch.equal(c[key], d[key]) and torch.equal(c[key], e[key]), key
    l, v = pp.get_heatmap_expected_value(np.random.default_rng(3).random((5, 19, 27), dtype=np.float32), np.full(5, 0.07))
    assert last() == 4


# --------------------------------------------------------------------------- tensor-core prefilter + full batches
def _exact_conv(hm, sigmas, k):
    """float32 exact convolved map of ONE heatmap of channel k (scipy, the reference's call)."""
    return oc.heatmap_expected_value(hm[None], np.asarray(sigmas)[k:k + 1], return_heatmap=True, conv="scipy")[2][0]


@pytest.mark.parametrize("case", ["blob", "uniform", "range", "negative", "bf16", "c4"])
def test_mma_prefilter_error_bound(pp, case):
    """The tensor-core kernel only PROPOSES candidates; its rigorous bound |Z(p) - R(p) - const| <= kMmaErr in units of
    the scaled range (pp_decode_mma.cuh) is what guarantees that the exact argmax is among them.  The debug entry point
    returns Z; check the bound against the exact convolution on every pixel."""
    import ctypes as C
    from probpose_pytorch_b200 import _lib
    from probpose_pytorch_b200.heatmap import _oks_table
    wl = synth.WORKLOADS[4 if case == "c4" else 2]
    K, (W, H) = wl.num_keypoints, wl.heatmap_size
    rng = np.random.default_rng(77)
    B = 3
    if case in ("blob", "bf16", "c4"):
        kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=71)
        tgt = _oracle_encode("argmax", wl, synth.jitter_keypoints(wl, kps, 72), vis)["heatmaps"]
        maps = synth.blob_predictions_numpy(tgt, synth.blob_params(tgt.shape[:2], 73), 74)
    elif case == "uniform":
        maps = rng.random((B, K, H, W), dtype=np.float32)
    elif case == "range":      # values spread over 40 binary orders of magnitude, tiny and huge overall scales
        maps = (rng.random((B, K, H, W)) * np.exp2(rng.integers(-40, 1, size=(B, K, H, W)))).astype(np.float32)
        maps[0] *= np.float32(1e-20)
        maps[1] *= np.float32(1e20)
    else:                      # negative: raw logits, small variation on a large offset
        maps = rng.normal(0.0, 1.0, size=(B, K, H, W)).astype(np.float32)
        maps[1] = (100.0 + 1e-2 * maps[1]).astype(np.float32)
        maps[2] = -np.abs(maps[2])
    t = torch.from_numpy(maps).cuda()
    if case == "bf16":
        t = t.bfloat16()
        maps = t.float().cpu().numpy()
    tab = _oks_table(wl.sigmas, K, H, W, t.device)
    assert tab.mma_tables is not None
    p = _lib.DecodeParams(B, K, H, W, _lib.dtype_code(t.dtype), 0, 1.0, 0.0, 0.0)
    locs = torch.empty((B, K, 2), device="cuda")
    vals = torch.empty((B, K), device="cuda")
    arg = torch.full((B, K), -1, dtype=torch.int32, device="cuda")
    pre = torch.full((B, K, H, W), float("nan"), device="cuda")
    scratch = torch.zeros(int(_lib.lib().pp_decode_expected_scratch_bytes_for(p)) // 4 + 1, dtype=torch.int32, device="cuda")
    fn = _lib.lib().pp_debug_decode_mma_prefilter
    fn.argtypes = [C.POINTER(_lib.DecodeParams), C.POINTER(_lib.OksTable)] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
    rc = fn(p, tab.descriptor(), _lib.ptr(t), _lib.ptr(locs), _lib.ptr(vals), _lib.ptr(arg), _lib.ptr(pre), _lib.ptr(scratch),
            scratch.numel() * 4, _lib.stream_ptr(t.device))
    _lib.check(rc, "pp_debug_decode_mma_prefilter")
    torch.cuda.synchronize()
    pre = pre.cpu().numpy().astype(np.float64)
    handed = set(scratch[4:4 + int(scratch[2])].cpu().tolist())
    worst = 0.0
    for b in range(B):
        for k in range(K):
            if b * K + k in handed:     # plateaus / no dynamic range: decoded by the general kernel instead
                continue
            R = _exact_conv(maps[b, k], wl.sigmas, k).astype(np.float64)
            rng_ = float(maps[b, k].max()) - float(maps[b, k].min())
            unit = np.exp2(np.floor(np.log2(np.float32(rng_))) + 1)      # 1 / sc: the scaled range lies in [0.5, 1)
            diff = (pre[b, k] - R) / unit
            # the constant is (sum of taps)^2 - 1 times min h, plus float32 rounding of Z / sc + vmin: centre it
            diff -= 0.5 * (diff.max() + diff.min())
            f32 = 2.0 ** -22 * max(abs(float(maps[b, k].max())), abs(float(maps[b, k].min()))) / unit
            worst = max(worst, float(np.abs(diff).max()) - f32)
            assert np.abs(diff).max() <= 2.2e-3 + f32, (case, b, k, np.abs(diff).max())
            assert int(arg[b, k]) == int(R.astype(np.float32).argmax())
    assert case in ("range", "negative") or not handed
    print(f"{case}: worst proposal error {worst:.2e} of the 2.2e-3 bound; {len(handed)} handed over")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_expected_decoder_full_c2_batch_vs_oracle(pp, dtype, expected_kernel):
    """ALL 4 352 heatmaps of BASELINE config 2 (B = 256, bench.py's inputs) against the oracle's scipy convolution:
    argmax and scores bit-exact, coordinates within 1e-5 -- on the persistent multi-heatmap-per-warp paths the
    benchmark times."""
    if expected_kernel not in ("auto", "mma-grid2", "nomma", "team1-grid3", "cta"):
        pytest.skip("covered by the other kernels' runs")
    wl = synth.WORKLOADS[2]
    B, K = wl.batch, wl.num_keypoints
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1002)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    blob = am.encode_batch(synth.jitter_keypoints(wl, kps, seed=5000), vis)["heatmaps"]
    amp = torch.from_numpy(synth.blob_params((B, K), seed=6000)).cuda()
    g = torch.Generator(device="cuda").manual_seed(7)
    pred = (blob * amp[:, :, None, None] + torch.rand(blob.shape, device="cuda", generator=g) * 0.02).clamp_(0, 1)
    if dtype == "bf16":
        pred = pred.bfloat16()
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = pm.decode_device(pred)
    host = pred.float().cpu().numpy()
    arg, locs, vals = dev["argmax"].cpu().numpy(), dev["locs"].cpu().numpy(), dev["vals"].cpu().numpy()
    mism = 0
    for b in range(B):
        l, v, conv = oc.heatmap_expected_value(host[b], wl.sigmas, return_heatmap=True, conv="scipy")
        mism += int(np.count_nonzero(arg[b] != conv.reshape(K, -1).argmax(1)))
        np.testing.assert_allclose(locs[b], l, rtol=RTOL32, atol=1e-5)
        assert np.array_equal(vals[b], v)
    assert mism == 0, f"{mism} argmax mismatches in {B * K} heatmaps"


@pytest.mark.parametrize("cid,dtype", [(4, "fp32"), (4, "bf16"), (5, "fp32"), (5, "bf16")])
def test_expected_decoder_full_size_sampled_vs_oracle(pp, cid, dtype):
    """BASELINE configs 4 (8 704 heatmaps of 96x72) and 5 (68 096 heatmaps, K = 133) decoded at full size with the
    auto-selected kernel; 2 048 randomly chosen heatmaps of each are checked against the oracle (scipy convolution):
    argmax / scores bit-exact, coordinates 1e-5."""
    wl = synth.WORKLOADS[cid]
    B, K = wl.batch, wl.num_keypoints
    W, H = wl.heatmap_size
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1000 + cid)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    blob = am.encode_batch(synth.jitter_keypoints(wl, kps, seed=5000), vis)["heatmaps"]
    amp = torch.from_numpy(synth.blob_params((B, K), seed=6000)).cuda()
    g = torch.Generator(device="cuda").manual_seed(9)
    pred = (blob * amp[:, :, None, None]).add_(torch.rand(blob.shape, device="cuda", generator=g) * 0.02).clamp_(0, 1)
    del blob
    if dtype == "bf16":
        pred = pred.bfloat16()
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = pm.decode_device(pred)
    from probpose_pytorch_b200 import _lib
    assert _lib.lib().pp_decode_expected_last_kernel() == 5
    pick = np.random.default_rng(cid).choice(B * K, size=2048, replace=False)
    flat = pred.reshape(B * K, H, W)[torch.from_numpy(pick).cuda()].float().cpu().numpy()
    arg = dev["argmax"].reshape(-1).cpu().numpy()[pick]
    locs = dev["locs"].reshape(-1, 2).cpu().numpy()[pick]
    vals = dev["vals"].reshape(-1).cpu().numpy()[pick]
    sig = np.asarray(wl.sigmas)
    mism = 0
    for i, n in enumerate(pick):
        k = int(n % K)
        l, v, conv = oc.heatmap_expected_value(flat[i][None], sig[k:k + 1], return_heatmap=True, conv="scipy")
        mism += int(arg[i] != conv.reshape(-1).argmax())
        np.testing.assert_allclose(locs[i], l[0], rtol=RTOL32, atol=1e-5)
        assert vals[i] == v[0]
    assert mism == 0, f"{mism} argmax mismatches in 2048 sampled heatmaps of C{cid}"


def test_dark_and_loss_persistent_loops_vs_oracle(pp, monkeypatch):
    """decode_dark_fast_kernel and oks_loss_fast_kernel with the launch capped at a few CTAs (PP_DARK_GRID / PP_LOSS_GRID):
    every CTA walks >= 8 heatmaps / units through its work queue, TMA refill and double-buffered stages -- the loops
    bench.py times -- and must still match the oracle."""
    wl = synth.WORKLOADS[2]
    B, K = 6, wl.num_keypoints
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=311)
    tgt = _oracle_encode("argmax", wl, kps, vis)["heatmaps"]
    clean = _oracle_encode("argmax", wl, synth.jitter_keypoints(wl, kps, 312), np.ones_like(vis))["heatmaps"]
    monkeypatch.setenv("PP_DARK_GRID", "4")
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = am.decode_device(torch.from_numpy(clean).cuda())
    for b in range(B):
        kp_ref, sc_ref = oc.decode_argmax_dark(clean[b], wl.input_size, wl.heatmap_size, backend="cv2")
        peaks, _ = oc.heatmap_maximum(clean[b])
        assert np.array_equal(dev["peaks"][b].cpu().numpy(), peaks)
        assert np.array_equal(dev["scores"][b].cpu().numpy(), sc_ref[0])
        ok = peaks[:, 0] >= 0
        bound = RTOL32 * np.maximum(np.abs(kp_ref[0]), 1.0) + _dark_tolerance(clean[b], peaks, wl, RTOL32)
        assert (np.abs(dev["keypoints"][b].cpu().numpy() - kp_ref[0])[ok] <= bound[ok]).all()
    # loss: 102 heatmaps, G = 2 per unit -> 51 units over 3 CTAs = 17 units per CTA through both stages
    monkeypatch.setenv("PP_LOSS_GRID", "3")
    pred = synth.blob_predictions_numpy(clean, synth.blob_params(clean.shape[:2], 313), 314)
    for dt, rtol in ((torch.float32, RTOL32), (torch.bfloat16, RTOL16)):
        o_ref = torch.from_numpy(pred).to(dt).float().requires_grad_(True)
        t_ref = torch.from_numpy(tgt).to(dt).float()
        w = torch.from_numpy(vis)
        l_ref = oc.oks_heatmap_loss(o_ref, t_ref, w, per_pixel=True, smoothing_weight=0.05, oks_type="minus").mean()
        l_ref.backward()
        o = torch.from_numpy(pred).cuda().to(dt).requires_grad_(True)
        l = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus").forward_mean(
            o, torch.from_numpy(tgt).cuda().to(dt), w.cuda())
        l.backward()
        assert abs(l.item() - l_ref.item()) <= rtol * abs(l_ref.item()) + 1e-9
        _close(o.grad.float().cpu().numpy(), o_ref.grad.numpy(), rtol)


def test_backward_with_unit_upstream(pp):
    """loss.backward(gradient=unit_upstream(...)) = loss.backward() without the ones_like fill and the rescale launch: the
    gradient written by the fused forward pass is returned as is; any other upstream value still rescales it."""
    wl = synth.WORKLOADS[2]
    B, K = 3, wl.num_keypoints
    rng = np.random.default_rng(23)
    out = torch.from_numpy(rng.random((B, K, 64, 48), dtype=np.float32)).cuda()
    tgt = torch.from_numpy(rng.random((B, K, 64, 48), dtype=np.float32)).cuda()
    w = torch.from_numpy((rng.random((B, K)) < 0.8).astype(np.float32)).cuda()
    loss_fn = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)
    grads = []
    for how in ("plain", "unit", "scaled"):
        o = out.clone().requires_grad_(True)
        loss = loss_fn.forward_mean(o, tgt, w)
        if how == "plain":
            loss.backward()
        elif how == "unit":
            loss.backward(gradient=pp.unit_upstream(loss.device, loss.dtype))
        else:
            (loss * 3.0).backward()
        grads.append(o.grad)
    assert torch.equal(grads[0], grads[1])
    assert torch.allclose(grads[2], 3.0 * grads[0], rtol=1e-6, atol=0)
    assert float(pp.unit_upstream(out.device)) == 1.0


# --------------------------------------------------------------------------- encode-inside-loss (one pass, 2 H W e bytes)
@pytest.mark.parametrize("cid,batch,dtype,grid", [(2, 6, "fp32", 0), (2, 6, "fp32", 3), (3, 5, "bf16", 0), (4, 3, "fp32", 2),
                                                 (5, 1, "fp32", 4), (1, 4, "fp32", 0)])
def test_loss_with_encoded_target_matches_encode_then_loss(pp, cid, batch, dtype, grid, monkeypatch):
    """forward_mean_encoded = the reference's encode (codec.py:11-70, per sample) followed by
    OKSHeatmapLoss(per_pixel=True).mean() (loss.py:428-431) and its autograd gradient -- computed by ONE kernel that
    never materialises the target.  Checked against the oracle's encode + loss; `grid` caps the launch so that every
    CTA walks many units (its double-buffered factor tables and TMA stages)."""
    if grid:
        monkeypatch.setenv("PP_LOSS_GRID", str(grid))
    wl = synth.WORKLOADS[cid]
    K = wl.num_keypoints
    kps, vis, _ = synth.make_keypoints(wl, batch=batch, seed=400 + cid)
    if cid == 1:
        kps[0, 0] = (-5000.0, -5000.0)      # far outside: the float64 map underflows, weight 0
        vis[0, 1] = 0.4                     # unlabelled: zero target, weight = the visibility value
    enc = _oracle_encode("argmax", wl, kps, vis)
    tgt = enc["heatmaps"]
    pred = synth.blob_predictions_numpy(_oracle_encode("argmax", wl, synth.jitter_keypoints(wl, kps, 401), np.ones_like(vis))["heatmaps"],
                                        synth.blob_params(tgt.shape[:2], 402), 403)
    dt = torch.float32 if dtype == "fp32" else torch.bfloat16
    rtol = RTOL32 if dtype == "fp32" else RTOL16
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    mod = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus")
    for explicit in (False, True):
        w = (np.random.default_rng(9).random((batch, K)) < 0.8).astype(np.float32) if explicit else enc["keypoint_weights"].astype(np.float32)
        o_ref = torch.from_numpy(pred).to(dt).float().requires_grad_(True)
        l_ref = oc.oks_heatmap_loss(o_ref, torch.from_numpy(tgt).to(dt).float() if dtype == "fp32" else torch.from_numpy(tgt),
                                    torch.from_numpy(w), per_pixel=True, smoothing_weight=0.05, oks_type="minus").mean()
        l_ref.backward()
        o = torch.from_numpy(pred).cuda().to(dt).requires_grad_(True)
        loss, encoded = mod.forward_mean_encoded(o, am, torch.from_numpy(kps).cuda(), torch.from_numpy(vis).cuda(),
                                                 torch.from_numpy(w).cuda() if explicit else None, return_encoded=True)
        (loss * 3.0).backward()
        assert abs(loss.item() - l_ref.item()) <= rtol * abs(l_ref.item()) + 1e-9, (loss.item(), l_ref.item())
        _close(o.grad.float().cpu().numpy(), 3.0 * o_ref.grad.numpy(), rtol)
        assert np.array_equal(encoded["keypoint_weights"].cpu().numpy(), enc["keypoint_weights"].astype(np.float32))
        assert np.array_equal(encoded["in_image"].cpu().numpy(), enc["in_image"])
        assert np.array_equal(encoded["annotated"].cpu().numpy(), enc["annotated"])
        # same numbers as the two-kernel path of this package
        o2 = torch.from_numpy(pred).cuda().to(dt).requires_grad_(True)
        e2 = am.encode_batch(kps, vis, dtype=dt)
        l2 = mod.forward_mean(o2, e2["heatmaps"], torch.from_numpy(w).cuda() if explicit else e2["keypoint_weights"])
        l2.backward()
        assert abs(loss.item() - l2.item()) <= rtol * abs(l2.item()) + 1e-9
        _close(o.grad.float().cpu().numpy(), 3.0 * o2.grad.float().cpu().numpy(), rtol)


def test_c_consumer_runs_kernels_without_torch(tmp_path):
    """tests/c/cabi_gpu.c -- plain C + the CUDA runtime, no Python, no torch -- allocates with cudaMalloc and launches
    pp_encode -> pp_decode_expected (tensor-core kernel) -> pp_oks_loss_forward_encoded through the C ABI."""
    import shutil
    import subprocess
    from pathlib import Path
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from probpose_pytorch_b200.build import build
    so = build()
    root = Path(__file__).resolve().parents[1]
    gcc = shutil.which("gcc")
    cuda = Path("/usr/local/cuda")
    if gcc is None or not (cuda / "include" / "cuda_runtime.h").exists():
        pytest.skip("gcc / CUDA runtime headers not available")
    exe = tmp_path / "cabi_gpu"
    subprocess.run([gcc, "-std=c99", "-Wall", "-I", str(root / "include"), "-I", str(cuda / "include"),
                    str(root / "tests" / "c" / "cabi_gpu.c"), "-o", str(exe), str(so), f"-L{cuda / 'lib64'}", "-lcudart", "-lm",
                    f"-Wl,-rpath,{so.parent}", f"-Wl,-rpath,{cuda / 'lib64'}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout + res.stderr


def test_expected_decoder_maximum_on_the_map_border(pp, expected_kernel):
    """Noise-only maps whose convolved maximum lies on the first / last row or column (and in the corners): the reflect
    boundary, the last row pair of a band and the clipped bounding box all meet there.  (Found by comparing the kernels
    on all 68 096 heatmaps of C5: one noise map with its maximum in the bottom row.)"""
    wl = synth.WORKLOADS[2]
    K, (W, H) = wl.num_keypoints, wl.heatmap_size
    rng = np.random.default_rng(123)
    spots = [(H - 1, 28), (0, 17), (31, 0), (40, W - 1), (0, 0), (H - 1, W - 1), (H - 1, 0), (0, W - 1), (H - 2, 5), (1, W - 2),
             (H - 1, 1), (H - 1, W - 2)]
    maps = (rng.random((len(spots), K, H, W)) * 0.02).astype(np.float32)
    for i, (y, x) in enumerate(spots):
        maps[i, :, y, x] = 0.033                      # one raised pixel: raw and convolved maximum sit at the border
        if i % 2:
            maps[i, :, max(y - 1, 0):y + 2, max(x - 1, 0):x + 2] += 0.01
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    dev = pm.decode_device(torch.from_numpy(maps).cuda())
    bad = []
    for b in range(len(spots)):
        locs, vals, conv = oc.heatmap_expected_value(maps[b], wl.sigmas, return_heatmap=True, conv="scipy")
        am = conv.reshape(K, -1).argmax(1)
        got = dev["argmax"][b].cpu().numpy()
        for k in np.nonzero(got != am)[0]:
            bad.append((spots[b], int(k), int(got[k]), int(am[k])))
        if not bad:
            np.testing.assert_allclose(dev["locs"][b].cpu().numpy(), locs, rtol=RTOL32, atol=1e-5)
            assert np.array_equal(dev["vals"][b].cpu().numpy(), vals)
    assert not bad, f"(spot, channel, got, want): {bad[:12]}"


def test_expected_decoder_maximum_in_a_short_last_band(pp, golden_dir, expected_kernel):
    """tests/golden/decode_short_band.npz: the six heatmaps (of 408 576: C5, six seeds, fp32 + bf16) on which round 1's
    team kernel disagreed with the tensor-core kernel -- and with the oracle.  Out-of-image keypoints whose tail peaks
    in the bottom row over a noise floor: the pruned region spans almost the whole map, its last band holds a single
    row pair and the row-task split divided by one through a reciprocal that wraps to zero (div_magic)."""
    g = np.load(golden_dir / "decode_short_band.npz")
    wl = synth.WORKLOADS[5]
    K, (W, H) = wl.num_keypoints, wl.heatmap_size
    sig = np.asarray(wl.sigmas)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    for hm, k in zip(g["maps"], g["channel"]):
        stack = np.zeros((2, K, H, W), dtype=np.float32)       # the map in its channel, twice (B = 2)
        stack[:, k] = hm
        dev = pm.decode_device(torch.from_numpy(stack).cuda())
        l, v, conv = oc.heatmap_expected_value(hm[None], sig[k:k + 1], return_heatmap=True, conv="scipy")
        for b in range(2):
            assert int(dev["argmax"][b, k]) == int(conv.reshape(-1).argmax()), (int(k), b)
            np.testing.assert_allclose(dev["locs"][b, k].cpu().numpy(), l[0], rtol=RTOL32, atol=1e-5)
            assert float(dev["vals"][b, k]) == float(v[0])


@pytest.mark.parametrize("cid,dtype", [(2, "fp32"), (2, "bf16"), (4, "fp32"), (5, "fp32")])
def test_dark_decoder_full_size_tensor_core_vs_cta_kernel_and_oracle(pp, cid, dtype, monkeypatch):
    """The argmax + DARK-UDP decoder at BASELINE batch sizes through the tensor-core kernel: peaks / scores bit-equal to
    the CTA-per-heatmap kernel on every heatmap, refined coordinates equal to it up to the float32 blur's operation
    order (the two kernels evaluate the 11 x 11 blur rows-first / columns-first on flat maps), and a 512-heatmap sample
    of blob-shaped channels against the oracle (cv2) with the conditioning-aware bound of the golden tests."""
    wl = synth.WORKLOADS[cid]
    B, K = wl.batch, wl.num_keypoints
    W, H = wl.heatmap_size
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1000 + cid)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    blob = am.encode_batch(synth.jitter_keypoints(wl, kps, seed=5000), vis)["heatmaps"]
    amp = torch.from_numpy(synth.blob_params((B, K), seed=6000)).cuda()
    g = torch.Generator(device="cuda").manual_seed(11)
    pred = (blob * amp[:, :, None, None]).add_(torch.rand(blob.shape, device="cuda", generator=g) * 0.02).clamp_(0, 1)
    del blob
    if dtype == "bf16":
        pred = pred.bfloat16()
    from probpose_pytorch_b200 import _lib
    a = am.decode_device(pred)
    assert _lib.lib().pp_decode_argmax_dark_last_kernel() == 5
    handed = int(a["_scratch"][2])
    monkeypatch.setenv("PP_DARK_MMA", "0")
    b = am.decode_device(pred)
    assert _lib.lib().pp_decode_argmax_dark_last_kernel() == 1
    assert torch.equal(a["peaks"], b["peaks"]) and torch.equal(a["scores"], b["scores"])
    diff = (a["keypoints"] - b["keypoints"]).abs().reshape(B * K, 2).max(dim=1).values
    strong = (pred.reshape(B * K, -1).float().max(dim=1).values >= 0.1)
    # input px.  Blobs on a clean floor take the same row-then-column arithmetic in both kernels; a noise floor makes
    # the CTA kernel blur columns first (its full-plane path), and DARK's inverse Hessian amplifies the 1e-7 difference
    # (bfloat16 maps: quantised blobs have near-singular Hessians here and there, so only the bulk is compared)
    if dtype == "fp32":
        assert float(diff[strong].max()) <= 0.1, float(diff[strong].max())
    assert float((diff[strong] <= 1e-3).float().mean()) >= 0.98
    assert handed <= B * K // 100, f"{handed} heatmaps handed on"
    # oracle on a sample of blob-shaped heatmaps (whole samples: the oracle decodes (K, H, W) stacks)
    rng = np.random.default_rng(cid)
    for bi in rng.choice(B, size=max(1, 512 // K), replace=False):
        hm = pred[bi].float().cpu().numpy()
        kp, sc = oc.decode_argmax_dark(hm, wl.input_size, wl.heatmap_size, backend="cv2")
        peaks, _ = oc.heatmap_maximum(hm)
        assert np.array_equal(a["peaks"][bi].cpu().numpy(), peaks) and np.array_equal(a["scores"][bi].cpu().numpy(), sc[0])
        ok = (peaks[:, 0] >= 0) & (hm.reshape(K, -1).max(axis=1) >= 0.1)
        bound = RTOL32 * np.maximum(np.abs(kp[0]), 1.0) + _dark_tolerance(hm, peaks, wl, RTOL32)
        if dtype == "bf16":
            bound = bound + RTOL16
        assert (np.abs(a["keypoints"][bi].cpu().numpy() - kp[0])[ok] <= bound[ok]).all()
