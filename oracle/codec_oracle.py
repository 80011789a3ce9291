"""CPU oracle: ProbPose codec + heatmap ops (encode, expected-OKS decode, argmax+DARK decode).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restated from the
behaviour of the reference at ``/root/reference`` -- every function cites the
file:line it follows.  Written vectorised over the keypoint axis; the
reference's per-sample (N == 1) semantics are kept, batched helpers simply loop
over samples (reference decoders are single-sample: heatmap.py:364,379,
codec.py:339).

dtype flow matters for bit parity and is kept as in the reference under
NumPy-2 promotion rules: integer pixel grids minus float32 keypoints are
float64 (codec.py:56-59); maps are stored as float32 (codec.py:45,69);
convolution accumulates in double and stores float32 (heatmap.py:335,362-364);
sub-pixel refinement is float32 (heatmap.py:136-165); DARK derivatives are
float32, the 2x2 solve is float64 (codec.py:361-373); final coordinates are
float64 (codec.py:237,541).
"""

from __future__ import annotations

import numpy as np

COCO17_SIGMAS = np.array(
    [0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072,
     0.062, 0.062, 0.107, 0.107, 0.087, 0.087, 0.089, 0.089]
)


# --------------------------------------------------------------------------- #
# encode
# --------------------------------------------------------------------------- #
def oks_variance_table(sigmas, H: int, W: int) -> np.ndarray:
    """Per-keypoint OKS variance ``s`` used as ``exp(-d^2 / (2 s))``.

    Follows codec.py:48,60-62 and heatmap.py:171,176-178:
    ``s = clip((2 sigma_k)^2 * sqrt(H/1.25 * W/1.25) * 2, 0.55, 3.0)``.
    """
    sig = np.asarray(sigmas)
    area = np.sqrt(H / 1.25 * W / 1.25)
    s = (sig * 2) ** 2 * area * 2
    return np.clip(s, 0.55, 3.0)


def generate_probmaps(heatmap_size, keypoints, keypoints_visible, sigmas, sigma=0.55):
    """OKS-shaped probability-map targets (codec.py:11-70).

    ``keypoints`` are in heatmap space, shape (N, K, 2); returns
    ``(heatmaps (K, H, W) float32, keypoint_weights (N, K))``.  Unlabelled
    keypoints (visible < 0.5, codec.py:53) leave a zero channel and keep their
    weight; labelled ones paint the full grid with no window truncation
    (codec.py:56-66) and get weight ``max(map) > 0`` in float64 (codec.py:68).
    """
    W, H = heatmap_size
    N, K, _ = keypoints.shape
    maps = np.zeros((K, H, W), dtype=np.float32)
    weights = keypoints_visible.copy()

    if sigma is not None and sigma > 0:          # codec.py:63-64
        two_s = np.full(K, 2 * sigma)
    else:
        two_s = 2 * oks_variance_table(np.asarray(sigmas)[:K], H, W)

    ys = np.arange(H).reshape(1, H, 1)
    xs = np.arange(W).reshape(1, 1, W)
    for n in range(N):
        labelled = ~(keypoints_visible[n] < 0.5)
        if not labelled.any():
            continue
        kx = keypoints[n, :, 0].reshape(K, 1, 1)
        ky = keypoints[n, :, 1].reshape(K, 1, 1)
        dx = xs - kx                              # int64 - float32 -> float64
        dy = ys - ky
        dist = np.sqrt(dx ** 2 + dy ** 2)
        m = np.exp(-(dist ** 2 / two_s.reshape(K, 1, 1)))
        peak_positive = (m.reshape(K, -1).max(axis=1) > 0).astype(int)
        for k in np.nonzero(labelled)[0]:
            maps[k] = m[k]
            weights[n, k] = peak_positive[k]
    return maps, weights


def encode(kind, input_size, heatmap_size, sigmas, keypoints, keypoints_visible=None,
           id_similarity=0.0, sigma=None):
    """``ProbMap.encode`` (codec.py:138-212) / ``ArgMaxProbMap.encode`` (codec.py:443-513).

    ``kind`` is ``"probmap"`` (default sigma 2.0, codec.py:122; also returns
    ``heatmap_keypoints``, codec.py:208) or ``"argmax"`` (default sigma -1,
    codec.py:426).  ``keypoints`` (1, K, 2) are in input-image space.
    """
    assert keypoints.shape[0] == 1, "only single-instance keypoint encoding"
    if sigma is None:
        sigma = 2.0 if kind == "probmap" else -1
    scale = ((np.array(input_size) - 1) / (np.array(heatmap_size) - 1)).astype(np.float32)
    if keypoints_visible is None:
        keypoints_visible = np.ones(keypoints.shape[:2], dtype=np.float32)
    maps, weights = generate_probmaps(heatmap_size, keypoints / scale, keypoints_visible, sigmas, sigma)
    x, y = keypoints[:, :, 0], keypoints[:, :, 1]
    in_image = (x >= 0) & (x < input_size[0]) & (y >= 0) & (y < input_size[1])
    out = dict(
        heatmaps=maps,
        keypoint_weights=weights,
        annotated=keypoints_visible > 0,
        in_image=in_image,
        keypoints_scaled=keypoints,
        identification_similarity=id_similarity,
    )
    if kind == "probmap":
        out["heatmap_keypoints"] = keypoints / scale
    return out


# --------------------------------------------------------------------------- #
# expected-OKS decoder
# --------------------------------------------------------------------------- #
def oks_kernels_2d(K: int, H: int, W: int, sigmas) -> list:
    """Normalised per-keypoint OKS kernels, float64 (heatmap.py:170-194).

    radius = ceil(3 s), diameter = 2 r + 1, ``exp(-dist^2/(2 s)) / sum``; the
    reference squares a square root (heatmap.py:187-188), kept for bit parity.
    """
    s = oks_variance_table(np.asarray(sigmas)[:K], H, W)
    out = []
    for k in range(K):
        r = int(np.ceil(s[k] * 3))
        ax = np.arange(2 * r + 1) - r
        gx, gy = np.meshgrid(ax, ax)
        dist = np.sqrt(gx ** 2 + gy ** 2)
        ker = np.exp(-(dist ** 2) / (2 * s[k]))
        out.append(ker / ker.sum())
    return out


def conv_reflect_f64(hm: np.ndarray, ker: np.ndarray) -> np.ndarray:
    """Restatement of ``scipy.ndimage.convolve(hm, ker, mode='reflect')`` as called
    at heatmap.py:361-362 (third-party; scipy is not vendored in the reference).

    Published algorithm (scipy ``NI_Correlate``): half-sample-symmetric
    extension (d c b a | a b c d | d c b a), for every output pixel a double
    accumulator runs over the kernel footprint in row-major order
    (``tmp += w * value``), and the result is stored in the input dtype
    (float32).  The kernel is symmetric, so convolution == correlation.
    """
    H, W = hm.shape
    d = ker.shape[0]
    r = d // 2
    pad = np.pad(hm.astype(np.float64), r, mode="symmetric")
    kf = ker[::-1, ::-1]
    acc = np.zeros((H, W), dtype=np.float64)
    for i in range(d):
        for j in range(d):
            acc += kf[i, j] * pad[i:i + H, j:j + W]
    return acc.astype(np.float32)


def subpixel_quadratic(conv: np.ndarray, locs: np.ndarray) -> np.ndarray:
    """1-D quadratic peak fit in x and y on the convolved maps, float32
    (heatmap.py:114-167).  Border peaks are left alone (heatmap.py:120-125),
    zero curvature divides by 1e-6 (heatmap.py:156-157), the shift is not clamped.
    """
    K, H, W = conv.shape
    x = locs[:, 0].astype(np.int32)
    y = locs[:, 1].astype(np.int32)
    ok = (x > 0) & (x < W - 1) & (y > 0) & (y < H - 1)
    out = locs.copy()
    if ok.any():
        kk = np.nonzero(ok)[0]
        xv, yv = x[ok], y[ok]
        c = conv[kk, yv, xv]
        l, r = conv[kk, yv, xv - 1], conv[kk, yv, xv + 1]
        u, dn = conv[kk, yv - 1, xv], conv[kk, yv + 1, xv]
        gx = (r - l) / 2.0
        gy = (dn - u) / 2.0
        hxx = r + l - 2 * c
        hyy = dn + u - 2 * c
        hxx = np.where(hxx != 0, hxx, 1e-6)
        hyy = np.where(hyy != 0, hyy, 1e-6)
        out[ok, 0] += -gx / hxx
        out[ok, 1] += -gy / hyy
    return out


def heatmap_expected_value(heatmaps: np.ndarray, sigmas, return_heatmap=False, conv="numpy"):
    """Expected-OKS decoder core for one sample (heatmap.py:291-395).

    ``heatmaps`` (K, H, W) float32.  Returns ``locs (K, 2) float32`` (argmax of
    the OKS-convolved map, first maximum wins, plus sub-pixel shift),
    ``vals (K,) float32`` read from the *unconvolved* map at the integer argmax
    (heatmap.py:375-379) and optionally the convolved maps.  ``conv='scipy'``
    calls the same third-party routine the reference calls; ``'numpy'`` uses
    the restatement above.
    """
    assert heatmaps.ndim == 3
    K, H, W = heatmaps.shape
    kernels = oks_kernels_2d(K, H, W, sigmas)
    out = np.zeros_like(heatmaps)
    for k in range(K):
        if conv == "scipy":
            from scipy.ndimage import convolve
            out[k] = convolve(heatmaps[k], kernels[k], mode="reflect")
        else:
            out[k] = conv_reflect_f64(heatmaps[k], kernels[k])
    flat = out.reshape(K, H * W).argmax(axis=1)
    ys, xs = np.unravel_index(flat, (H, W))
    locs = np.stack((xs, ys), axis=-1).astype(np.float32)
    locs = subpixel_quadratic(out, locs)
    vals = heatmaps[np.arange(K), ys, xs]
    if return_heatmap:
        return locs, vals, out
    return locs, vals


def decode_expected(heatmaps, input_size, heatmap_size, sigmas, conv="numpy"):
    """``ProbMap.decode`` for one sample (codec.py:214-239): expected-OKS core,
    then ``/ [W-1, H-1] * input_size`` in float64 (not the inverse of encode's
    ``(in-1)/(hm-1)``; the asymmetry is the reference's)."""
    W, H = heatmap_size
    locs, vals = heatmap_expected_value(heatmaps.copy(), sigmas, conv=conv)
    kps = locs[None] / [W - 1, H - 1] * input_size
    return kps, vals[None]


# --------------------------------------------------------------------------- #
# argmax + DARK-UDP decoder
# --------------------------------------------------------------------------- #
def heatmap_maximum(heatmaps: np.ndarray):
    """Flat argmax / max per heatmap; location -1 where max <= 0 (heatmap.py:13-52)."""
    assert heatmaps.ndim in (3, 4)
    H, W = heatmaps.shape[-2:]
    lead = heatmaps.shape[:-2]
    flat = heatmaps.reshape(-1, H * W)
    idx = flat.argmax(axis=1)
    ys, xs = np.unravel_index(idx, (H, W))
    locs = np.stack((xs, ys), axis=-1).astype(np.float32)
    vals = flat.max(axis=1)
    locs[vals <= 0.0] = -1
    return locs.reshape(lead + (2,)), vals.reshape(lead)


def gaussian_taps_f32(ksize: int = 11) -> np.ndarray:
    """``cv2.getGaussianKernel(ksize, 0, CV_32F)`` restated (third-party; opencv-python
    4.11.0.86 pinned in requirements.txt:5): sigma = 0.3*((ksize-1)*0.5-1)+0.8,
    taps exp(-x^2/(2 sigma^2)) normalised in double, then rounded to float32."""
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (t / t.sum()).astype(np.float32)


def blur_zero_pad_f32(hm: np.ndarray, ksize: int = 11) -> np.ndarray:
    """The blur of codec.py:306-311 for one map: zero-pad by (ksize-1)/2, 11x11
    Gaussian, crop.  Because of the explicit zero padding cv2's own border mode
    never reaches the cropped region.  cv2 filters separably in float32 (row
    pass, then column pass); so does this."""
    H, W = hm.shape
    b = (ksize - 1) // 2
    taps = gaussian_taps_f32(ksize)
    pad = np.zeros((H + 2 * b, W + 2 * b), dtype=np.float32)
    pad[b:b + H, b:b + W] = hm
    rows = np.zeros((H + 2 * b, W), dtype=np.float32)
    for j in range(ksize):
        rows += taps[j] * pad[:, j:j + W]
    out = np.zeros((H, W), dtype=np.float32)
    for i in range(ksize):
        out += taps[i] * rows[i:i + H, :]
    return out


def gaussian_blur_modulate(heatmaps: np.ndarray, ksize: int = 11, backend="numpy") -> np.ndarray:
    """``gaussian_blur`` (codec.py:284-313): blur every map and rescale it so its
    maximum is preserved (``*= max_before / (max_after + 1e-12)``, float32).
    In place, like the reference."""
    assert ksize % 2 == 1
    b = (ksize - 1) // 2
    K, H, W = heatmaps.shape
    for k in range(K):
        top = np.max(heatmaps[k])
        if backend == "cv2":
            import cv2
            pad = np.zeros((H + 2 * b, W + 2 * b), dtype=np.float32)
            pad[b:-b, b:-b] = heatmaps[k]
            blurred = cv2.GaussianBlur(pad, (ksize, ksize), 0)[b:-b, b:-b].copy()
        else:
            blurred = blur_zero_pad_f32(heatmaps[k], ksize)
        heatmaps[k] = blurred
        heatmaps[k] *= top / (np.max(heatmaps[k]) + 1e-12)
    return heatmaps


def dark_udp_refine(locs: np.ndarray, heatmaps: np.ndarray, ksize: int = 11, backend="numpy"):
    """``refine_keypoints_dark_udp`` (codec.py:315-375) for N == 1.

    ``locs`` (1, K, 2) float32 integer peaks, modified in place; ``heatmaps``
    (K, H, W) float32, modified in place (blur, clip to [1e-3, 50], log).  The
    seven neighbours are gathered from the edge-padded, *flattened* stack
    exactly like the reference (codec.py:346-359), which includes its
    behaviour for the (-1,-1) sentinel of empty channels (reads run into the
    neighbouring channel; Appendix B-7 of SURVEY.md).
    """
    N, K = locs.shape[:2]
    H, W = heatmaps.shape[1:]
    heatmaps = gaussian_blur_modulate(heatmaps, ksize, backend)
    np.clip(heatmaps, 1e-3, 50.0, heatmaps)
    np.log(heatmaps, heatmaps)
    flat = np.pad(heatmaps, ((0, 0), (1, 1), (1, 1)), mode="edge").flatten()
    row = W + 2
    for n in range(N):
        at = locs[n, :, 0] + 1 + (locs[n, :, 1] + 1) * row
        at += row * (H + 2) * np.arange(0, K)
        at = at.astype(int).reshape(-1, 1)
        c = flat[at]
        xp, xm = flat[at + 1], flat[at - 1]
        yp, ym = flat[at + row], flat[at - row]
        pp, mm = flat[at + row + 1], flat[at - row - 1]
        g = np.concatenate([0.5 * (xp - xm), 0.5 * (yp - ym)], axis=1).reshape(K, 2, 1)
        hxx = xp - 2 * c + xm
        hyy = yp - 2 * c + ym
        hxy = 0.5 * (pp - xp - yp + c + c - xm - ym + mm)
        hess = np.concatenate([hxx, hxy, hxy, hyy], axis=1).reshape(K, 2, 2)
        inv = np.linalg.pinv(hess + np.finfo(np.float32).eps * np.eye(2))
        locs[n] -= np.einsum("imn,ink->imk", inv, g).squeeze()
    return locs


def decode_argmax_dark(heatmaps, input_size, heatmap_size, ksize: int = 11, backend="numpy"):
    """``ArgMaxProbMap.decode`` for one sample (codec.py:515-543)."""
    W, H = heatmap_size
    hm = heatmaps.copy()
    peaks, scores = heatmap_maximum(hm)
    kps = dark_udp_refine(peaks[None].copy(), hm, ksize, backend)
    kps = kps / [W - 1, H - 1] * input_size
    return kps, scores[None]


# --------------------------------------------------------------------------- #
# head tail
# --------------------------------------------------------------------------- #
def head_tail(x: np.ndarray, temperature: float = 0.5) -> np.ndarray:
    """Post-conv tail of ``ProbMapHead.forward_heatmap`` with ``normalize=None``
    (head.py:526-532): ``clamp(x / temperature, 0, 1)`` in float32."""
    return np.clip(x.astype(np.float32) / np.float32(temperature), np.float32(0), np.float32(1))
