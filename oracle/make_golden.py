"""Generate tests/golden/* by running the REFERENCE ITSELF (imported from /root/reference).

Run in the build container only (the reference mount does not exist on the GPU
box):  ``python -m oracle.make_golden``.  The fixtures pin the oracle
(tests/test_oracle_golden.py) and give the GPU parity tests reference outputs
that travel.  Nothing from the reference is copied: only its numerical outputs
on seeded synthetic inputs are stored.
"""

from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    sys.path.insert(0, REF)
    sys.path.insert(0, str(ROOT))
    from probpose import codec as rc, heatmap as rh, loss as rl  # the reference
    from probpose_pytorch_b200 import synth

    OUT.mkdir(parents=True, exist_ok=True)
    hashes: dict = {}

    # ---------------- encode: full arrays (small) ----------------
    wl = synth.WORKLOADS[3]
    kps, vis, _ = synth.make_keypoints(wl, batch=2)
    enc = {"keypoints": kps, "visible": vis}
    for kind, cls in (("probmap", rc.ProbMap), ("argmax", rc.ArgMaxProbMap)):
        codec = cls(wl.input_size, wl.heatmap_size, wl.sigmas)
        outs = [codec.encode(kps[b:b + 1], vis[b:b + 1]) for b in range(kps.shape[0])]
        enc[f"{kind}_heatmaps"] = np.stack([o["heatmaps"] for o in outs])
        enc[f"{kind}_weights"] = np.concatenate([o["keypoint_weights"] for o in outs])
        enc[f"{kind}_in_image"] = np.concatenate([o["in_image"] for o in outs])
        enc[f"{kind}_annotated"] = np.concatenate([o["annotated"] for o in outs])
    np.savez_compressed(OUT / "encode_small.npz", **enc)

    # ---------------- encode: hashes for the larger shapes ----------------
    for cid, batch in ((1, 4), (4, 2), (5, 1)):
        w = synth.WORKLOADS[cid]
        for dt in (np.float32, np.float64):
            k, v, _ = synth.make_keypoints(w, batch=batch, dtype=dt)
            for kind, cls in (("probmap", rc.ProbMap), ("argmax", rc.ArgMaxProbMap)):
                codec = cls(w.input_size, w.heatmap_size, w.sigmas)
                outs = [codec.encode(k[b:b + 1], v[b:b + 1]) for b in range(batch)]
                hm = np.stack([o["heatmaps"] for o in outs])
                wt = np.concatenate([o["keypoint_weights"] for o in outs])
                hashes[f"encode/C{cid}/B{batch}/{np.dtype(dt).name}/{kind}"] = {
                    "heatmaps_sha256": _sha(hm), "weights_sha256": _sha(wt.astype(np.float32)),
                    "sum": float(hm.astype(np.float64).sum()), "max": float(hm.max())}

    # ---------------- decoders ----------------
    wl = synth.WORKLOADS[3]
    kps, vis, _ = synth.make_keypoints(wl, batch=3, seed=77)
    am = rc.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pm = rc.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    jit = synth.jitter_keypoints(wl, kps, seed=78)
    blob_src = np.stack([am.encode(jit[b:b + 1], vis[b:b + 1])["heatmaps"] for b in range(3)])
    clean = np.stack([am.encode(kps[b:b + 1], np.ones_like(vis[b:b + 1]))["heatmaps"] for b in range(3)])
    amp = synth.blob_params(blob_src.shape[:2], seed=79)
    blob = synth.blob_predictions_numpy(blob_src, amp, seed=80)
    uni = synth.uniform_predictions_numpy((1,) + blob.shape[1:], seed=81)

    dec = {"blob": blob, "uniform": uni, "clean": clean}
    for name, arr in (("blob", blob), ("uniform", uni), ("clean", clean)):
        locs, vals, kp, sc = [], [], [], []
        for b in range(arr.shape[0]):
            l, v, conv = rh.get_heatmap_expected_value(arr[b].copy(), wl.sigmas, return_heatmap=True)
            k, s = pm.decode(arr[b])
            locs.append(l); vals.append(v); kp.append(k[0]); sc.append(s[0])
            if b == 0:
                dec[f"{name}_conv0"] = conv
                dec[f"{name}_argmax0"] = conv.reshape(conv.shape[0], -1).argmax(axis=1).astype(np.int32)
        dec[f"{name}_locs"] = np.stack(locs)
        dec[f"{name}_vals"] = np.stack(vals)
        dec[f"{name}_keypoints"] = np.stack(kp)
        dec[f"{name}_scores"] = np.stack(sc)
    dark = dec          # one file: inputs once, both decoders' outputs
    for name, arr in (("blob", blob), ("clean", clean)):
        kp, sc, pk = [], [], []
        for b in range(arr.shape[0]):
            k, s = am.decode(arr[b])
            p, _ = rh.get_heatmap_maximum(arr[b])
            kp.append(k[0]); sc.append(s[0]); pk.append(p)
        dark[f"{name}_dark_keypoints"] = np.stack(kp)
        dark[f"{name}_dark_scores"] = np.stack(sc)
        dark[f"{name}_peaks"] = np.stack(pk)
    np.savez_compressed(OUT / "decode.npz", **dec)

    # ---------------- loss ----------------
    g = torch.Generator().manual_seed(1234)
    B, K, H, W = 2, 3, 24, 20
    out = torch.rand(B, K, H, W, generator=g)
    tgt = torch.rand(B, K, H, W, generator=g)
    tgt[1, 2] = 0
    tw = (torch.rand(B, K, generator=g) < 0.8).float()
    mask = (torch.rand(B, 1, H, W, generator=g) < 0.9).float()
    loss = {"output": out.numpy(), "target": tgt.numpy(), "target_weights": tw.numpy(), "mask": mask.numpy()}
    variants = {
        "train": dict(smoothing_weight=0.05, oks_type="minus"),            # loss.py:348-352
        "both": dict(smoothing_weight=0.2, gaussian_weight=0.1, oks_type="both", loss_weight=2.0),
        "plus_skip": dict(oks_type="plus", skip_empty_channel=True),
    }
    for vname, kw in variants.items():
        mod = rl.OKSHeatmapLoss(use_target_weight=True, **kw)
        for wname, w, m in (("w", tw, None), ("wm", tw, mask), ("none", None, None)):
            for mname, mode in (("pixel", dict(per_pixel=True)), ("kpt", dict(per_keypoint=True)), ("mean", dict())):
                o = out.clone().requires_grad_(True)
                l = mod(o, tgt, w, m, **mode)
                red = l.mean() if mname == "pixel" else l.sum()
                red.backward()
                loss[f"{vname}/{wname}/{mname}/value"] = l.detach().numpy()
                loss[f"{vname}/{wname}/{mname}/grad"] = o.grad.numpy()
    np.savez_compressed(OUT / "loss.npz", **loss)

    # ---------------- training targets from heatmaps (ProbPoseLoss helpers) ----------------
    ploss = rl.ProbPoseLoss(rc.Codec(am))
    wgt = (np.random.default_rng(90).random((3, 17)) < 0.8).astype(np.int64)
    wgt[2] = 0                                       # a sample without any valid keypoint
    t_oks, t_w = ploss._oks_from_heatmaps(torch.from_numpy(clean), torch.from_numpy(blob), torch.from_numpy(wgt),
                                          heatmap_size=wl.heatmap_size)
    t_err = ploss._error_from_heatmaps(torch.from_numpy(clean), torch.from_numpy(blob))
    np.savez_compressed(OUT / "targets.npz", weight=wgt, oks=t_oks.numpy(), oks_weights=t_w.numpy(), error=t_err)

    # ---------------- known answers held by the reference's own tests ----------------
    codec = rc.ArgMaxProbMap((768, 768), (192, 192), np.array([0.1] * 20))
    enc = codec.encode(np.array([[[96.0, 96.0]]]), np.array([[1.0]]), np.array([[1.0]]))
    hm = enc["heatmaps"][None]
    zero_loss = rl.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus")(
        torch.zeros(hm.shape), torch.from_numpy(hm), torch.tensor([[1.0]]))
    known = {
        "tests/test_loss.py": {
            "heatmap_shape": list(hm.shape), "target_max": float(hm.max()), "zero_prediction_loss": float(zero_loss),
            "expected_target_max": 0.9970669150352478, "expected_loss": 0.0,
            "heatmap_sha256": _sha(hm)},
        "versions": {"numpy": np.__version__, "torch": torch.__version__},
    }
    import cv2, scipy
    known["versions"].update(cv2=cv2.__version__, scipy=scipy.__version__)
    (OUT / "hashes.json").write_text(json.dumps({"hashes": hashes, "known": known}, indent=1, sort_keys=True))
    for p in sorted(OUT.iterdir()):
        print(f"{p.name:28s} {p.stat().st_size / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
