"""Generate tests/golden/probpose_loss.npz: the REFERENCE's ``ProbPoseLoss.forward`` (imported from
/root/reference, CPU) on seeded inputs.  Build container only:  ``python -m oracle.make_golden_probpose_loss``."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"


def main() -> None:
    sys.path.insert(0, REF)
    sys.path.insert(0, str(ROOT))
    from probpose import codec as rc, loss as rl  # the reference
    from probpose_pytorch_b200 import synth

    wl = synth.WORKLOADS[3]
    B = 4
    kps, vis, visibility = synth.make_keypoints(wl, batch=B, seed=311)
    am = rc.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    enc = [am.encode(kps[b:b + 1], vis[b:b + 1]) for b in range(B)]
    gt = {
        "heatmaps": np.stack([e["heatmaps"] for e in enc]),
        "in_image": np.concatenate([e["in_image"] for e in enc]),
        "keypoints_visible": np.concatenate([e["annotated"] for e in enc]),
        "keypoints_visibility": (visibility > 0.5),
    }
    jit = synth.jitter_keypoints(wl, kps, seed=312)
    src = np.stack([am.encode(jit[b:b + 1], np.ones_like(vis[b:b + 1]))["heatmaps"] for b in range(B)])
    amp = synth.blob_params(src.shape[:2], seed=313)
    rng = np.random.default_rng(314)
    # predictions: blobs on a faint smooth floor (DARK on pure noise is ill-conditioned; see tests/test_gpu_parity.py)
    dt_heatmaps = np.clip(src * amp[:, :, None, None] + 0.002, 0, 1).astype(np.float32)
    heads = [rng.uniform(0.05, 0.95, size=(B, 17, 1, 1)).astype(np.float32) for _ in range(3)]
    heads.append(rng.uniform(0.0, 6.0, size=(B, 17, 1, 1)).astype(np.float32))
    kw = (rng.random((B, 17)) < 0.85).astype(np.float32)
    out = {f"gt/{k}": v for k, v in gt.items()}
    out.update({"dt_heatmaps": dt_heatmaps, "dt_probs": heads[0], "dt_vis": heads[1], "dt_oks": heads[2], "dt_errs": heads[3],
                "keypoint_weights": kw})
    gt_t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in gt.items()}
    for name, freeze, kwargs in (("frozen", True, {}), ("live", False, {}),
                                 ("zeros", False, dict(learn_heatmaps_from_zeros=True)),
                                 ("weights", True, dict(keypoint_weights=torch.from_numpy(kw)))):
        mod = rl.ProbPoseLoss(rc.Codec(am), freeze_error=freeze)
        pred = [torch.from_numpy(x).clone().requires_grad_(True) for x in (dt_heatmaps, *heads)]
        np.random.seed(99)
        losses, acc = mod(gt_t, tuple(pred), compute_acc=True, **kwargs)
        total = sum(losses.values())
        total.backward()
        for k, v in losses.items():
            out[f"{name}/loss/{k}"] = v.detach().numpy()
        for k, v in acc.items():
            out[f"{name}/acc/{k}"] = np.asarray(v.detach().numpy() if isinstance(v, torch.Tensor) else v)
        for n, p in zip(("heatmaps", "probs", "vis", "oks", "errs"), pred):
            out[f"{name}/grad/{n}"] = p.grad.numpy()
    np.savez_compressed(OUT / "probpose_loss.npz", **out)
    print({k: (np.asarray(v).shape, float(np.asarray(v).ravel()[0])) for k, v in out.items() if "/loss/" in k or "/acc/" in k})


if __name__ == "__main__":
    main()
