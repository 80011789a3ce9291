"""Generate tests/golden/metrics.npz by running the REFERENCE's metric functions (imported from
/root/reference) on seeded inputs.  Build container only:  ``python -m oracle.make_golden_metrics``."""

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
    from probpose import loss as rl  # the reference

    g = np.load(OUT / "decode.npz")
    output, target = g["blob"], g["clean"]
    N, K, H, W = output.shape
    rng = np.random.default_rng(77)
    mask = rng.random((N, K)) < 0.8
    mask[:, 3] = False                      # a keypoint that is never valid -> acc = -1
    out = {"mask": mask}
    for thr in (0.05, 0.2):
        acc, avg, cnt = rl.pose_pck_accuracy(output.copy(), target.copy(), mask.copy(), thr=thr)
        out[f"pose/{thr}/acc"], out[f"pose/{thr}/avg"], out[f"pose/{thr}/cnt"] = acc, np.float64(avg), np.int64(cnt)
    norm64 = rng.uniform(20, 60, size=(N, 2))
    norm64[0, 0] = 0.0                      # masks the whole first instance
    norm64[1, 1] = -3.0                     # replaced by 1e6
    out["norm64"] = norm64
    acc, avg, cnt = rl.pose_pck_accuracy(output.copy(), target.copy(), mask.copy(), thr=0.1, normalize=norm64.copy())
    out["pose/norm64/acc"], out["pose/norm64/avg"], out["pose/norm64/cnt"] = acc, np.float64(avg), np.int64(cnt)
    # coordinates directly, float32 normalisation (float32 arithmetic in the reference)
    pred = rng.uniform(0, 48, size=(N, K, 2)).astype(np.float32)
    gt = (pred + rng.normal(0, 2.0, size=pred.shape)).astype(np.float32)
    norm32 = np.full((N, 2), 40.0, dtype=np.float32)
    out["pred"], out["gt"], out["norm32"] = pred, gt, norm32
    acc, avg, cnt = rl.keypoint_pck_accuracy(pred.copy(), gt.copy(), mask.copy(), 0.05, norm32.copy())
    out["kpt/acc"], out["kpt/avg"], out["kpt/cnt"] = acc, np.float64(avg), np.int64(cnt)
    none = np.zeros_like(mask)
    acc, avg, cnt = rl.keypoint_pck_accuracy(pred.copy(), gt.copy(), none, 0.05, norm32.copy())
    out["kpt_none/acc"], out["kpt_none/avg"], out["kpt_none/cnt"] = acc, np.float64(avg), np.int64(cnt)
    # mask-select metrics of ProbPoseLoss (self is unused by both)
    dt = rng.random((N, K)).astype(np.float32)
    gtb = (rng.random((N, K)) < 0.6).astype(np.float32)
    out["scalar_dt"], out["scalar_gt"] = dt, gtb
    a, t = rl.ProbPoseLoss.get_binary_accuracy(None, torch.from_numpy(dt), torch.from_numpy(gtb), torch.from_numpy(mask))
    out["binary/acc"], out["binary/thr"] = a.numpy(), t.numpy()
    out["mae"] = rl.ProbPoseLoss.get_mae(None, torch.from_numpy(dt), torch.from_numpy(gtb), torch.from_numpy(mask)).numpy()
    np.savez_compressed(OUT / "metrics.npz", **out)
    print("wrote", OUT / "metrics.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
