"""Fused OKS loss forward+backward kernel under its tuning knobs (PP_LOSS_T rows per thread, PP_LOSS_G heatmaps per
CTA, PP_LOSS_STAGES) on C2- and C5-sized inputs.  Usage: python tools/loss_knob_sweep.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import _lib
from probpose_pytorch_b200.loss import _Prepared
from decode_split import timed


def main():
    dev = torch.device("cuda")
    loss_fn = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)
    for name, (B, K) in (("C2", (256, 17)), ("C5/4", (128, 133))):
        out = torch.rand(B, K, 64, 48, device=dev)
        tgt = torch.rand(B, K, 64, 48, device=dev)
        w = torch.ones(B, K, device=dev)
        for knobs in ({}, {"PP_LOSS_T": "16"}, {"PP_LOSS_T": "11"}, {"PP_LOSS_T": "6"}, {"PP_LOSS_G": "1"}, {"PP_LOSS_G": "3"},
                      {"PP_LOSS_T": "16", "PP_LOSS_G": "3"}, {"PP_LOSS_T": "16", "PP_LOSS_G": "4"}):
            for k in ("PP_LOSS_T", "PP_LOSS_G"):
                os.environ.pop(k, None)
            os.environ.update(knobs)
            prep = _Prepared(loss_fn, out, tgt, w, None, _lib.PP_LOSS_PIXEL_MEAN)
            us, _ = timed(lambda: prep.forward(want_grad=True), iters=30)
            print(f"{name} {knobs}: {us:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
