"""Expected-OKS decoder variants at B = 256 and B = 1024 (C2 shapes): separates the per-heatmap cost from the
launch / tail overhead.  Usage: python tools/decode_scale_probe.py [config_id]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth
from decode_split import timed


def make(B, wl, dev):
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1002)
    jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=5000)).to(dev)
    blob = am.encode_batch(jit, torch.from_numpy(vis).to(dev))["heatmaps"]
    amp = torch.from_numpy(synth.blob_params((B, wl.num_keypoints), seed=6000)).to(dev)
    return (blob * amp[:, :, None, None] + torch.rand_like(blob) * 0.02).clamp_(0, 1).contiguous()


def main():
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    wl = synth.WORKLOADS[cid]
    dev = torch.device("cuda")
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    preds = {B: make(B, wl, dev) for B in ((128, 256, 1024) if cid == 2 else (wl.batch // 4, wl.batch))}
    for mode in (1, 2, 0):
        os.environ["PP_DECODE_WARP"] = "1" if mode else "0"
        os.environ["PP_DECODE_TEAM"] = str(max(mode, 1))
        row = []
        for B, t in preds.items():
            us, _ = timed(lambda: pm.decode_device(t), iters=20)
            row.append(f"B={B}: {us:7.1f} us ({us * 1e3 / (B * wl.num_keypoints):5.2f} ns/hm)")
        print(f"TEAM={mode}  " + "  ".join(row), flush=True)


if __name__ == "__main__":
    main()
