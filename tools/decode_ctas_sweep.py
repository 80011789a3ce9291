"""Expected-OKS decoder on the bench's mixed C2 predictions at 1..6 resident CTAs per SM (PP_DECODE_CTAS):
how the kernel scales with the number of heatmaps in flight per SM.  Usage: python tools/decode_ctas_sweep.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth
from decode_split import timed


def main():
    wl = synth.WORKLOADS[2]
    B, K = wl.batch, wl.num_keypoints
    dev = torch.device("cuda")
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1002)
    vis_d = torch.from_numpy(vis).to(dev)
    jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=5000)).to(dev)
    blob = am.encode_batch(jit, vis_d)["heatmaps"]
    amp = torch.from_numpy(synth.blob_params((B, K), seed=6000)).to(dev)
    pred = (blob * amp[:, :, None, None] + torch.rand_like(blob) * 0.02).clamp_(0, 1).contiguous()
    noise = (torch.rand_like(blob) * 0.02).contiguous()
    clean = (blob * amp[:, :, None, None]).contiguous()
    for mode in (1, 0):
        os.environ["PP_DECODE_WARP"] = "1" if mode else "0"
        os.environ["PP_DECODE_TEAM"] = str(max(mode, 1))
        for cap in (2, 0):
            os.environ["PP_DECODE_CTAS"] = str(cap)
            us, _ = timed(lambda: pm.decode_device(pred))
            print(f"TEAM={mode} PP_DECODE_CTAS={cap}: expected decode {us:7.1f} us", flush=True)
        os.environ["PP_DECODE_CTAS"] = "0"
        for name, t in (("noise only", noise), ("clean blobs", clean), ("zeros", torch.zeros_like(blob)),
                        ("mixed bf16", pred.bfloat16())):
            us, _ = timed(lambda: pm.decode_device(t))
            print(f"TEAM={mode} {name}: {us:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
