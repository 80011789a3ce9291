"""Time the expected-OKS decoder on blob-only, noise-only and mixed heatmaps (per-heatmap cost of the
tile path vs the full path).  Usage: python tools/decode_split.py [config_id]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth


def timed(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3, keep


def main():
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    wl = synth.WORKLOADS[cid]
    B = min(wl.batch, 256)
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    kps, vis, _ = synth.make_keypoints(wl, batch=B)
    inside = synth.Workload(wl.config_id, "in", B, wl.num_keypoints, wl.input_size, wl.heatmap_size, False)
    kin, _, _ = synth.make_keypoints(inside, batch=B)
    dev = torch.device("cuda")
    ones = torch.ones((B, wl.num_keypoints), device=dev)
    clean = am.encode_batch(kin, ones)["heatmaps"]
    noise = torch.rand_like(clean) * 0.02
    blob_noise = (clean * 0.7 + noise).clamp_(0, 1)
    uniform = torch.rand_like(clean)
    zeros = torch.zeros_like(clean)
    n = B * wl.num_keypoints
    for name, t in (("clean blobs", clean), ("blobs + U(0,.02) noise", blob_noise), ("noise only U(0,.02)", noise),
                    ("uniform U(0,1)", uniform), ("all zero", zeros)):
        us, _ = timed(lambda: pm.decode_device(t))
        ud, _ = timed(lambda: am.decode_device(t))
        print(f"{name:26s} expected {us:8.1f} us {us * 1e3 / n:6.1f} ns/hm {n * t[0, 0].numel() * t.element_size() / us / 1e3:7.1f} GB/s"
              f" | dark {ud:8.1f} us {ud * 1e3 / n:6.1f} ns/hm {n * t[0, 0].numel() * t.element_size() / ud / 1e3:7.1f} GB/s")
    for k in (0, 5, 9, 11):
        sub = noise[:, k:k + 1].expand(-1, wl.num_keypoints, -1, -1).contiguous()
        # decode every channel with channel k's kernel by repeating sigma k
        pk = pp.ProbMap(wl.input_size, wl.heatmap_size, np.full(wl.num_keypoints, wl.sigmas[k]))
        us, _ = timed(lambda: pk.decode_device(sub))
        us2, _ = timed(lambda: pk.decode_device(clean))
        print(f"sigma[{k}]={wl.sigmas[k]:.3f}: noise-only {us * 1e3 / n:7.1f} ns/heatmap, clean {us2 * 1e3 / n:7.1f} ns/heatmap")


if __name__ == "__main__":
    main()
