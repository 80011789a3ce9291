"""Launch the Sparsemax tail forward / backward a few times on bench-like logits (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth

dev = torch.device("cuda:0")
wl = synth.WORKLOADS[2]
B, K = wl.batch, wl.num_keypoints
am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1)
jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=51)).to(dev)
blob = am.encode_batch(jit, torch.from_numpy(vis).to(dev))["heatmaps"]
amp = torch.from_numpy(synth.blob_params((B, K), seed=61)).to(dev)
pred = (blob * amp[:, :, None, None] + torch.rand_like(blob) * 0.02).clamp_(0, 1)
logits = ((pred - 0.3) * 4.0).contiguous().requires_grad_(True)
up = torch.rand_like(pred)
for _ in range(3):
    logits.grad = None
    y = pp.heatmap_tail(logits, 0.5, normalize=1.0)
    y.backward(up)
torch.cuda.synchronize()
print("ok", float(y.sum()) / (B * K))
