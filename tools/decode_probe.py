"""Launch the expected-OKS decoder (and optionally DARK) a few times on the bench's C2 predictions (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth

dev = torch.device("cuda:0")
wl = synth.WORKLOADS[int(os.environ.get("PP_PROBE_CONFIG", "2"))]
B, K = wl.batch, wl.num_keypoints
am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1002)
jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=5000)).to(dev)
blob = am.encode_batch(jit, torch.from_numpy(vis).to(dev))["heatmaps"]
amp = torch.from_numpy(synth.blob_params((B, K), seed=6000)).to(dev)
pred = (blob * amp[:, :, None, None] + torch.rand_like(blob) * 0.02).clamp_(0, 1).contiguous()
if os.environ.get("PP_PROBE_NOISE"):   # flat / noisy maps only: the dense regime of the decoders
    pred = (torch.rand_like(blob) * 0.02).contiguous()
for _ in range(3):
    out = pm.decode_device(pred)
    if os.environ.get("PP_PROBE_DARK"):
        out2 = am.decode_device(pred)
torch.cuda.synchronize()
print("ok", float(out["keypoints"].sum()))
