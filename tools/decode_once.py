"""A handful of expected-OKS decodes of one BASELINE configuration (for an ncu capture of one launch).
Usage: python tools/decode_once.py <config> <batch> [dark]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth
from decode_mma_probe import make

cid, B = int(sys.argv[1]), int(sys.argv[2])
wl = synth.WORKLOADS[cid]
dev = torch.device("cuda")
cls = pp.ArgMaxProbMap if len(sys.argv) > 3 else pp.ProbMap
pm = cls(wl.input_size, wl.heatmap_size, wl.sigmas)
pred = make(B, wl, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    flush.zero_()
    out = pm.decode_device(pred)
torch.cuda.synchronize()
print("ok", tuple(out["keypoints"].shape))
