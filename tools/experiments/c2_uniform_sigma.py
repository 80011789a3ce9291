"""Does the latency of the per-channel operand-table reads bound the tensor-core decoder at C2?  Same maps, decoded
once with the COCO sigmas (17 channels, ~10 distinct tables) and once with one sigma for every channel (one table, always
L1-resident after the first touch).  Also the grid-cap sweep (fewer warps per SM).  Usage: python tools/experiments/c2_uniform_sigma.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth
from decode_mma_probe import make
from decode_split import timed

wl = synth.WORKLOADS[2]
dev = torch.device("cuda")
pred = make(256, wl, dev)
for name, sig in (("coco", wl.sigmas), ("uniform", np.full_like(wl.sigmas, float(np.median(wl.sigmas))))):
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, sig)
    for grid in (0, 444, 296, 148):
        os.environ["PP_DECODE_GRID"] = str(grid)
        us, _ = timed(lambda: pm.decode_device(pred), iters=50)
        print(f"{name} sigmas, grid cap {grid or 'none'}: {us:.1f} us", flush=True)
