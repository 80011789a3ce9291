"""argmax + DARK-UDP decode timings: tensor-core kernel (pp_decode_mma.cuh, kDark) against the CTA-per-heatmap kernel."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import _lib, synth
from decode_mma_probe import make
from decode_split import timed

dev = torch.device("cuda")
for cid, B in ((2, 256), (2, 1024), (4, 512), (5, 512)):
    wl = synth.WORKLOADS[cid]
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pred = make(B, wl, dev)
    n = B * wl.num_keypoints
    row = []
    for mma in ("1", "0"):
        os.environ["PP_DARK_MMA"] = mma
        us, out = timed(lambda: am.decode_device(pred), iters=20)
        torch.cuda.synchronize()
        row.append(f"kernel {_lib.lib().pp_decode_argmax_dark_last_kernel()} ({int(out['_scratch'][2]) if mma == '1' else 0} handed on): "
                   f"{us:8.1f} us {us * 1e3 / n:6.2f} ns/hm {n * pred[0, 0].numel() * 4 / us / 1e3:7.1f} GB/s")
    print(f"C{cid} B={B}: " + " | ".join(row), flush=True)
