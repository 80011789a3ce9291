"""Developer tool: per-phase cycle breakdown of the expected-OKS decode kernel.
Builds a separate library with -DPP_PHASE_TIMING (does not touch the product .so) and runs it."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
CSRC = ROOT / "probpose_pytorch_b200" / "csrc"
DBG = Path("/tmp/libprobpose_b200_timing.so")


def build():
    from probpose_pytorch_b200.build import SOURCES
    srcs = [str(CSRC / f) for f in SOURCES]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DPP_PHASE_TIMING",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-shared", "-cudart", "static",
           "-o", str(DBG), *srcs]
    subprocess.run(cmd, check=True)


def main():
    if not DBG.exists() or "--rebuild" in sys.argv:
        build()
    from probpose_pytorch_b200 import _lib
    _lib.LIB_PATH = DBG
    import probpose_pytorch_b200 as pp
    from probpose_pytorch_b200 import synth
    L = _lib.lib()
    wl = synth.WORKLOADS[2]
    B = 256
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    inside = synth.Workload(2, "in", B, 17, wl.input_size, wl.heatmap_size, False)
    kin, _, _ = synth.make_keypoints(inside, batch=B)
    dev = torch.device("cuda")
    clean = am.encode_batch(kin, torch.ones((B, 17), device=dev))["heatmaps"]
    noise = torch.rand_like(clean) * 0.02
    import os
    forced = os.environ.get("PP_DECODE_WARP", "1")
    os.environ["PP_DECODE_WARP"] = forced           # the team kernel unless PP_DECODE_WARP=0 asks for the CTA kernel
    if forced != "0":
        names = ["top/tables", "TMA wait", "scan A (+publish)", "B: p0+L", "C: bbox", "column pass", "row pass", "candidates",
                 "exact values", "-", "outputs", "end sync + TMA issue"]
    else:
        names = ["top/tables", "TMA wait", "scan A", "B: p0+L", "C: bbox", "D: gather/col", "E/F: prefilter", "candidates",
                 "G: exact cands", "neighbours", "H: outputs", "final barrier"]
    buf = (C.c_ulonglong * 16)()
    for label, t in (("clean blobs (tile path)", clean), ("noise only (full path)", noise), ("zeros", torch.zeros_like(clean))):
        pm.decode_device(t)
        L.pp_debug_phase_cycles(buf, 1)
        for _ in range(5):
            pm.decode_device(t)
        L.pp_debug_phase_cycles(buf, 1)
        n = 5 * B * 17
        tot = sum(buf[i] for i in range(12))
        print(f"== {label}: {tot / n:9.0f} cycles per heatmap per CTA / team")
        for i, nm in enumerate(names):
            print(f"   {nm:16s} {buf[i] / n:9.0f}  {100 * buf[i] / max(tot, 1):5.1f}%")


if __name__ == "__main__":
    main()
