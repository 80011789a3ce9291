"""Expected-OKS decode timings, tensor-core kernel (pp_decode_mma.cuh) against the general kernels, on the bench's
inputs: mixed / noise-only / clean maps at C2, C3-per-GPU, B=1024, C4, C5, fp32 and bf16.  CUDA events around graph
replays.  Usage: python tools/decode_mma_probe.py [quick]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import _lib, synth
from decode_split import timed


def make(B, wl, dev, kind="mixed"):
    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1002)
    jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=5000)).to(dev)
    blob = am.encode_batch(jit, torch.from_numpy(vis).to(dev))["heatmaps"]
    if kind == "noise":
        return (torch.rand_like(blob) * 0.02).contiguous()
    amp = torch.from_numpy(synth.blob_params((B, wl.num_keypoints), seed=6000)).to(dev)
    if kind == "clean":
        return (blob * amp[:, :, None, None]).contiguous()
    return (blob * amp[:, :, None, None]).add_(torch.rand_like(blob) * 0.02).clamp_(0, 1).contiguous()


def main():
    quick = len(sys.argv) > 1
    dev = torch.device("cuda")
    cases = [(2, 256, "mixed"), (2, 256, "noise"), (2, 256, "clean"), (2, 128, "mixed"), (2, 1024, "mixed"),
             (4, 512, "mixed"), (5, 512, "mixed")]
    if quick:
        cases = cases[:3]
    for cid, B, kind in cases:
        wl = synth.WORKLOADS[cid]
        pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
        pred = make(B, wl, dev, kind)
        n = B * wl.num_keypoints
        for dt in (torch.float32, torch.bfloat16):
            x = pred.to(dt)
            row = []
            ref = None
            for mma in ("1", "minb3", "0"):
                os.environ["PP_DECODE_MMA"] = "0" if mma == "0" else "1"
                os.environ["PP_DECODE_MINB"] = "3" if mma == "minb3" else "4"
                us, out = timed(lambda: pm.decode_device(x), iters=20)
                kern = _lib.lib().pp_decode_expected_last_kernel()
                torch.cuda.synchronize()
                handed = int(out["_scratch"][2]) if kern == 5 else 0
                row.append(f"kernel {kern} ({handed} handed on): {us:8.1f} us {us * 1e3 / n:6.2f} ns/hm {n * x[0, 0].numel() * x.element_size() / us / 1e3:7.1f} GB/s")
                if ref is None:
                    ref = out
                else:
                    diff = {k: int((ref[k] != out[k]).sum()) for k in ("argmax", "vals", "locs") if not torch.equal(ref[k], out[k])}
                    row.append("same" if not diff else f"DIFFERENT {diff}")
            print(f"C{cid} B={B} {kind} {str(dt)[6:]}: " + " | ".join(row), flush=True)
        del pred


if __name__ == "__main__":
    main()
