"""Per-heatmap timeline of the tensor-core expected-OKS decoder (debug instance, pp_debug_decode_mma_timeline):
where a warp's time goes per phase, and when the warps of the launch finish.
Usage: python tools/decode_timeline.py [config=2] [batch=256]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import _lib, synth
from decode_mma_probe import make


def main():
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    wl = synth.WORKLOADS[cid]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else wl.batch
    dev = torch.device("cuda")
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pred = make(B, wl, dev)
    K = wl.num_keypoints
    W, H = wl.heatmap_size
    from probpose_pytorch_b200.heatmap import _oks_table
    tab = _oks_table(wl.sigmas, K, H, W, dev)
    p = _lib.DecodeParams(B, K, H, W, _lib.dtype_code(pred.dtype), 0, 1.0, float(wl.input_size[0]), float(wl.input_size[1]))
    locs = torch.empty((B, K, 2), device=dev)
    vals = torch.empty((B, K), device=dev)
    arg = torch.empty((B, K), dtype=torch.int32, device=dev)
    times = torch.zeros((B * K, 8), dtype=torch.int64, device=dev)
    scratch = torch.zeros(int(_lib.lib().pp_decode_expected_scratch_bytes_for(p)) // 4 + 1, dtype=torch.int32, device=dev)
    fn = _lib.lib().pp_debug_decode_mma_timeline
    fn.argtypes = [C.POINTER(_lib.DecodeParams), C.POINTER(_lib.OksTable)] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for rep in range(3):
        flush.zero_()
        times.zero_()
        torch.cuda.synchronize()
        rc = fn(p, tab.descriptor(), _lib.ptr(pred), _lib.ptr(locs), _lib.ptr(vals), _lib.ptr(arg), _lib.ptr(times),
                _lib.ptr(scratch), scratch.numel() * 4, _lib.stream_ptr(dev))
        _lib.check(rc, "pp_debug_decode_mma_timeline")
        torch.cuda.synchronize()
    t = times.cpu().numpy().astype(np.int64)
    done = t[:, 6] != 0
    t = t[done]
    ghz = 1.9
    ns0 = t[:, 0] - t[:, 0].min()
    it = t[:, 7] & 255
    cnt = (t[:, 7] >> 8) & 0xFFFFFF
    smid = t[:, 7] >> 32
    ph = np.diff(t[:, 1:7], axis=1) / ghz / 1e3          # us per phase
    names = ["plane wait", "scan", "fragments + sweep", "re-evaluation", "outputs"]
    print(f"C{cid} B={B}: {len(t)} heatmaps decoded by the tensor-core kernel, {int((~done).sum())} handed on; SMs seen: {len(set(smid.tolist()))}")
    for i in sorted(set(it.tolist()))[:4]:
        m = it == i
        print(f"  item #{i} of a warp ({int(m.sum())} heatmaps): start {ns0[m].mean() / 1e3:6.1f} us (min {ns0[m].min() / 1e3:.1f}, max {ns0[m].max() / 1e3:.1f}) | "
              + " | ".join(f"{n} {ph[m, j].mean():5.2f}" for j, n in enumerate(names)) + f" | total {ph[m].sum(1).mean():5.2f} (p95 {np.percentile(ph[m].sum(1), 95):5.2f}, max {ph[m].sum(1).max():5.2f}) us")
    end = ns0 / 1e3 + ph.sum(1)
    print(f"  heatmaps finished by: 50 % {np.percentile(end, 50):.1f} us, 90 % {np.percentile(end, 90):.1f}, 99 % {np.percentile(end, 99):.1f}, last {end.max():.1f} us after the first warp started")
    print(f"  candidates per heatmap: mean {cnt.mean():.1f}, p95 {np.percentile(cnt, 95):.0f}, max {cnt.max()}; re-evaluation us vs candidates: "
          + ", ".join(f"{c}:{ph[cnt == c, 3].mean():.2f}" for c in sorted(set(cnt.tolist()))[:8]))
    rad = np.asarray([int(r) for r in tab.radius.cpu().tolist()]) if hasattr(tab, "radius") else None
    if rad is not None:
        k = np.nonzero(done)[0] % K
        for r in sorted(set(rad.tolist())):
            m = rad[k] == r
            print(f"  radius {r}: {int(m.sum())} heatmaps, total {ph[m].sum(1).mean():5.2f} us, sweep {ph[m, 2].mean():5.2f}, re-evaluation {ph[m, 3].mean():5.2f}")


if __name__ == "__main__":
    main()
