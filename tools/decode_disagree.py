"""Find heatmaps on which the tensor-core decoder and the general kernels disagree and ask the oracle who is right.
Usage: python tools/decode_disagree.py [config] [seeds]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import oracle as oc
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth
from decode_mma_probe import make


def main():
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    wl = synth.WORKLOADS[cid]
    B, K = wl.batch, wl.num_keypoints
    W, H = wl.heatmap_size
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    sig = np.asarray(wl.sigmas)
    dev = torch.device("cuda")
    for seed in range(seeds):
        torch.manual_seed(seed)
        for dt in (torch.float32, torch.bfloat16):
            pred = make(B, wl, dev).to(dt)
            outs = {}
            for name, env in (("mma", {"PP_DECODE_MMA": "1"}), ("general", {"PP_DECODE_MMA": "0"})):
                os.environ.update(env)
                o = pm.decode_device(pred)
                outs[name] = {k: o[k].reshape(B * K, -1).cpu() for k in ("argmax", "vals", "locs")}
            bad = torch.nonzero((outs["mma"]["argmax"] != outs["general"]["argmax"]).any(1) |
                                (outs["mma"]["locs"] != outs["general"]["locs"]).any(1)).reshape(-1).tolist()
            print(f"seed {seed} {str(dt)[6:]}: {len(bad)} of {B * K} heatmaps differ", flush=True)
            flat = pred.reshape(B * K, H, W)
            for n in bad[:10]:
                k = n % K
                hm = flat[n].float().cpu().numpy()
                l, v, conv = oc.heatmap_expected_value(hm[None], sig[k:k + 1], return_heatmap=True, conv="scipy")
                am = int(conv.reshape(-1).argmax())
                top = np.sort(conv.reshape(-1))[-3:]
                row = [f"hm {n} (k={k}, radius {int(np.ceil(3 * np.clip((sig[k] * 2) ** 2 * np.sqrt(H / 1.25 * W / 1.25) * 2, 0.55, 3.0)))}) oracle argmax {am}",
                       f"top3 {top}", f"range [{hm.min():.3g}, {hm.max():.3g}]"]
                for name in ("mma", "general"):
                    a = int(outs[name]["argmax"][n])
                    ok = a == am and np.allclose(outs[name]["locs"][n].numpy(), l[0], rtol=1e-5, atol=1e-5)
                    row.append(f"{name}: argmax {a} locs {outs[name]['locs'][n].tolist()} {'OK' if ok else 'WRONG'}")
                # the same heatmap alone (B = 1, all channels zero but this one's) through every general kernel
                solo = torch.zeros((1, K, H, W), dtype=pred.dtype, device=dev)
                solo[0, k] = flat[n]
                for mode, env in (("team1", {"PP_DECODE_WARP": "1", "PP_DECODE_TEAM": "1"}), ("team2", {"PP_DECODE_WARP": "1", "PP_DECODE_TEAM": "2"}),
                                  ("team1-fulltaps", {"PP_DECODE_WARP": "1", "PP_DECODE_TEAM": "1", "PP_DECODE_FULLTAPS": "1"}),
                                  ("cta", {"PP_DECODE_WARP": "0"})):
                    os.environ.update({"PP_DECODE_MMA": "0", "PP_DECODE_FULLTAPS": "0", **env})
                    a = int(pm.decode_device(solo)["argmax"][0, k])
                    row.append(f"solo {mode}: {a} {'OK' if a == am else 'WRONG'}")
                    for key in ("PP_DECODE_WARP", "PP_DECODE_TEAM", "PP_DECODE_FULLTAPS"):
                        os.environ.pop(key, None)
                print("   " + " | ".join(row), flush=True)
                np.save(f"gpurun_out/disagree_c{cid}_{str(dt)[6:]}_{n}_k{k}.npy", hm)
            del pred


if __name__ == "__main__":
    main()
