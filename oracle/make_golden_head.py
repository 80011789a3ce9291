"""Generate tests/golden/head.npz: the REFERENCE's ``ProbMapHead`` (imported from /root/reference, CPU, random
init, seeded) run on a seeded feature map.  Stored: the activations that enter the tail of ``forward_heatmap``
(output of ``final_layer``, captured with a forward hook, head.py:523-525) and what the reference returns
(``forward_heatmap`` and the 5-tuple of ``forward``, head.py:487-534).  The GPU tests feed the stored activations
to this package's fused tail and must reproduce the reference's heatmaps bit for bit.
Build container only:  ``python -m oracle.make_golden_head``."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"


def main() -> None:
    sys.path.insert(0, REF)
    from probpose.head import ProbMapHead  # the reference

    torch.manual_seed(1234)
    head = ProbMapHead(768, 17, [(4, 3), (2, 2), (2, 2)], (256, 256), (4, 4)).eval()
    # random-init heads emit nearly constant maps; rescale the final layer so that the activations spread over the
    # clamp's three regimes (< 0, inside, > 1 after / temperature)
    x = torch.randn(2, 768, 16, 12)
    with torch.no_grad():
        feat = head.conv_layers(head.deconv_layers(x))
        raw = head.final_layer(feat) - head.final_layer.bias[None, :, None, None]
        head.final_layer.weight.mul_(0.4 / float(raw.std()))
        head.final_layer.bias.fill_(0.2)
    captured = {}
    head.final_layer.register_forward_hook(lambda m, i, o: captured.__setitem__("pre_tail", o.detach().clone()))
    with torch.no_grad():
        hm = head.forward_heatmap(x)
        out5 = head(x)
    pre = captured["pre_tail"].numpy()
    t = float(head.temperature)
    frac = [(pre / t < 0).mean(), ((pre / t >= 0) & (pre / t <= 1)).mean(), (pre / t > 1).mean()]
    np.savez_compressed(OUT / "head.npz", pre_tail=pre, heatmaps=hm.numpy(), temperature=np.float64(t),
                        forward_heatmaps=out5[0].numpy(), **{f"forward_{n}": o.numpy() for n, o in
                                                            zip(("probabilities", "visibilities", "oks", "error"), out5[1:])})
    print("pre_tail", pre.shape, "regimes (<0, in, >1):", [round(float(f), 3) for f in frac], "temperature", t)


if __name__ == "__main__":
    main()
