"""Do the decoder and the loss kernel share SMs when launched on two streams?  Times each alone and both
together (CUDA graph, events) for several CTAs-per-SM caps (PP_DECODE_CTAS / PP_LOSS_CTAS are read at launch)."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probpose_pytorch_b200 as pp
from probpose_pytorch_b200 import synth, _lib
from probpose_pytorch_b200.loss import _Prepared

dev = torch.device("cuda:0")
wl = synth.WORKLOADS[2]
B, K = wl.batch, wl.num_keypoints
am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
loss_fn = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)
sets = []
for s in range(6):
    kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=s)
    enc = am.encode_batch(torch.from_numpy(kps).to(dev), torch.from_numpy(vis).to(dev))
    jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=50 + s)).to(dev)
    blob = am.encode_batch(jit, torch.from_numpy(vis).to(dev))["heatmaps"]
    amp = torch.from_numpy(synth.blob_params((B, K), seed=60 + s)).to(dev)
    pred = (blob * amp[:, :, None, None] + torch.rand_like(blob) * 0.02).clamp_(0, 1).contiguous()
    sets.append(dict(pred=pred, prep=_Prepared(loss_fn, pred, enc["heatmaps"], enc["keypoint_weights"], None, _lib.PP_LOSS_PIXEL_MEAN)))
side = torch.cuda.Stream(device=dev)

def both(s, do_dec, do_loss):
    cur = torch.cuda.current_stream(dev)
    keep = []
    if do_dec:
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            keep.append(pm.decode_device(s["pred"]))
    if do_loss:
        keep.append(s["prep"].forward(want_grad=True))
    if do_dec:
        cur.wait_stream(side)
    return keep

def timeit(do_dec, do_loss, iters=200):
    for s in sets:
        both(s, do_dec, do_loss)
    torch.cuda.synchronize()
    gs = []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            k = both(s, do_dec, do_loss)
        gs.append((g, k))
    for g, _ in gs:
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        gs[i % len(gs)][0].replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters

for d, l in [(0, 0), (3, 2), (4, 2), (3, 3), (2, 2), (4, 1), (5, 1)]:
    os.environ["PP_DECODE_CTAS"], os.environ["PP_LOSS_CTAS"] = str(d), str(l)
    print(f"D={d} L={l}: decode {timeit(True, False):.1f} us, loss {timeit(False, True):.1f} us, both {timeit(True, True):.1f} us", flush=True)
