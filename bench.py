#!/usr/bin/env python
"""Benchmark of the ProbPose heatmap hot path: heatmaps/s of encode + decode + OKS loss fwd/bwd.

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA, this repo)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the CPU port on host cores
    torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU, NCCL

One *step* = one pass of the hot path over one batch of synthetic input (per GPU):
    target encode  (keypoints -> (B,K,H,W) maps)             write 1 x HW x e
    decode         (predicted maps -> keypoint records)      read  1 x HW x e
    loss fwd+bwd   (read prediction + target, write grad)    read 2, write 1
=> 5 x H x W x e algorithmic bytes per heatmap.  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "heatmaps/sec encode+decode+loss (BxKxHxW)"
UNIT = "heatmaps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


def workload_text(wl):
    return (wl.name + " -- heatmap path only (target encode + expected-OKS decode + OKS loss fwd/bwd); "
            "the ViT backbone / head convolutions are not on the path")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configuration (2..5) whose shapes are used")
    ap.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    ap.add_argument("--sets", type=int, default=4, help="rotating buffer sets (working set must exceed L2)")
    ap.add_argument("--sample", type=int, default=0, help="reference arm: images per step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--exchange-every", type=int, default=0,
                    help="multi-GPU: steps per all-gather bucket (0 = number of buffer sets; 1 = one collective per step)")
    ap.add_argument("--serial", action="store_true", help="run decode after encode on one stream (no fork/join)")
    ap.add_argument("--exchange", choices=["mailbox", "nccl"], default="mailbox",
                    help="multi-GPU record + loss exchange: stores into every rank's mailbox over NVLink peer memory from the "
                         "record-packing kernel (default), or bucketed NCCL all-gathers")
    return ap.parse_args()


# =============================================================================================
# reference arm: the oracle port (CPU restatement of the reference, checked bit-exact against it)
# on the host cores.  This is the only place besides the tests that executes oracle/.
# =============================================================================================
def _cpu_chunk(job):
    """encode -> expected-OKS decode -> loss fwd+bwd for a chunk of images, as the reference does it:
    per-sample NumPy encode (codec.py:138), per-sample scipy convolution decode (codec.py:214) and the
    torch-CPU loss with autograd (loss.py:428-431)."""
    import numpy as np
    import torch

    import oracle as oc
    from probpose_pytorch_b200 import synth

    cfg, lo, hi, seed = job
    torch.set_num_threads(1)
    wl = synth.WORKLOADS[cfg]
    kps, vis, _ = synth.make_keypoints(wl, batch=hi, seed=seed)
    kps, vis = kps[lo:hi], vis[lo:hi]
    n = hi - lo
    t0 = time.perf_counter()
    tgt = np.stack([oc.encode("argmax", wl.input_size, wl.heatmap_size, wl.sigmas, kps[b:b + 1], vis[b:b + 1])["heatmaps"]
                    for b in range(n)])
    t1 = time.perf_counter()
    pred = np.clip(tgt * 0.8 + 0.01, 0, 1).astype(np.float32)   # stand-in prediction (not timed as encode)
    t2 = time.perf_counter()
    for b in range(n):
        oc.decode_expected(pred[b], wl.input_size, wl.heatmap_size, wl.sigmas, conv="scipy")
    t3 = time.perf_counter()
    o = torch.from_numpy(pred).requires_grad_(True)
    w = torch.from_numpy(vis)
    loss = oc.oks_heatmap_loss(o, torch.from_numpy(tgt), w, per_pixel=True, smoothing_weight=0.05, oks_type="minus").mean()
    loss.backward()
    t4 = time.perf_counter()
    return n, (t1 - t0), (t3 - t2), (t4 - t3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    from probpose_pytorch_b200 import synth

    wl = synth.WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    sample = args.sample or min(wl.batch, max(cores, 64))
    ctx = mp.get_context("fork")

    def one_step(pool, images, seed):
        per = max(1, (images + cores - 1) // cores)
        jobs = [(args.config, lo, min(images, lo + per), seed) for lo in range(0, images, per)]
        t0 = time.perf_counter()
        parts = pool.map(_cpu_chunk, jobs)
        return time.perf_counter() - t0, parts

    with ctx.Pool(cores) as pool:
        t_probe, _ = one_step(pool, sample, 0)          # also warms the workers (imports)
        t_probe, _ = one_step(pool, sample, 0)
        budget = 150.0
        total_steps = args.steps + args.warmup
        if t_probe * total_steps > budget:               # keep the whole run within a few minutes
            sample = max(cores // 2 or 1, int(sample * budget / (t_probe * total_steps)))
        for i in range(args.warmup):
            one_step(pool, sample, i)
        t0 = time.perf_counter()
        enc = dec = los = 0.0
        for i in range(args.steps):
            _, parts = one_step(pool, sample, 100 + i)
            enc += sum(p[1] for p in parts); dec += sum(p[2] for p in parts); los += sum(p[3] for p in parts)
        dt = time.perf_counter() - t0
    hms = sample * wl.num_keypoints * args.steps
    value = hms / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(wl), "heatmap": list(wl.heatmap_size), "keypoints": wl.num_keypoints,
                   "images_per_step": sample, "arithmetic": "f32 maps; f64 encode / convolution accumulation (NumPy, SciPy)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images x {wl.num_keypoints} keypoints per step, {args.steps} steps; "
                                   "oracle port (NumPy encode, scipy.ndimage decode, torch-CPU loss+autograd), "
                                   f"{cores} worker processes",
                         "cpu_seconds_split": {"encode": enc, "decode": dec, "loss": los}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =============================================================================================
# product arm
# =============================================================================================
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the GPU is being timed."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.mark = index, [], False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.perf_counter(), mhz, reasons, util, self.mark))
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        timed = [s for s in self.samples if s[4] == "timed"]
        pool = timed or [s for s in self.samples if s[4] is not None] or self.samples
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost",
                 0x100: "display_clock_setting"}
        seen = set()
        for s in pool:
            for bit, name in names.items():
                if s[2] & bit:
                    seen.add(name)
        return {"sm_mhz": statistics.median(s[1] for s in pool), "sm_max_mhz": self.max_mhz, "reasons": sorted(seen),
                "samples": len(pool), "window": "timed region" if timed else "whole measurement phase"}


def run_product(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import probpose_pytorch_b200 as pp
    from probpose_pytorch_b200 import _lib, synth
    from probpose_pytorch_b200 import distributed as ppd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (product arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    wl = synth.WORKLOADS[args.config]
    B, K = wl.batch, wl.num_keypoints            # per-GPU batch: weak scaling (C3 = 128 images per GPU)
    if args.config == 3:
        B = wl.batch // 8
    W, H = wl.heatmap_size
    tdtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    esize = 4 if args.dtype == "fp32" else 2
    hm_bytes = H * W * esize
    n_hm = B * K

    am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
    codec = pp.Codec(pm)
    loss_fn = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)

    # ---- synthetic inputs: `sets` distinct buffer sets so that consecutive steps never find their
    # inputs in L2 (per set: prediction + target + gradient = 3 x B*K*H*W*e bytes)
    sets = []
    rng = np.random.default_rng(1000 + args.config + 17 * rank)
    for s in range(args.sets):
        kps, vis, _ = synth.make_keypoints(wl, batch=B, seed=1000 + args.config + 97 * rank + s)
        kps_d = torch.from_numpy(kps).to(dev)
        vis_d = torch.from_numpy(vis).to(dev)
        jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=5000 + s)).to(dev)
        blob = am.encode_batch(jit, vis_d)["heatmaps"]
        amp = torch.from_numpy(synth.blob_params((B, K), seed=6000 + s)).to(dev)
        noise = torch.rand(blob.shape, device=dev) * 0.02
        pred = (blob * amp[:, :, None, None] + noise).clamp_(0, 1).to(tdtype).contiguous()
        heads = [torch.rand((B, K, 1, 1), device=dev) for _ in range(4)]
        sets.append(dict(kps=kps_d, vis=vis_d, pred=pred, heads=heads, kps_host=torch.from_numpy(kps).pin_memory(),
                         vis_host=torch.from_numpy(vis).pin_memory(), pred_host=pred.cpu().pin_memory()))
        del blob, noise
    set_bytes = 3 * n_hm * hm_bytes
    torch.cuda.synchronize()

    side = torch.cuda.Stream(device=dev)

    # ---- multi-GPU exchange of the step's results (keypoint records + loss): a mailbox per rank in NVLink peer
    # memory, written by the record-packing kernel itself; one slot per rotating buffer set
    mailbox, exchange_note = None, ""
    if world > 1 and args.exchange == "mailbox":
        try:
            mailbox = ppd.PeerMailbox(B, K, len(sets), dev)
            exchange_note = f"records + loss stored into every rank's mailbox over NVLink by pp_pack_records ({len(sets)} slots)"
        except Exception as e:   # symmetric memory unavailable on this box: the NCCL path still measures the step
            exchange_note = f"NCCL (symmetric memory unavailable: {type(e).__name__}: {str(e)[:80]})"
            mailbox = None

    def step_device(s, slot=0):
        """The hot path on device-resident inputs.  The decode only depends on the prediction, so it runs on
        a second stream next to encode -> loss (fork / join; inside the captured graph these are two branches);
        the records are packed (and, on several GPUs, published together with the loss) after the join."""
        cur = torch.cuda.current_stream(dev)
        pred5 = (s["pred"], *s["heads"])
        rec = None
        if not args.serial:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                dec = pm.decode_device(s["pred"])
                # the records are packed right behind the decode and, on several GPUs, stored into every rank's
                # mailbox by the same kernel; the loss joins them after the join (commit)
                rec = codec.pack_records(dec, pred5, mailbox=mailbox, slot=slot)
        enc = am.encode_batch(s["kps"], s["vis"], dtype=tdtype)
        if args.serial:
            dec = pm.decode_device(s["pred"])
        out = s["pred"].detach().requires_grad_(True)
        loss = loss_fn.forward_mean(out, enc["heatmaps"], enc["keypoint_weights"])
        loss.backward()
        if not args.serial:
            cur.wait_stream(side)
        if rec is None:
            rec = codec.pack_records(dec, pred5, mailbox=mailbox, slot=slot)
        if mailbox is not None:
            mailbox.commit(slot, loss.detach())
        return rec, loss.detach(), out.grad

    # ---- CUDA graphs of the step, one per buffer set
    graphs, results = [], []
    use_graph = not args.no_graph
    for j, s in enumerate(sets):
        for _ in range(2):
            step_device(s, j)
    torch.cuda.synchronize()
    if use_graph:
        for j, s in enumerate(sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                results.append(step_device(s, j))
            graphs.append(g)
        torch.cuda.synchronize()

    def run_step(i):
        j = i % len(sets)
        if use_graph:
            graphs[j].replay()
            rec, loss, _ = results[j]
            if mailbox is not None:
                mailbox.published(j)      # the replayed graph published slot j again
        else:
            rec, loss, _ = step_device(sets[j], j)
        if world > 1 and mailbox is None:
            # the only exchange on the path: keypoint records + loss partials.  It is latency bound, so the results
            # of `--exchange-every` consecutive steps (<= the number of rotating buffer sets, whose outputs are
            # still intact) travel in ONE all-gather, issued asynchronously so that it overlaps the next steps'
            # kernels (bounded number in flight).
            bucket.append((rec, loss))
            if len(bucket) == exchange_every:
                pending.append(ppd.exchange_bucket([b[0] for b in bucket], [b[1] for b in bucket], async_op=True))
                bucket.clear()
                if len(pending) > 2:
                    pending.popleft().wait()
        return rec, loss

    import collections
    pending = collections.deque()

    bucket = []
    exchange_every = max(1, min(args.exchange_every or len(sets), len(sets)))

    def drain():
        if bucket:      # a partial bucket at the end of a run still travels
            pending.append(ppd.exchange_bucket([b[0] for b in bucket], [b[1] for b in bucket], async_op=True))
            bucket.clear()
        while pending:
            pending.popleft().wait()

    sampler = ClockSampler(local)
    sampler.start()
    sampler.mark = "warm"
    for i in range(max(args.warmup, 3)):
        run_step(i)
    drain()
    torch.cuda.synchronize()

    # ---- timed region: EXACTLY args.steps steps, barrier + synchronize on both sides
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark = "timed"
    ev0.record()
    for i in range(args.steps):
        run_step(i)
    drain()          # the main stream waits for the last exchanges: they are inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    sampler.mark = "post"
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    exchange_checked = None
    if mailbox is not None:
        # what arrived in this rank's mailbox for the last step must be what an all-gather of the ranks' results gives
        j = (args.steps - 1) % len(sets)
        got_rec, got_loss = mailbox.read(j)
        rec_l, loss_l = (results[j][0], results[j][1]) if use_graph else step_device(sets[j], j)[:2]
        want = ppd.exchange_step_results(rec_l, loss_l.to(torch.float64))
        exchange_checked = bool(torch.equal(got_rec.view(want.records.shape), want.records)
                                and torch.allclose(got_loss.mean(), want.loss.double(), rtol=1e-6, atol=0))
        if not exchange_checked:
            raise RuntimeError("mailbox exchange does not match the all-gather of the ranks' results")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * n_hm * args.steps / (ms * 1e-3)

    # ---- per-kernel timings (CUDA events on the launching stream, same rotating sets)
    def time_kernel(fn, iters):
        """Average device time of one launch: each buffer set's launch is captured in a CUDA graph so that
        the Python/ctypes call overhead (~20 us) does not hide a ~10 us kernel; events on the launching stream."""
        for s in sets:
            fn(s)
        torch.cuda.synchronize()
        gs = []
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = fn(s)
            gs.append((g, keep))
        for g, _ in gs:
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            gs[i % len(gs)][0].replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / iters

    targets = [am.encode_batch(s["kps"], s["vis"], dtype=tdtype) for s in sets]
    for s, t in zip(sets, targets):
        s["tgt"], s["w"] = t["heatmaps"], t["keypoint_weights"]
    grads = [torch.empty_like(s["pred"]) for s in sets]

    def k_encode(s):
        return am.encode_batch(s["kps"], s["vis"], dtype=tdtype)

    def k_decode(s):
        return pm.decode_device(s["pred"])

    def k_dark(s):
        return am.decode_device(s["pred"])

    prep_cache = {}

    def k_loss(s):
        from probpose_pytorch_b200.loss import _Prepared
        key = id(s)
        if key not in prep_cache:
            prep_cache[key] = _Prepared(loss_fn, s["pred"], s["tgt"], s["w"], None, _lib.PP_LOSS_PIXEL_MEAN)
        return prep_cache[key].forward(want_grad=True)

    # head tails (section 8 a10 / f-2), not part of the step: logits with the spread of a trained head
    for s in sets:
        s["logits"] = ((s["pred"].float() - 0.3) * 4.0).to(tdtype).requires_grad_(True)
        s["up"] = torch.rand_like(s["pred"])

    def k_tail(s):
        return pp.heatmap_tail(s["logits"].detach(), 0.5)

    def k_sparse_fwd(s):
        return pp.heatmap_tail(s["logits"].detach(), 0.5, normalize=1.0)

    def k_sparse_fwd_bwd(s):
        s["logits"].grad = None
        y = pp.heatmap_tail(s["logits"], 0.5, normalize=1.0)
        y.backward(s["up"])
        return y, s["logits"].grad

    sampler.mark = "kernels"
    iters = max(20, min(400, args.steps))
    kt = {"encode": time_kernel(k_encode, iters), "decode_expected": time_kernel(k_decode, iters),
          "decode_dark": time_kernel(k_dark, iters), "loss_fwd_bwd": time_kernel(k_loss, iters),
          "head_tail": time_kernel(k_tail, iters), "sparsemax_tail_fwd": time_kernel(k_sparse_fwd, iters),
          "sparsemax_tail_fwd_bwd": time_kernel(k_sparse_fwd_bwd, iters)}
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    alg = {"encode": 1, "decode_expected": 1, "decode_dark": 1, "loss_fwd_bwd": 3, "head_tail": 2,
           "sparsemax_tail_fwd": 2, "sparsemax_tail_fwd_bwd": 5}
    kernels = {k: {"ms": 1e3 * t, "GBps": alg[k] * n_hm * hm_bytes / t / 1e9,
                   "frac": alg[k] * n_hm * hm_bytes / t / 1e9 / peak,
                   "heatmaps_per_s": n_hm / t} for k, t in kt.items()}
    # the dominant kernel = the one of the three step kernels with the largest share of the step's device time
    step_kernels = ("encode", "decode_expected", "loss_fwd_bwd")
    dom = max(step_kernels, key=lambda k: kt[k])
    decode_kernel = {0: "conv_exact_kernel + argmax_from_conv_kernel", 1: "decode_expected_fast_kernel (one CTA per heatmap)",
                     2: "decode_expected_warp_kernel (one warp per heatmap)", 3: "decode_expected_dense_kernel",
                     4: "decode_expected_kernel (generic)"}.get(_lib.lib().pp_decode_expected_last_kernel(), "?")
    names = {"encode": "encode_kernel (OKS target encode, separable fp64 factors)",
             "decode_expected": decode_kernel + ": expected-OKS decode, TMA-staged plane, pruned separable prefilter + exact fp64 "
                                                "re-evaluation; bounded by per-heatmap instruction latency, not by HBM",
             "loss_fwd_bwd": "oks_loss_fast_kernel (fused OKS loss forward+backward, TMA-staged)"}
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        traffic = json.loads(tfile.read_text()).get(f"{dom}/C{args.config}/{args.dtype}")
    serial = sum(kt[k] for k in step_kernels)
    roofline = {"bound": "hbm", "kernel": names[dom],
                "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom] * n_hm * hm_bytes,
                "share_of_step": {k: kt[k] / serial for k in step_kernels},
                "other_step_kernels": {k: {"kernel": names[k], "achieved": kernels[k]["GBps"], "frac": kernels[k]["frac"]}
                                       for k in step_kernels if k != dom},
                "step": {"GBps": 5 * n_hm * hm_bytes * args.steps / (ms * 1e-3) / 1e9,
                         "frac": 5 * n_hm * hm_bytes * args.steps / (ms * 1e-3) / 1e9 / peak,
                         "note": "whole step, 5 x H x W x e bytes per heatmap"}}

    # ---- end to end through the public API with HOST buffers (pinned): every step's inputs (keypoints,
    # visibility, predicted heatmaps, the four scalar heads) are copied host -> device and its results (decoded
    # records + loss) device -> host inside the timed region.  `pipelined_steps` double-buffers the copies so
    # that the PCIe transfer of step i+1 overlaps the kernels of step i.
    from probpose_pytorch_b200.host_io import pipelined_steps
    for s in sets:
        s["heads_host"] = [h.cpu().pin_memory() for h in s["heads"]]

    def host_batch(s):
        return (s["kps_host"], s["vis_host"], s["pred_host"], *s["heads_host"])

    def step_fn(kps, vis, pred, *heads):
        enc = am.encode_batch(kps, vis, dtype=tdtype)
        rec = codec.decode_device((pred, *heads))
        out = pred.detach().requires_grad_(True)
        loss = loss_fn.forward_mean(out, enc["heatmaps"], enc["keypoint_weights"])
        loss.backward()
        return rec, loss.detach()

    def run_e2e(n):
        last = None
        for last in pipelined_steps((host_batch(sets[i % len(sets)]) for i in range(n)), step_fn, dev):
            pass
        return last

    sampler.mark = "e2e"
    e2e_steps = max(5, min(50, args.steps))
    run_e2e(4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rec_h, loss_h = run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    s0 = sets[0]
    h2d = sum(t.numel() * t.element_size() for t in host_batch(s0))
    d2h = rec_h.numel() * rec_h.element_size() + loss_h.numel() * loss_h.element_size()
    # the PCIe copy that bounds this number, timed alone (same pinned buffer, CUDA events)
    dst = torch.empty_like(s0["pred"])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(s0["pred_host"], non_blocking=True)
    a.record()
    for _ in range(5):
        dst.copy_(s0["pred_host"], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    h2d_ms = a.elapsed_time(b) / 5
    pred_bytes = s0["pred_host"].numel() * s0["pred_host"].element_size()
    e2e = {"value": world * n_hm * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "h2d_alone": {"ms": h2d_ms, "GBps": pred_bytes / h2d_ms / 1e6,
                         "note": "host->device copy of one step's predicted heatmaps, timed alone: the floor of e2e"},
           "note": "public API (host_io.pipelined_steps around encode_batch, Codec.decode_device, "
                   "OKSHeatmapLoss.forward_mean + backward) on pinned host buffers; copies of neighbouring steps "
                   "overlap the kernels"}

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                   "--config", str(args.config)]
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            ref = json.loads(res.stdout.strip().splitlines()[-1])
            cpu_baseline = ref["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}

    if rank == 0:
        launches_per_step = 6 + (1 if mailbox is not None else 0)   # encode, decode, loss, loss finalize, grad-scale check, record packing (+ mailbox commit)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "fp32" else "bf16",
            "data": "synthetic",
            "config": {"workload": workload_text(wl),
                       "arithmetic": "f32 maps; f64 for the encode exponentials and the exact argmax check"
                                     if args.dtype == "fp32" else "bf16 maps, f32 arithmetic",
                       "batch_per_gpu": B, "keypoints": K, "heatmap": [W, H], "heatmaps_per_step_per_gpu": n_hm,
                       "l2": f"rotating {args.sets} buffer sets x {set_bytes / 1e6:.0f} MB (> 126 MB L2)",
                       "launch": ("CUDA graph replay" if use_graph else "eager")
                                 + (", one stream" if args.serial else ", decode on a second stream beside encode->loss"),
                       "parallelism": f"dp{world} (batch sharded by image)"
                                      + ("" if world == 1 else
                                         f"; {exchange_note}, checked against an all-gather: {exchange_checked}" if mailbox is not None else
                                         f"; records + loss of {exchange_every} step(s) per asynchronous all-gather"
                                         + (f" [{exchange_note}]" if exchange_note else ""))},
            "roofline": roofline, "kernels": kernels, "e2e": e2e, "cpu_baseline": cpu_baseline,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
