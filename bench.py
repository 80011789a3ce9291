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

The headline (`value`, `ms_per_step`, `roofline`, `e2e`) is the largest single-GPU configuration of BASELINE.json
(C5: COCO-WholeBody 133 keypoints, B = 512, 64x48; 836.8 MB per tensor -- nothing fits L2); `configs` carries the same
measurement for C2 (the ViT-B head decode shape, B = 256), C3 (the 8-GPU training shape: B = 1024 split over the ranks,
strong scaling) and C4 (96x72, B = 512; plus its decode-only sweep, split over the ranks), each with its own
`ms_per_step`, kernel times and roofline fractions.  `fused_step` is the same step with the target encode folded into
the loss kernel (`OKSHeatmapLoss.forward_mean_encoded`: the target is never written or read, 3 x HW x e per heatmap).
After the timed region a seeded sample of the timed buffers is checked against the CPU oracle (`parity_check`).
"""

from __future__ import annotations

import argparse
import collections
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
HEADLINE = 5               # BASELINE.json config the N = 1 line is quoted on (largest single-GPU configuration)


def workload_text(wl):
    return (wl.name + " -- heatmap path only (target encode + expected-OKS decode + OKS loss fwd/bwd); "
            "the ViT backbone / head convolutions are not on the path")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", type=int, default=HEADLINE, help="BASELINE.json configuration (2..5) of the headline")
    ap.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    ap.add_argument("--sets", type=int, default=0, help="rotating buffer sets (0 = enough for the working set to exceed L2)")
    ap.add_argument("--sample", type=int, default=0, help="reference arm: images per step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subconfigs", action="store_true", help="headline configuration only (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--exchange-every", type=int, default=0,
                    help="multi-GPU, --exchange nccl: steps per all-gather bucket (0 = number of buffer sets)")
    ap.add_argument("--serial", action="store_true", help="run decode after encode on one stream (no fork/join)")
    ap.add_argument("--consumer", choices=["branch", "side"], default="branch",
                    help="multi-GPU mailbox consumer inside the step graph: a branch of its own, or behind the record packing")
    ap.add_argument("--plain-backward", action="store_true",
                    help="loss.backward() with autograd's implicit ones_like gradient instead of backward(unit_upstream)")
    ap.add_argument("--publish", choices=["deferred", "finalize"], default="deferred",
                    help="multi-GPU mailbox: the loss of a step is published at the start of the next step (deferred) or by "
                         "the loss' own finalize kernel at the end of its step")
    ap.add_argument("--exchange", choices=["mailbox", "nccl"], default="mailbox",
                    help="multi-GPU record + loss exchange: stores into every rank's mailbox over NVLink peer memory from the "
                         "record-packing kernel (default), or bucketed NCCL all-gathers")
    return ap.parse_args()


# =============================================================================================
# reference arm: the oracle port (CPU restatement of the reference, checked bit-exact against it)
# on the host cores.  This is the only place besides the tests and parity_check that executes oracle/.
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
    sample = args.sample or min(wl.batch, max(cores, 32))
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

    def summary(self, marks=None):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        timed = [s for s in self.samples if s[4] is not None and (marks is None or s[4] in marks) and str(s[4]).startswith("timed")]
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


def hbm_peak():
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        return float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


DECODE_KERNELS = {0: "conv_exact_kernel + argmax_from_conv_kernel", 1: "decode_expected_fast_kernel (one CTA per heatmap)",
                  2: "decode_expected_warp_kernel (one warp per heatmap, pruned float32 prefilter)",
                  3: "decode_expected_dense_kernel", 4: "decode_expected_kernel (generic)",
                  5: "decode_expected_mma_kernel (one warp per heatmap, tensor-core Toeplitz prefilter + exact fp64 re-evaluation)"}


class Bench:
    """One workload (a BASELINE.json configuration, this rank's share of it) on this rank's GPU."""

    def __init__(self, ctx, cfg: int, batch: int, label: str, scaling: str):
        import numpy as np
        import torch

        import probpose_pytorch_b200 as pp
        from probpose_pytorch_b200 import synth

        self.ctx, self.cfg, self.label, self.scaling = ctx, cfg, label, scaling
        args, dev, rank = ctx["args"], ctx["dev"], ctx["rank"]
        self.torch, self.pp, self.np = torch, pp, np
        self.wl = wl = synth.WORKLOADS[cfg]
        self.B, self.K = batch, wl.num_keypoints
        self.W, self.H = wl.heatmap_size
        self.tdtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
        self.esize = 4 if args.dtype == "fp32" else 2
        self.hm_bytes = self.H * self.W * self.esize
        self.n_hm = self.B * self.K
        tensor_mb = self.n_hm * self.hm_bytes / 1e6
        # enough rotating sets that consecutive steps never find their inputs in the 126 MB L2
        self.nsets = args.sets or (2 if 3 * tensor_mb > 400 else 4)
        self.am = pp.ArgMaxProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
        self.pm = pp.ProbMap(wl.input_size, wl.heatmap_size, wl.sigmas)
        self.codec = pp.Codec(self.pm)
        self.loss_fn = pp.OKSHeatmapLoss(use_target_weight=True, smoothing_weight=0.05, oks_type="minus", check_target=False)
        self.sets = []
        for s in range(self.nsets):
            kps, vis, _ = synth.make_keypoints(wl, batch=self.B, seed=1000 + cfg + 97 * rank + s)
            kps_d, vis_d = torch.from_numpy(kps).to(dev), torch.from_numpy(vis).to(dev)
            jit = torch.from_numpy(synth.jitter_keypoints(wl, kps, seed=5000 + s)).to(dev)
            pred = self.am.encode_batch(jit, vis_d)["heatmaps"]
            amp = torch.from_numpy(synth.blob_params((self.B, self.K), seed=6000 + s)).to(dev)
            g = torch.Generator(device=dev).manual_seed(7000 + 13 * rank + s)
            pred.mul_(amp[:, :, None, None]).add_(torch.rand(pred.shape, device=dev, generator=g).mul_(0.02)).clamp_(0, 1)
            pred = pred.to(self.tdtype).contiguous()
            heads = [torch.rand((self.B, self.K, 1, 1), device=dev, generator=g) for _ in range(4)]
            self.sets.append(dict(kps=kps_d, vis=vis_d, pred=pred, heads=heads, kps_np=kps, vis_np=vis))
        self.set_bytes = 3 * self.n_hm * self.hm_bytes
        self.side = torch.cuda.Stream(device=dev)
        self.cons = torch.cuda.Stream(device=dev)
        self._sync_token = torch.zeros(1, device=dev)
        self.mailbox, self.exchange_note = None, ""
        torch.cuda.synchronize()

    # ---- the step on device-resident inputs -------------------------------------------------------------
    def step(self, s, slot=0, fused=False, mailbox=None):
        """encode -> loss on the current stream, decode + record packing beside them on a second stream (fork / join;
        inside a captured graph these are two branches).  `fused`: the encode happens inside the loss kernel.
        With a mailbox (multi-GPU) the step also carries the exchange: its records are stored into every rank's mailbox
        by the packing kernel; its loss lands in the mailbox's loss slot and is published at the START of the next step
        (commit_deferred: the NVLink round trips of a publication are not at the tail of the step that produced it);
        and a third branch consumes -- waits for, copies, acknowledges -- what ALL ranks published two steps ago.  Two
        steps of slack: that wait is normally over before it starts, and no rank runs more than two steps ahead."""
        torch, args = self.torch, self.ctx["args"]
        cur = torch.cuda.current_stream(self.ctx["dev"])
        pred5 = (s["pred"], *s["heads"])
        rec = None

        def exchange_side():
            if args.publish == "deferred":
                mailbox.commit_deferred((slot - 1) % mailbox.slots)
            mailbox.read_async((slot - 2) % mailbox.slots)

        if not args.serial:
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                dec = self.pm.decode_device(s["pred"])
                rec = self.codec.pack_records(dec, pred5, mailbox=mailbox, slot=slot)
                if mailbox is not None and args.consumer == "side":
                    exchange_side()
            if mailbox is not None and args.consumer == "branch":
                self.cons.wait_stream(cur)
                with torch.cuda.stream(self.cons):
                    exchange_side()
        out = s["pred"].detach().requires_grad_(True)
        tgt = None
        pub = loss_out = None
        if mailbox is not None:
            if args.publish == "deferred":
                loss_out = mailbox.loss_slot(slot)
            else:
                pub = mailbox.descriptor(slot)          # the loss' finalize kernel publishes it
        if fused:
            loss = self.loss_fn.forward_mean_encoded(out, self.am, s["kps"], s["vis"], publish=pub, loss_out=loss_out)
        else:
            enc = self.am.encode_batch(s["kps"], s["vis"], dtype=self.tdtype)
            tgt = enc["heatmaps"]
            loss = self.loss_fn.forward_mean(out, tgt, enc["keypoint_weights"], publish=pub, loss_out=loss_out)
        if args.serial:
            dec = self.pm.decode_device(s["pred"])
        if args.plain_backward:
            loss.backward()
        else:      # the same without the ones_like fill and the no-op rescale launch (loss.unit_upstream)
            loss.backward(gradient=self.pp.unit_upstream(loss.device, loss.dtype))
        if not args.serial:
            cur.wait_stream(self.side)
            if mailbox is not None and args.consumer == "branch":
                cur.wait_stream(self.cons)
        if rec is None:
            rec = self.codec.pack_records(dec, pred5, mailbox=mailbox, slot=slot)
            if mailbox is not None:
                exchange_side()
        if mailbox is not None and pub is not None:
            mailbox.loss_enqueued(slot)
        return dict(rec=rec, loss=loss.detach(), grad=out.grad, dec=dec, tgt=tgt)

    def _graphs(self, fn, n_graphs):
        """Warm up and capture one CUDA graph per step phase: phase g works on buffer set g % sets and mailbox slot g.
        The phases are visited cyclically, always: the in-graph consumer of phase g acknowledges phase g - 2."""
        torch = self.torch
        for _ in range(2):
            for g in range(n_graphs):
                fn(self.sets[g % len(self.sets)], g)
        torch.cuda.synchronize()
        graphs, results = [], []
        if not self.ctx["args"].no_graph:
            for g in range(n_graphs):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    results.append(fn(self.sets[g % len(self.sets)], g))
                graphs.append(gr)
            torch.cuda.synchronize()
        return graphs, results

    def time_step(self, steps, warmup, fused=False, exchange=True, mark="timed"):
        """ms for exactly `steps` steps (max over ranks), the last results and what the exchange did."""
        torch, ctx = self.torch, self.ctx
        import torch.distributed as dist
        from probpose_pytorch_b200 import distributed as ppd
        args, world, dev = ctx["args"], ctx["world"], ctx["dev"]
        mailbox = None
        note = ""
        n_phases = len(self.sets)
        if world > 1 and exchange and args.exchange == "mailbox":
            try:
                if self.mailbox is None:
                    self.mailbox = ppd.PeerMailbox(self.B, self.K, max(4, len(self.sets)), dev)
                mailbox = self.mailbox
                n_phases = mailbox.slots
                how = ("the loss at the start of the next step (pp_mailbox_commit_deferred)" if args.publish == "deferred"
                       else "the loss by the loss' finalize kernel")
                note = (f"records stored into every rank's mailbox over NVLink by pp_pack_records, {how} ({mailbox.slots} slots, "
                        "flow control); every rank consumes (waits for, copies, acknowledges) every step's slot two steps later, "
                        "inside the step's CUDA graph")
            except Exception as e:   # symmetric memory unavailable on this box: the NCCL path still measures the step
                note = f"NCCL (symmetric memory unavailable: {type(e).__name__}: {str(e)[:80]})"
        graphs, results = self._graphs(lambda s, g: self.step(s, g, fused=fused, mailbox=mailbox), n_phases)
        use_graph = bool(graphs)
        pending, bucket = collections.deque(), []
        every = max(1, min(args.exchange_every or len(self.sets), len(self.sets)))
        nccl = world > 1 and exchange and mailbox is None

        def run_step(i):
            g = i % n_phases
            if use_graph:
                graphs[g].replay()
                res = results[g]
                if mailbox is not None:
                    mailbox.published(g)      # the replayed graph published slot g again
            else:
                res = self.step(self.sets[g % len(self.sets)], g, fused=fused, mailbox=mailbox)
            if nccl:
                bucket.append((res["rec"], res["loss"]))
                if len(bucket) == every:
                    pending.append(ppd.exchange_bucket([b[0] for b in bucket], [b[1] for b in bucket], async_op=True))
                    bucket.clear()
                    if len(pending) > 2:
                        pending.popleft().wait()
            return res

        def drain():
            if bucket:
                pending.append(ppd.exchange_bucket([b[0] for b in bucket], [b[1] for b in bucket], async_op=True))
                bucket.clear()
            while pending:
                pending.popleft().wait()

        sampler = ctx["sampler"]
        sampler.mark = "warm"
        n_warm = -(-max(warmup, 3) // n_phases) * n_phases     # whole cycles of phases: the timed loop starts at phase 0
        for i in range(n_warm):
            run_step(i)
        drain()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if world > 1:
            # still outside the timed region: a tiny collective on this stream, so that ev0 is recorded when it
            # completes -- within a microsecond or two on every rank -- instead of whenever each rank's CPU gets round to it
            # after the barrier (tens of microseconds apart: in a 2 ms region that is several percent, and the exchange's
            # final wait would charge the latest starter's delay to everybody)
            dist.all_reduce(self._sync_token)
        sampler.mark = mark
        ev0.record()
        res = None
        for i in range(steps):
            res = run_step(i)
        if mailbox is not None and args.publish == "deferred":
            mailbox.commit_deferred((steps - 1) % n_phases)    # the last step's loss: published inside the timed region too
        drain()          # the main stream waits for the last exchanges: they are inside the timed region
        ev1.record()
        torch.cuda.synchronize()
        sampler.mark = "post"
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        checked = None
        if mailbox is not None:
            # finish the cycle of phases (the next user of this mailbox starts at phase 0 again), then consume the two
            # publications nobody has acknowledged yet; check_async raises if a consumer wait timed out or a producer
            # waited in vain for an acknowledgement
            for i in range(steps, -(-steps // n_phases) * n_phases):
                run_step(i)
            total = -(-steps // n_phases) * n_phases
            if args.publish == "deferred":
                mailbox.commit_deferred((total - 1) % n_phases)
            outs = {}
            for i in (total - 2, total - 1):
                outs[i % n_phases] = mailbox.read_async(i % n_phases)
            mailbox.check_async()
            # what arrived in this rank's mailbox for the last step must be what an all-gather of the ranks' results gives
            g_last = (total - 1) % n_phases
            got_rec, got_loss = outs[g_last]
            last_res = results[g_last] if use_graph else res
            want = ppd.exchange_step_results(last_res["rec"], last_res["loss"].to(torch.float64))
            checked = bool(torch.equal(got_rec.view(want.records.shape), want.records)
                           and torch.allclose(got_loss.mean(), want.loss.double(), rtol=1e-6, atol=0))
            if not checked:
                raise RuntimeError("mailbox exchange does not match the all-gather of the ranks' results")
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        last = (results[(steps - 1) % n_phases] if use_graph else res)
        last = dict(last, _set=((steps - 1) % n_phases) % len(self.sets))
        info = {"exchange": (note + f"; consumed steps inside the timed region: {max(0, steps - 2)}; last step checked against an "
                             f"all-gather: {checked}") if mailbox is not None else
                (f"records + loss of {every} step(s) per asynchronous NCCL all-gather" + (f" [{note}]" if note else "")) if nccl else "",
                "launch": ("CUDA graph replay" if use_graph else "eager")
                          + (", one stream" if args.serial else ", decode on a second stream beside encode->loss")}
        del graphs
        return ms, last, info

    def time_kernel(self, fn, iters):
        """Average device time of one launch: each buffer set's launch is captured in a CUDA graph so that
        the Python/ctypes call overhead (~20 us) does not hide a ~10 us kernel; events on the launching stream."""
        torch = self.torch
        for s in self.sets:
            fn(s)
        torch.cuda.synchronize()
        gs = []
        for s in self.sets:
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
        t = a.elapsed_time(b) * 1e-3 / iters
        del gs
        return t

    def kernel_times(self, iters, extra=False):
        torch, pp = self.torch, self.pp
        from probpose_pytorch_b200 import _lib
        from probpose_pytorch_b200.loss import _Prepared, _PreparedEncoded
        for s in self.sets:
            t = self.am.encode_batch(s["kps"], s["vis"], dtype=self.tdtype)
            s["tgt"], s["w"] = t["heatmaps"], t["keypoint_weights"]
        prep, prep_e = {}, {}

        def k_loss(s):
            if id(s) not in prep:
                prep[id(s)] = _Prepared(self.loss_fn, s["pred"], s["tgt"], s["w"], None, _lib.PP_LOSS_PIXEL_MEAN)
            return prep[id(s)].forward(want_grad=True)

        def k_loss_encoded(s):
            if id(s) not in prep_e:
                prep_e[id(s)] = _PreparedEncoded(self.loss_fn, s["pred"], self.am, s["kps"], s["vis"], None)
            return prep_e[id(s)].forward(want_grad=True)

        fns = {"encode": lambda s: self.am.encode_batch(s["kps"], s["vis"], dtype=self.tdtype),
               "decode_expected": lambda s: self.pm.decode_device(s["pred"]),
               "loss_fwd_bwd": k_loss, "loss_encoded_fwd_bwd": k_loss_encoded}
        alg = {"encode": 1, "decode_expected": 1, "loss_fwd_bwd": 3, "loss_encoded_fwd_bwd": 2}
        if extra:
            for s in self.sets:
                s["logits"] = ((s["pred"].float() - 0.3) * 4.0).to(self.tdtype)
            fns.update({"decode_dark": lambda s: self.am.decode_device(s["pred"]),
                        "head_tail": lambda s: pp.heatmap_tail(s["logits"], 0.5)})
            alg.update({"decode_dark": 1, "head_tail": 2})
        kt = {k: self.time_kernel(f, iters) for k, f in fns.items()}
        self.decode_kernel = DECODE_KERNELS.get(_lib.lib().pp_decode_expected_last_kernel(), "?")
        peak, _ = hbm_peak()
        out = {k: {"ms": 1e3 * t, "GBps": alg[k] * self.n_hm * self.hm_bytes / t / 1e9,
                   "frac": alg[k] * self.n_hm * self.hm_bytes / t / 1e9 / peak, "heatmaps_per_s": self.n_hm / t,
                   "algorithmic_planes": alg[k]} for k, t in kt.items()}
        for s in self.sets:
            s.pop("tgt", None); s.pop("logits", None)
        return out

    def record(self, steps, warmup, extra_kernels=False, mark="timed"):
        """ms_per_step, throughput, kernel times and roofline fractions of this workload."""
        world = self.ctx["world"]
        peak, peak_src = hbm_peak()
        ms, last, info = self.time_step(steps, warmup, fused=False, mark=mark)
        ms_f, last_f, _ = self.time_step(steps, warmup, fused=True, mark=mark + "_fused")
        ms_nox = self.time_step(steps, warmup, fused=False, exchange=False, mark="other")[0] if world > 1 else None
        kernels = self.kernel_times(max(20, min(200, steps)), extra=extra_kernels)
        n_job = world * self.n_hm
        per_step = ms / steps
        names = {"encode": "encode_kernel (OKS target encode, separable fp64 factors)",
                 "decode_expected": self.decode_kernel,
                 "loss_fwd_bwd": "oks_loss_fast_kernel (fused OKS loss forward+backward, TMA-staged planes)"}
        step_kernels = ("encode", "decode_expected", "loss_fwd_bwd")
        dom = max(step_kernels, key=lambda k: kernels[k]["ms"])
        serial = sum(kernels[k]["ms"] for k in step_kernels)
        traffic = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            traffic = json.loads(tfile.read_text()).get(f"{dom}/C{self.cfg}/{self.ctx['args'].dtype}")
        gb = lambda planes, ms_step: planes * self.n_hm * self.hm_bytes / (ms_step * 1e-3) / 1e9
        rec = {
            "workload": self.wl.name, "scaling": self.scaling, "batch_per_gpu": self.B, "keypoints": self.K,
            "heatmap": [self.W, self.H], "heatmaps_per_step_per_gpu": self.n_hm, "tensor_MB": self.n_hm * self.hm_bytes / 1e6,
            "l2": f"rotating {self.nsets} buffer sets x {self.set_bytes / 1e6:.0f} MB (> 126 MB L2)",
            "steps": steps, "ms_per_step": per_step, "value": n_job * steps / (ms * 1e-3), "unit": UNIT,
            "launch": info["launch"], "exchange": info["exchange"],
            "ms_per_step_without_exchange": (ms_nox / steps) if ms_nox is not None else None,
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_planes"] * self.n_hm * self.hm_bytes,
                         "share_of_step": {k: kernels[k]["ms"] / serial for k in step_kernels},
                         "other_step_kernels": {k: {"kernel": names[k], "achieved": kernels[k]["GBps"], "frac": kernels[k]["frac"]}
                                                for k in step_kernels if k != dom},
                         "step": {"GBps": gb(5, per_step), "frac": gb(5, per_step) / peak,
                                  "note": "whole step, 5 x H x W x e bytes per heatmap (write target, read prediction twice, "
                                          "read target, write gradient)"}},
            "fused_step": {"what": "same step with the target encode inside the loss kernel (forward_mean_encoded): the target is "
                                   "never written or read -- 3 x H x W x e bytes per heatmap (prediction read by decode and by "
                                   "the loss, gradient written)",
                           "ms_per_step": ms_f / steps, "value": n_job * steps / (ms_f * 1e-3),
                           "GBps_own_bytes": gb(3, ms_f / steps), "frac_own_bytes": gb(3, ms_f / steps) / peak,
                           "frac_if_counted_as_5_planes": gb(5, ms_f / steps) / peak},
            "kernels": kernels,
        }
        return rec, last, last_f


def parity_check(bench, last, last_f, n_sample=64):
    """A seeded sample of the TIMED buffers against the CPU oracle: decode outputs of the last timed step (argmax / scores
    bit-exact, coordinates 1e-5), the encoded targets and the loss gradient on the sampled heatmaps, and the loss value
    of a sub-batch.  The oracle is the checker here, never the thing measured."""
    import numpy as np
    import torch

    import oracle as oc
    wl, B, K, H, W = bench.wl, bench.B, bench.K, bench.H, bench.W
    j = (last["_set"] if "_set" in last else 0)
    s = bench.sets[j]
    rng = np.random.default_rng(2024)
    pick = np.sort(rng.choice(B * K, size=min(n_sample, B * K), replace=False))
    pick_t = torch.from_numpy(pick).to(s["pred"].device)
    pred = s["pred"].reshape(B * K, H, W)[pick_t].float().cpu().numpy()
    sig = np.asarray(wl.sigmas)
    arg = last["dec"]["argmax"].reshape(-1)[pick_t].cpu().numpy()
    vals = last["dec"]["vals"].reshape(-1)[pick_t].cpu().numpy()
    locs = last["dec"]["locs"].reshape(-1, 2)[pick_t].cpu().numpy()
    tgt = last["tgt"].reshape(B * K, H, W)[pick_t].float().cpu().numpy()
    grad = last["grad"].reshape(B * K, H, W)[pick_t].float().cpu().numpy()
    grad_f = last_f["grad"].reshape(B * K, H, W)[pick_t].float().cpu().numpy()
    out = {"heatmaps_checked": int(pick.size), "argmax_mismatch": 0, "score_mismatch": 0, "locs_max_rel": 0.0,
           "target_max_rel": 0.0, "grad_max_rel": 0.0, "fused_grad_max_rel": 0.0}
    scale_n = 1.0 / (B * K * H * W)
    for i, n in enumerate(pick):
        b, k = divmod(int(n), K)
        l, v, conv = oc.heatmap_expected_value(pred[i][None], sig[k:k + 1], return_heatmap=True, conv="scipy")
        out["argmax_mismatch"] += int(arg[i] != conv.reshape(-1).argmax())
        out["score_mismatch"] += int(vals[i] != v[0])
        out["locs_max_rel"] = max(out["locs_max_rel"], float(np.max(np.abs(locs[i] - l[0]) / np.maximum(np.abs(l[0]), 1.0))))
        enc = oc.encode("argmax", wl.input_size, wl.heatmap_size, wl.sigmas, s["kps_np"][b:b + 1], s["vis_np"][b:b + 1])
        t_ref = enc["heatmaps"][k]
        if bench.esize == 2:
            t_ref = torch.from_numpy(t_ref).bfloat16().float().numpy()
        out["target_max_rel"] = max(out["target_max_rel"], float(np.max(np.abs(tgt[i] - t_ref) / np.maximum(np.abs(t_ref), 1e-30)
                                                                        * (np.abs(t_ref) > 1e-30))))
        w = float(enc["keypoint_weights"][0, k])
        g_ref = oc.oks_heatmap_loss_grad_closed_form(torch.from_numpy(pred[i][None, None]), torch.from_numpy(t_ref[None, None]),
                                                     torch.full((1, 1, 1, 1), w), smoothing_weight=0.05, oks_type="minus")
        g_ref = np.asarray(g_ref)[0, 0] * scale_n * (H * W)     # the closed form is the gradient of the mean over ONE heatmap
        den = np.abs(g_ref).max() * 0.05 + np.abs(g_ref) + 1e-30
        out["grad_max_rel"] = max(out["grad_max_rel"], float(np.max(np.abs(grad[i] - g_ref) / den)))
        out["fused_grad_max_rel"] = max(out["fused_grad_max_rel"], float(np.max(np.abs(grad_f[i] - g_ref) / den)))
    # loss value of a sub-batch (the mean over the whole batch would take the CPU minutes)
    nb = max(1, min(B, 1024 // K))
    o = s["pred"][:nb].float().cpu()
    encs = [oc.encode("argmax", wl.input_size, wl.heatmap_size, wl.sigmas, s["kps_np"][b:b + 1], s["vis_np"][b:b + 1]) for b in range(nb)]
    t = torch.from_numpy(np.stack([e["heatmaps"] for e in encs]))
    wts = torch.from_numpy(np.concatenate([e["keypoint_weights"] for e in encs]).astype(np.float32))
    if bench.esize == 2:
        t = t.bfloat16().float()
    l_ref = float(oc.oks_heatmap_loss(o, t, wts, per_pixel=True, smoothing_weight=0.05, oks_type="minus").mean())
    enc_d = bench.am.encode_batch(s["kps"][:nb], s["vis"][:nb], dtype=bench.tdtype)
    l_got = float(bench.loss_fn.forward_mean(s["pred"][:nb], enc_d["heatmaps"], enc_d["keypoint_weights"]))
    l_fused = float(bench.loss_fn.forward_mean_encoded(s["pred"][:nb], bench.am, s["kps"][:nb], s["vis"][:nb]))
    out["loss_value_rel"] = abs(l_got - l_ref) / abs(l_ref)
    out["fused_loss_value_rel"] = abs(l_fused - l_ref) / abs(l_ref)
    tol = 1e-5 if bench.esize == 4 else 1e-2
    out["tolerance"] = tol
    out["ok"] = bool(out["argmax_mismatch"] == 0 and out["score_mismatch"] == 0 and out["locs_max_rel"] <= 1e-5
                     and out["target_max_rel"] <= tol and out["grad_max_rel"] <= tol and out["fused_grad_max_rel"] <= tol
                     and out["loss_value_rel"] <= tol and out["fused_loss_value_rel"] <= tol)
    return out


def run_e2e(bench, e2e_steps):
    """The same metric end to end through the public API with HOST buffers (pinned): every step's inputs (keypoints,
    visibility, predicted heatmaps, the four scalar heads) are copied host -> device and its results (decoded records +
    loss) device -> host inside the timed region; `pipelined_steps` double-buffers the copies."""
    import torch
    import torch.distributed as dist
    from probpose_pytorch_b200.host_io import pipelined_steps
    ctx = bench.ctx
    dev, world = ctx["dev"], ctx["world"]
    hosts = []
    for s in bench.sets:
        hosts.append((torch.from_numpy(s["kps_np"]).pin_memory(), torch.from_numpy(s["vis_np"]).pin_memory(),
                      s["pred"].cpu().pin_memory(), *[h.cpu().pin_memory() for h in s["heads"]]))

    def step_fn(kps, vis, pred, *heads):
        enc = bench.am.encode_batch(kps, vis, dtype=bench.tdtype)
        rec = bench.codec.decode_device((pred, *heads))
        out = pred.detach().requires_grad_(True)
        loss = bench.loss_fn.forward_mean(out, enc["heatmaps"], enc["keypoint_weights"])
        loss.backward()
        return rec, loss.detach()

    def run(n):
        last = None
        for last in pipelined_steps((hosts[i % len(hosts)] for i in range(n)), step_fn, dev):
            pass
        return last

    run(3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    rec_h, loss_h = run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    h2d = sum(t.numel() * t.element_size() for t in hosts[0])
    d2h = rec_h.numel() * rec_h.element_size() + loss_h.numel() * loss_h.element_size()
    # the PCIe copy that bounds this number, timed alone (same pinned buffer, CUDA events)
    dst = torch.empty_like(bench.sets[0]["pred"])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(hosts[0][2], non_blocking=True)
    a.record()
    for _ in range(3):
        dst.copy_(hosts[0][2], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    h2d_ms = a.elapsed_time(b) / 3
    pred_bytes = hosts[0][2].numel() * hosts[0][2].element_size()
    if world > 1:
        t = torch.tensor([h2d_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_ms = float(t)
    return {"value": world * bench.n_hm * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
            "h2d_alone": {"ms": h2d_ms, "GBps": pred_bytes / h2d_ms / 1e6, "ranks_copying_at_once": world,
                          "ceiling_heatmaps_per_s": world * bench.n_hm / (h2d_ms * 1e-3),
                          "note": "host->device copy of one step's predicted heatmaps timed alone (slowest rank when all ranks copy "
                                  "at once): the PCIe / host-memory ceiling of e2e"},
            "note": "public API (host_io.pipelined_steps around encode_batch, Codec.decode_device, "
                    "OKSHeatmapLoss.forward_mean + backward) on pinned host buffers; copies of neighbouring steps "
                    "overlap the kernels"}


def run_product(args):
    import torch
    import torch.distributed as dist

    from probpose_pytorch_b200 import _lib, synth

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
    sampler = ClockSampler(local)
    sampler.start()
    ctx = dict(args=args, dev=dev, world=world, rank=rank, sampler=sampler)

    # ---- headline: the configuration's full batch on every GPU (weak scaling)
    wl = synth.WORKLOADS[args.config]
    head = Bench(ctx, args.config, wl.batch if args.config != 3 else wl.batch // 8, f"C{args.config}", "weak")
    rec, last, last_f = head.record(args.steps, args.warmup, extra_kernels=True, mark="timed")
    parity = None
    if rank == 0:
        parity = parity_check(head, last, last_f)
    sampler.mark = "e2e"
    e2e = run_e2e(head, max(5, min(20, args.steps)))
    clocks = sampler.summary(marks={"timed", "timed_fused"})
    del last, last_f
    head.sets.clear()
    del head
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, same measurement, shorter runs
    configs = {}
    if not args.no_subconfigs:
        sub_steps = max(10, min(args.steps, 100))
        plan = [("C2", 2, synth.WORKLOADS[2].batch, "weak (B = 256 per GPU)"),
                ("C3", 3, max(1, synth.WORKLOADS[3].batch // world), f"strong (B = 1024 split over {world} GPU(s), loss + records exchanged)"),
                ("C4", 4, max(1, synth.WORKLOADS[4].batch // world), f"strong (B = 512 split over {world} GPU(s))"),
                ("C3_per_gpu_128", 3, synth.WORKLOADS[3].batch // 8, "weak (B = 128 per GPU, the per-GPU share of C3 as BASELINE "
                 "quotes it: 128/GPU x 8; loss + records exchanged)")]
        for label, cid, batch, scaling in plan:
            if cid == args.config:
                continue
            b = Bench(ctx, cid, batch, label, scaling)
            # these records are not bound to --steps: time at least ~30 ms so that launch / rank skew (tens of microseconds)
            # stays well below a percent of a region of 0.1 ms steps
            est = b.time_kernel(lambda s: b.step(s)["loss"], 5)
            if world > 1:
                tt = torch.tensor([est], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                est = float(tt)
            steps_here = int(min(1000, max(sub_steps, 0.03 / max(est, 1e-6))))
            r, _, _ = b.record(steps_here, 5, mark="timed_" + label)
            r["clocks"] = sampler.summary(marks={"timed_" + label, "timed_" + label + "_fused"})
            if cid == 4:
                # BASELINE configs[3]: decode throughput sweep over 1/2/4/8 GPUs -- decode only, B = 512 split over the ranks
                t = b.time_kernel(lambda s: b.pm.decode_device(s["pred"]), sub_steps)
                if world > 1:
                    tt = torch.tensor([t], device=dev)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    t = float(tt)
                r["decode_sweep"] = {"what": "expected-OKS decode only, B = 512 split over the ranks (strong scaling), device time of the "
                                             "slowest rank", "n_gpus": world, "ms": 1e3 * t, "value": world * b.n_hm / t, "unit": "heatmaps/s"}
            configs[label] = r
            b.sets.clear()
            del b
            torch.cuda.empty_cache()

    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                   "--config", str(args.config)]
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            cpu_baseline = json.loads(res.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}

    if rank == 0:
        # kernels of ours inside the timed region, per step: encode, decode + its hand-over launch, record packing, loss,
        # loss finalize (+ the gradient rescale launch with --plain-backward) (+ the mailbox's deferred commit and consumer)
        launches_per_step = 6 + (1 if args.plain_backward else 0) + (2 if world > 1 and args.exchange == "mailbox" else 0)
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": workload_text(wl),
                       "arithmetic": "f32 maps; f64 for the encode exponentials and the exact argmax check; f16 tensor-core proposal "
                                     "step in the decoder" if args.dtype == "fp32" else "bf16 maps, f32 arithmetic",
                       "batch_per_gpu": rec["batch_per_gpu"], "keypoints": rec["keypoints"], "heatmap": rec["heatmap"],
                       "heatmaps_per_step_per_gpu": rec["heatmaps_per_step_per_gpu"], "l2": rec["l2"], "launch": rec["launch"],
                       "parallelism": f"dp{world} (batch sharded by image)" + (f"; {rec['exchange']}" if rec["exchange"] else "")},
            "ms_per_step_without_exchange": rec["ms_per_step_without_exchange"],
            "roofline": rec["roofline"], "kernels": rec["kernels"], "fused_step": rec["fused_step"], "e2e": e2e,
            "parity_check": parity, "configs": configs, "cpu_baseline": cpu_baseline,
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
