#!/usr/bin/env python
"""bench.py -- HistoGAN training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload train|hist|rehisto]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Two workloads:

train (default; BASELINE.json configs[2]/[3], what "training images/sec at 256^2" is
  quoted on): one `Trainer.train()` step of HistoGAN -- D phase + G phase with the
  histogram loss, gradient penalty every 4th and path-length regulariser every 32nd
  step, DiffGrad updates -- at 256x256, network_capacity 16, batch 32 per GPU, synthetic
  images + random target histograms.  Under torchrun the batch shards over the ranks
  (weak scaling) and gradients are averaged with an NCCL all-reduce before each
  optimiser step.
hist (BASELINE.json configs[1]): RGBuvHistBlock forward + Hellinger loss + backward on
  32 x 3 x 256 x 256 per GPU, h = 64, insz = 256 (all 65 536 pixels enter the histogram;
  the Trainer's insz = 150 alongside); no collective.  The train line carries this
  block's us/img and achieved GB/s in `hist_block`.

One JSON line on stdout (rank 0).  `value` = images/s with the batch resident in HBM;
`e2e` = the same through the public API with the batch in pinned HOST memory (H2D of
images + target histograms and D2H of the loss scalars inside the timed region).
`--impl reference` times the reference's CPU algorithm (torch-CPU restatement in
oracle/, kind "port": the reference itself is Python and does not travel to the GPU
box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

B_PER_GPU = 32
S = 256
CAPACITY = 16
H_BINS = 64
ALPHA = 2.0
L2_FLUSH_BYTES = 256 << 20
FIRST_STEP = 2501       # past the reference's "evaluate every 100 steps < 2500" window
PL_STEP = 2528          # a multiple of 32: gradient penalty AND path-length regulariser


def window_start(steps):
    """first trainer.steps value of the timed window: centred on PL_STEP so that any K >= 2 holds
    one path-length step and every 4th step a gradient penalty (K = 32 is exactly the steady-state
    mix; shorter windows over-weight the PL step, i.e. are conservative)"""
    return max(FIRST_STEP, PL_STEP - steps // 2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0),
                d.get("bf16_tflops_sustained", 1400.0), "measured")
    return 6650.0, 1590.0, 1400.0, "fallback"


def make_inputs(rank, B=B_PER_GPU, device="cpu"):
    """generator-like images relu(randn*0.5+0.3) and random target histograms."""
    g = torch.Generator().manual_seed(rank)
    x = torch.relu(torch.randn(B, 3, S, S, generator=g) * 0.5 + 0.3)
    t = torch.rand(B, 3, H_BINS, H_BINS, generator=g)
    t = t / t.sum(dim=(1, 2, 3), keepdim=True)
    return x.to(device), t.to(device)


# ------------------------------------------------------------ clock sampler --

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if sm:
            # median over the samples taken under load (power above the idle/active midpoint)
            mid = (min(power) + max(power)) / 2
            load = sorted(c for c, p in zip(sm, power) if p >= mid) or sorted(sm)
            out.update(sm_mhz=load[len(load) // 2], sm_max_mhz=max(smax), samples=len(sm),
                       samples_under_load=len(load), power_w_max=max(power))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------ helpers --

class Dist:
    def __init__(self, gpus):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world == gpus, f"--gpus {gpus} but WORLD_SIZE={self.world}"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, seconds):
        if self.world > 1:
            t = torch.tensor([seconds], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return t.item()
        return seconds

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def time_call(fn, flush, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    ds = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ds.append(s.elapsed_time(e) * 1e-3)
    return sum(ds) / len(ds)


# -------------------------------------------------------- histogram workload --

def hist_section(dv, steps, warmup):
    """fwd + Hellinger + bwd of the histogram block; returns (dict, t_dev, t_e2e, launches)."""
    from histogan_b200 import RGBuvHistBlock, hellinger_loss, _lib
    lib = _lib.load()
    dev = dv.dev
    x_host, t_host = make_inputs(dv.rank)
    x_pin, t_pin = x_host.pin_memory(), t_host.pin_memory()
    x_dev, t_dev_ = x_pin.to(dev), t_pin.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    blk = RGBuvHistBlock(h=H_BINS, insz=256, device=dev)
    blk150 = RGBuvHistBlock(h=H_BINS, insz=150, device=dev)

    def step(xd, td, block=blk):
        xg = xd.detach().requires_grad_(True)
        hist = block(F.relu(xg))                     # histoGAN/histoGAN.py:955
        loss = hellinger_loss(td, hist, ALPHA)       # :957-960
        loss.backward()
        return loss, hist, xg.grad

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        dv.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(n)]
        for s, e in evs:
            flush.zero_()                            # 26.7 MB of input < 126 MB L2: flush it
            s.record(); fn(); e.record()
        dv.barrier()
        return dv.max_over_ranks(sum(s.elapsed_time(e) for s, e in evs) * 1e-3)

    n0 = lib.hg_launch_count()
    t_dev = timed(lambda: step(x_dev, t_dev_), steps, warmup)
    launches = (lib.hg_launch_count() - n0) // (steps + warmup) * steps

    loss_h = torch.empty((), dtype=torch.float32).pin_memory()
    hist_h = torch.empty(B_PER_GPU, 3, H_BINS, H_BINS).pin_memory()
    grad_h = torch.empty(B_PER_GPU, 3, S, S).pin_memory()

    def e2e_step():
        xd = x_pin.to(dev, non_blocking=True)
        td = t_pin.to(dev, non_blocking=True)
        loss, hist, gx = step(xd, td)
        loss_h.copy_(loss, non_blocking=True)
        hist_h.copy_(hist.detach(), non_blocking=True)
        grad_h.copy_(gx, non_blocking=True)

    t_e2e = timed(e2e_step, steps, max(1, warmup // 2))
    xg = x_dev.detach().requires_grad_(True)
    hist_keep = blk(F.relu(xg))
    g_up = torch.rand_like(hist_keep)
    t_fwd = time_call(lambda: blk(x_dev), flush)
    t_bwd = time_call(lambda: torch.autograd.grad(hist_keep, xg, g_up, retain_graph=True), flush)
    n150 = max(3, steps // 2)
    t150 = timed(lambda: step(x_dev, t_dev_, blk150), n150, 2)

    hbm_peak, _, _, peak_kind = load_peaks()
    N = S * S
    bytes_fwd, bytes_all = 835_584, 2_555_904         # per image, SURVEY 8(d)
    flops_fwd = 24_576 * N
    n_imgs = B_PER_GPU * dv.world * steps
    info = {
        "images_per_s": round(n_imgs / t_dev, 2),
        "us_per_image": round(t_dev / n_imgs * dv.world * 1e6, 3),
        "ms_per_step": round(t_dev / steps * 1e3, 4),
        "e2e_images_per_s": round(n_imgs / t_e2e, 2),
        "e2e_h2d_bytes_per_step": x_pin.numel() * 4 + t_pin.numel() * 4,
        "e2e_d2h_bytes_per_step": 4 + hist_h.numel() * 4 + grad_h.numel() * 4,
        "insz150_images_per_s": round(B_PER_GPU * dv.world * n150 / t150, 2),
        "fwd_ms": round(t_fwd * 1e3, 4), "bwd_ms": round(t_bwd * 1e3, 4),
        "hbm_gbs_fwd": round(bytes_fwd * B_PER_GPU / t_fwd / 1e9, 2),
        "hbm_gbs_bwd": round((bytes_all - bytes_fwd) * B_PER_GPU / t_bwd / 1e9, 2),
        "hbm_frac_bwd": round((bytes_all - bytes_fwd) * B_PER_GPU / t_bwd / 1e9 / hbm_peak, 5),
        "fwd_tflops": round(flops_fwd * B_PER_GPU / t_fwd / 1e12, 2),
        "bwd_tflops": round(2 * flops_fwd * B_PER_GPU / t_bwd / 1e12, 2),
        "peak_kind": peak_kind,
    }
    return info, t_dev, t_e2e, launches


def run_hist(args):
    from histogan_b200 import _lib
    dv = Dist(args.gpus)
    _lib.check(_lib.load().hg_device_check(dv.local_rank), "hg_device_check")
    sampler = ClockSampler(dv.local_rank) if dv.rank == 0 else None
    info, t_dev, t_e2e, launches = hist_section(dv, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else {}
    if dv.rank == 0:
        hbm_peak, _, _, peak_kind = load_peaks()
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        fma_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        out = {
            "metric": "images/sec", "value": info["images_per_s"], "unit": "images/s",
            "n_gpus": dv.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RGBuvHistBlock fwd + Hellinger loss + bwd, 32x3x256x256 per GPU, "
                                   "h=64, insz=256 (N=65536 px/img), inverse-quadratic sigma=0.02",
                       "global_batch": B_PER_GPU * dv.world,
                       "parallelism": f"dp{dv.world} (no collective)",
                       "l2_flush": "256 MiB memset between steps, outside the event windows"},
            "e2e": {"value": info["e2e_images_per_s"], "unit": "images/s",
                    "h2d_bytes_per_step": info["e2e_h2d_bytes_per_step"],
                    "d2h_bytes_per_step": info["e2e_d2h_bytes_per_step"]},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": "hist_bwd_fast_kernel (hg_hist_bwd: prep + pixel + adjoint)",
                "achieved": info["hbm_gbs_bwd"], "peak": hbm_peak, "unit": "GB/s",
                "frac": info["hbm_frac_bwd"], "traffic": None, "peak_kind": peak_kind,
                "note": "compute-bound op (1.9 kFLOP/B, SURVEY 8d): the HBM fraction is capped "
                        "near 1-2 %; fp32-FMA pipe fractions in `compute`",
                "compute": {"fma_peak_tflops_at_measured_clock": round(fma_peak, 1),
                            "fwd_frac_of_fma_peak": round(info["fwd_tflops"] / fma_peak, 4),
                            "bwd_frac_of_fma_peak": round(info["bwd_tflops"] / fma_peak, 4)}},
            "hist_block": info,
        }
        if dv.world == 1:
            out["cpu_baseline"] = cpu_baseline_hist()
        emit(out)
    dv.close()


# ------------------------------------------------------------ train workload --

class HostLoader:
    """batches in pinned host memory: Trainer.train copies them to the device each step
    (batch['images'].cuda(), histoGAN/histoGAN.py:895-898) -- the e2e data path."""

    def __init__(self, rank, eval_only=False):
        x, t = make_inputs(rank)
        self.x = torch.rand(x.shape, generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
        self.t = t.pin_memory()
        self.eval_only = eval_only

    def __iter__(self):
        return self

    def __next__(self):
        if self.eval_only:
            return {"histograms": self.t[:4]}
        return {"images": self.x, "histograms": self.t}


class DeviceLoader(HostLoader):
    def __init__(self, rank, dev, eval_only=False):
        super().__init__(rank, eval_only)
        self.x, self.t = self.x.to(dev), self.t.to(dev)


def _conv_flops(B, cin, cout, k, oh):
    return 2.0 * B * oh * oh * cout * cin * k * k


def conv_layer_table():
    """(net, Cin, Cout, k, stride, H_in) of every conv in G (incl. toRGB) and D at 256^2, cap 16."""
    g_f = [4 * CAPACITY] + [CAPACITY * 2 ** (i + 1) for i in range(7)][::-1]
    layers = []
    for i, (ci, co) in enumerate(zip(g_f[:-1], g_f[1:])):
        h = 4 * 2 ** i
        layers += [("G", ci, co, 3, 1, h), ("G", co, co, 3, 1, h), ("G", co, 3, 1, 1, h)]
    d_f = [3] + [CAPACITY * 2 ** i for i in range(8)]
    for i, (ci, co) in enumerate(zip(d_f[:-1], d_f[1:])):
        h = 256 // 2 ** i
        layers += [("D", ci, co, 1, 1, h), ("D", ci, co, 3, 1, h), ("D", co, co, 3, 1, h)]
        if i != len(d_f) - 2:
            layers.append(("D", co, co, 3, 2, h))
    return layers


def train_step_flops(B):
    """algorithmic conv FLOPs of one Trainer.train step (no GP / PL): G fwd x2 (the D-phase
    pass has no backward) + G bwd (2x fwd), D fwd x3 + D bwd x2 (2x fwd each)."""
    g = sum(_conv_flops(B, ci, co, k, h // s) for n, ci, co, k, s, h in conv_layer_table() if n == "G")
    d = sum(_conv_flops(B, ci, co, k, h // s) for n, ci, co, k, s, h in conv_layer_table() if n == "D")
    return g * (2 + 2) + d * (3 + 2 * 2), g, d


def graph_time(fns, reps=3):
    """device time per call of `fns` (a list of closures, one launch set each) replayed as ONE CUDA
    graph: no host issue time between the launches -- what the Trainer's captured phases see."""
    for f in fns[:2]:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3 / len(fns))
    del g
    return sorted(ts)[len(ts) // 2]


def conv_microbench(dev, B, rotate_bytes=320 << 20):
    """Device time of the three tcgen05 conv primitives (forward, input gradient, weight gradient) on
    every G/D layer shape of the step (channel counts padded as the networks carry them; the two image-input
    layers of D on the direct kernels that run them in the step).
    Each layer is captured as a CUDA graph of `n` launches on `n` DIFFERENT operand sets whose total
    size exceeds the 126 MB L2 (operands cold, as in the step), and the graph replay is timed."""
    from histogan_b200 import conv, ops
    tot = {"fwd": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    rows = []
    for net, ci, co, k, s, h in conv_layer_table():
        if co == 3:                 # toRGB runs the dedicated (CUDA-core) kernels
            continue
        cip, cop = ops._round_up(ci), ops._round_up(co)
        oh = h // s
        if ops.SMALL_CIN and conv.small_ok(ci, co, k, s, k // 2):
            # the image-input layers run the CUDA-core kernels of conv_small.cu on the planar image (as
            # DiscriminatorBlock 0 does): time those, not a channel-padded tensor-core launch
            n = 3
            imgs = [torch.randn(B, ci, h, h, device=dev) for _ in range(n)]
            dys = [conv.tf32_round(torch.randn(B, cop, h, h, device=dev)).contiguous(memory_format=torch.channels_last)
                   for _ in range(n)]
            w = torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5
            bias = torch.randn(co, device=dev)
            f = _conv_flops(B, ci, co, k, oh)
            t = {"fwd": graph_time([lambda i=i: conv.conv_small_fwd(imgs[i], w, cop, bias=bias, lrelu=True, round_tf32=True)
                                    for i in range(n)]),
                 "dgrad": graph_time([lambda i=i: conv.conv_small_dgrad(dys[i], w, ci) for i in range(n)]),
                 "wgrad": graph_time([lambda i=i: conv.conv_small_wgrad(dys[i], imgs[i], (co, ci, k, k)) for i in range(n)])}
            for kind, v in t.items():
                tot[kind][0] += f
                tot[kind][1] += v
            byt = 4 * (B * ci * h * h + B * cop * oh * oh)
            rows.append([f"{net} {ci}->{co} k{k} s{s} @{h} (CUDA cores)", round(f / t["fwd"] / 1e12, 1),
                         round(f / t["dgrad"] / 1e12, 1), round(f / t["wgrad"] / 1e12, 1), round(t["fwd"] * 1e6, 1),
                         round(t["dgrad"] * 1e6, 1), round(t["wgrad"] * 1e6, 1), round(byt / t["fwd"] / 1e9)])
            del imgs, dys
            continue
        per_set = 4 * (B * cip * h * h + B * cop * oh * oh + cop * cip * k * k)
        n = max(2, min(48, -(-rotate_bytes // per_set)))
        xs = [conv.tf32_round(torch.randn(B, cip, h, h, device=dev)).contiguous(memory_format=torch.channels_last)
              for _ in range(n)]
        dys = [conv.tf32_round(torch.randn(B, cop, oh, oh, device=dev)).contiguous(memory_format=torch.channels_last)
               for _ in range(n)]
        ws = [(torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
              for _ in range(n)]
        wp = [conv.pack_weight(w, 0) for w in ws]
        for w in ws:                # the dgrad operands (cached on the weight tensors)
            ops._raw_grad_input(dys[0], w, s, k // 2, (h, h), dy_rounded=True, padded_io=True)
        f = _conv_flops(B, ci, co, k, oh)
        t = {"fwd": graph_time([lambda i=i: conv.conv2d_nhwc(xs[i], wp[i], s, k // 2, cout=cop) for i in range(n)]),
             "dgrad": graph_time([lambda i=i: ops._raw_grad_input(dys[i], ws[i], s, k // 2, (h, h), dy_rounded=True,
                                                                  padded_io=True) for i in range(n)]),
             "wgrad": graph_time([lambda i=i: conv.conv2d_wgrad_nhwc(dys[i], xs[i], k, s, k // 2) for i in range(n)])}
        for kind, v in t.items():
            tot[kind][0] += f
            tot[kind][1] += v
        rows.append([f"{net} {ci}->{co} k{k} s{s} @{h}", round(f / t["fwd"] / 1e12, 1), round(f / t["dgrad"] / 1e12, 1),
                     round(f / t["wgrad"] / 1e12, 1), round(t["fwd"] * 1e6, 1), round(t["dgrad"] * 1e6, 1),
                     round(t["wgrad"] * 1e6, 1), round(per_set / t["fwd"] / 1e9)])
        del xs, dys, ws, wp
        torch.cuda.empty_cache()
    agg = {k: v[0] / v[1] / 1e12 for k, v in tot.items()}
    agg["all"] = sum(v[0] for v in tot.values()) / sum(v[1] for v in tot.values()) / 1e12
    agg["pass_us"] = {k: round(v[1] * 1e6, 1) for k, v in tot.items()}
    return agg, rows


def run_train(args):
    from histogan_b200 import _lib
    from histogan_b200.trainer import Trainer
    dv = Dist(args.gpus)
    lib = _lib.load()
    _lib.check(lib.hg_device_check(dv.local_rank), "hg_device_check")
    out_dir = os.path.join(ROOT, "gpurun_out", f"bench_rank{dv.rank}")
    torch.manual_seed(1234 + dv.rank)
    tr = Trainer("bench", os.path.join(out_dir, "results"), os.path.join(out_dir, "models"),
                 image_size=S, network_capacity=CAPACITY, batch_size=B_PER_GPU,
                 gradient_accumulate_every=1, hist_insz=150, hist_resizing="interpolation",
                 save_every=10 ** 9, fast_rng=True,
                 cuda_graphs=os.environ.get("HG_CUDA_GRAPHS", "1") != "0")
    tr.loader_evaluate = DeviceLoader(dv.rank, dv.dev, eval_only=True)
    sampler = ClockSampler(dv.local_rank) if dv.rank == 0 else None

    def timed(loader, steps, warmup):
        tr.loader = loader
        first = window_start(steps)
        tr.steps = first - warmup
        for _ in range(warmup):
            tr.train(alpha=ALPHA)
        dv.barrier()
        n0 = lib.hg_launch_count() + tr.graph_replayed_launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            tr.train(alpha=ALPHA)
        e.record()
        dv.barrier()
        assert tr.steps == first + steps
        return (dv.max_over_ranks(s.elapsed_time(e) * 1e-3),
                lib.hg_launch_count() + tr.graph_replayed_launches - n0)

    def step_kind_ms(first, n, stride):
        """median device time of `n` single steps starting at trainer.steps = first, first+stride.."""
        ts = []
        for i in range(n):
            tr.steps = first + i * stride
            dv.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); tr.train(alpha=ALPHA); e.record()
            dv.barrier()
            ts.append(dv.max_over_ranks(s.elapsed_time(e) * 1e-3))
        return round(sorted(ts)[len(ts) // 2] * 1e3, 3)

    # allocator / one-time kernel-attribute warm-up and capture of every graph variant (plain,
    # gradient-penalty, path-length) beyond the requested W (not timed)
    tr.loader = DeviceLoader(dv.rank, dv.dev)
    for st in (PL_STEP, PL_STEP + 1, PL_STEP + 2, PL_STEP + 4):
        tr.steps = st
        tr.train(alpha=ALPHA)
    t_dev, launches = timed(DeviceLoader(dv.rank, dv.dev), args.steps, args.warmup)
    step_ms = {"plain": step_kind_ms(PL_STEP + 1, 5, 4), "gp": step_kind_ms(PL_STEP + 4, 3, 8),
               "gp+pl": step_kind_ms(PL_STEP, 3, 32)}
    step_ms["steady_state_32"] = round((24 * step_ms["plain"] + 7 * step_ms["gp"] + step_ms["gp+pl"]) / 32, 3)
    host = HostLoader(dv.rank)
    t_e2e, _ = timed(host, args.steps, 1)
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    losses = {"d": tr.d_loss, "g": tr.g_loss, "h": tr.h_loss, "gp": tr.last_gp_loss}
    tr_graphs = tr.cuda_graphs
    del tr
    torch.cuda.empty_cache()
    if os.environ.get("HG_BENCH_LIGHT"):        # development runs: the step timing only
        if dv.rank == 0:
            emit({"metric": "training images/sec", "value": round(B_PER_GPU * dv.world * args.steps / t_dev, 2),
                  "unit": "images/s", "n_gpus": dv.world, "steps": args.steps, "ms_per_step": round(t_dev / args.steps * 1e3, 3),
                  "e2e": {"value": round(B_PER_GPU * dv.world * args.steps / t_e2e, 2)},
                  "config": {"step_ms": step_ms, "light": True}, "clocks": sampler.stop() if sampler else {}})
        dv.close()
        return
    hist_info, _, _, _ = hist_section(dv, 5, 3)
    conv_agg, conv_rows = ({}, [])
    if dv.rank == 0:
        conv_agg, conv_rows = conv_microbench(dv.dev, B_PER_GPU)
    conv_tf = conv_agg.get("fwd")
    clocks = sampler.stop() if sampler else {}

    if dv.rank == 0:
        hbm_peak, bf16_peak, bf16_sust, peak_kind = load_peaks()
        n_imgs = B_PER_GPU * dv.world * args.steps
        step_flops, g_f, d_f = train_step_flops(B_PER_GPU)
        step_s = t_dev / args.steps
        out = {
            "metric": "training images/sec", "value": round(n_imgs / t_dev, 2), "unit": "images/s",
            "n_gpus": dv.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_s * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 operands (round-to-nearest) / fp32 accumulate + fp32",
            "data": "synthetic",
            "config": {"workload": "HistoGAN Trainer.train step (D phase + G phase + histogram loss, "
                                   "GP every 4th / PL every 32nd step, DiffGrad), 256x256, "
                                   "network_capacity=16, batch 32 per GPU, hist_insz=150 interpolation",
                       "global_batch": B_PER_GPU * dv.world,
                       "parallelism": f"dp{dv.world}" + (" + NCCL grad all-reduce" if dv.world > 1 else ""),
                       "l2_flush": "not needed: the step's working set (GBs of activations) >> 126 MB L2",
                       "timed_steps": f"trainer.steps {window_start(args.steps)}.."
                                      f"{window_start(args.steps) + args.steps - 1}: "
                                      f"{sum(1 for k in range(window_start(args.steps), window_start(args.steps) + args.steps) if k % 4 == 0)} "
                                      f"gradient-penalty and "
                                      f"{sum(1 for k in range(window_start(args.steps), window_start(args.steps) + args.steps) if k % 32 == 0)} "
                                      f"path-length step(s) inside the window",
                       "step_ms": step_ms,
                       "rng": "device-side latent / noise generation (fast_rng=True)",
                       "cuda_graphs": bool(tr_graphs),
                       "peak_mem_gib": round(mem_gb, 2), "final_losses": losses},
            "e2e": {"value": round(n_imgs / t_e2e, 2), "unit": "images/s",
                    "h2d_bytes_per_step": host.x.numel() * 4 + 2 * host.t.numel() * 4,
                    "d2h_bytes_per_step": 4 * 6},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {
                "bound": "tensor", "kernel": "conv_tf32_kernel (hg_conv2d_fwd): forward convolution of every G/D "
                                             "layer of the step, flop-weighted; each layer timed as a CUDA-graph "
                                             "replay over operand sets > L2 (device time, no host issue gaps)",
                "achieved": round(conv_tf, 1), "peak": bf16_peak,
                "unit": "TFLOP/s", "frac": round(conv_tf / bf16_peak, 4),
                # dram__bytes_read+write of ONE launch of this kernel on the 32->32 3x3 @256^2 layer
                # (ncu --set full, profiles/r02_conv_fwd_32ch_256_ncu.md: 268.5 MB read + 222.6 MB written at
                # capture time, the rest of y still in L2); algorithmic = 536.9 MB
                "traffic": 491.1e6, "traffic_layer": "32->32 3x3 @256^2, batch 32: algorithmic 536.9e6 B "
                                                     "(x read once + y written once)",
                "peak_kind": peak_kind,
                "note": "operands are TF32 (half the bf16 rate): fraction of the TF32 ceiling = 2x frac",
                "dgrad_tflops": round(conv_agg["dgrad"], 1), "wgrad_tflops": round(conv_agg["wgrad"], 1),
                "all_three_tflops": round(conv_agg["all"], 1), "pass_us": conv_agg["pass_us"],
                "per_layer_columns": ["layer", "fwd TF/s", "dgrad TF/s", "wgrad TF/s", "fwd us", "dgrad us",
                                      "wgrad us", "fwd algorithmic GB/s (x + y + w once; HBM peak in hbm_peak_gbs)"],
                "hbm_peak_gbs": hbm_peak,
                "per_layer_tflops": conv_rows,
                "step_algorithmic_conv_tflop": round(step_flops / 1e12, 3),
                "step_achieved_tflops": round(step_flops / step_s / 1e12, 1)},
            "hist_block": hist_info,
        }
        if dv.world == 1:
            out["cpu_baseline"] = cpu_baseline_train()
        emit(out)
    dv.close()


# ---------------------------------------------------------- rehisto workload --
RH_BATCH = 16            # BASELINE config C5: ReHistoGAN step, 256^2, batch 16
RH_ALPHA, RH_BETA, RH_GAMMA = 32.0, 1.5, 4.0


class RecolorLoader:
    def __init__(self, rank, dev=None):
        g = torch.Generator().manual_seed(200 + rank)
        self.x = torch.rand(RH_BATCH, 3, S, S, generator=g)
        t = torch.rand(RH_BATCH, 3, H_BINS, H_BINS, generator=g)
        self.t = t / t.sum(dim=(1, 2, 3), keepdim=True)
        if torch.cuda.is_available():
            self.x, self.t = self.x.pin_memory(), self.t.pin_memory()
        if dev is not None:
            self.x, self.t = self.x.to(dev), self.t.to(dev)

    def __iter__(self):
        return self

    def __next__(self):
        return {"images": self.x, "histograms": self.t}


def rehisto_step_flops(B):
    """algorithmic conv FLOPs of one recoloringTrainer.train step (no GP): encoder-decoder +
    head forward x2 (the D-phase pass has no backward) + backward (2x), D fwd x3 + D bwd x2."""
    c = CAPACITY
    enc = [c * 2 ** i for i in range(7)]                      # 16 .. 1024
    f = _conv_flops(B, 3, c, 3, S)                            # mapping
    for i, (ci, co) in enumerate(zip(enc[:-1], enc[1:])):
        h = S // 2 ** i
        f += _conv_flops(B, ci, co, 1, h) + _conv_flops(B, ci, co, 3, h) + _conv_flops(B, co, co, 3, h)
        f += _conv_flops(B, co, co, 3, h // 2)
    dec = enc[::-1][:5]                                        # 1024 .. 64
    for i, (ci, co) in enumerate(zip(dec[:-1], dec[1:])):
        h = 4 * 2 ** i
        f += _conv_flops(B, ci, ci, 3, h) + _conv_flops(B, 2 * ci, co, 3, h) + _conv_flops(B, ci, co, 1, h)
        f += _conv_flops(B, co, co, 3, h) + _conv_flops(B, co, 3, 1, h)
    f += _conv_flops(B, dec[-1], 8 * c, 1, 64)                 # decoder_mapping
    f += _conv_flops(B, 4 * c, 4 * c, 3, 128) + _conv_flops(B, 2 * c, 2 * c, 3, 256)     # skip mod-convs
    for ci, co, h in ((8 * c, 4 * c, 128), (4 * c, 2 * c, 256)):                         # head
        f += _conv_flops(B, ci, co, 3, h) + _conv_flops(B, co, co, 3, h) + _conv_flops(B, co, 3, 1, h)
    d = sum(_conv_flops(B, ci, co, k, h // s) for n, ci, co, k, s, h in conv_layer_table() if n == "D")
    return f * (2 + 2) + d * (3 + 2 * 2), f, d


def make_cpu_rehisto_step():
    """closure: one ReHistoGAN step (D phase + G phase with all four loss terms; no GP, no
    optimiser update) at B=1 on the CPU with the oracle's reference-style arithmetic."""
    from histogan_b200 import rehistogan as rh
    from histogan_b200.gan import Discriminator, HistVectorizer
    from oracle import gan_oracle as go
    from oracle import rehisto_oracle as ro

    def sd_of(m, seed):
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        return {k: v.requires_grad_(True) for k, v in go.seeded_state_dict(shapes, seed).items()}

    with torch.device("meta"):
        mods = (rh.RecoloringEncoderDecoder(S, CAPACITY, skip_conn_to_GAN=True), HistVectorizer(H_BINS, 512, 8),
                rh.RecoloringGAN(S, 512, CAPACITY), Discriminator(S, CAPACITY))
    sd_ed, sd_h, sd_g, sd_d = [sd_of(m, i + 11) for i, m in enumerate(mods)]
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(min(avail, 32))
    ld = RecolorLoader(0)
    x, t = ld.x[:1].clone(), ld.t[:1].clone()
    nz = torch.rand(1, S, S, 1)

    def step():
        with torch.no_grad():
            fake = ro.g_phase(sd_ed, sd_h, sd_g, sd_d, x, t, nz, S, variance=False)["generated"]
        d_loss = (F.relu(1 + go.discriminator(sd_d, x, S)) + F.relu(1 - go.discriminator(sd_d, fake, S))).mean()
        torch.autograd.grad(d_loss, list(sd_d.values()))
        out = ro.g_phase(sd_ed, sd_h, sd_g, sd_d, x, t, nz, S, RH_ALPHA, RH_BETA, RH_GAMMA)
        params = [v for k, v in list(sd_ed.items()) + list(sd_h.items()) + list(sd_g.items())
                  if "conv_out_rgb" not in k]
        torch.autograd.grad(out["gen_loss"], params)
    return step


def run_rehisto(args):
    from histogan_b200 import _lib
    from histogan_b200.rehistogan import recoloringTrainer
    dv = Dist(args.gpus)
    lib = _lib.load()
    _lib.check(lib.hg_device_check(dv.local_rank), "hg_device_check")
    out_dir = os.path.join(ROOT, "gpurun_out", f"bench_rehisto_rank{dv.rank}")
    torch.manual_seed(4321 + dv.rank)
    tr = recoloringTrainer("bench", os.path.join(out_dir, "results"), os.path.join(out_dir, "models"),
                           image_size=S, network_capacity=CAPACITY, batch_size=RH_BATCH,
                           gradient_accumulate_every=1, skip_conn_to_GAN=True, initialize_gan=True,
                           save_every=10 ** 9, fast_rng=True,
                           cuda_graphs=os.environ.get("HG_CUDA_GRAPHS", "1") != "0")
    sampler = ClockSampler(dv.local_rank) if dv.rank == 0 else None

    def timed(loader, steps, warmup):
        tr.loader = loader
        tr.steps = FIRST_STEP - warmup
        for _ in range(warmup):
            tr.train(RH_ALPHA, RH_BETA, RH_GAMMA)
        dv.barrier()
        n0 = lib.hg_launch_count() + tr.graph_replayed_launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            tr.train(RH_ALPHA, RH_BETA, RH_GAMMA)
        e.record()
        dv.barrier()
        return (dv.max_over_ranks(s.elapsed_time(e) * 1e-3),
                lib.hg_launch_count() + tr.graph_replayed_launches - n0)

    timed(RecolorLoader(dv.rank, dv.dev), 4, 2)        # allocator warm-up + graph capture (both D variants)
    t_dev, launches = timed(RecolorLoader(dv.rank, dv.dev), args.steps, args.warmup)
    host = RecolorLoader(dv.rank)
    t_e2e, _ = timed(host, args.steps, 1)
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30
    losses = {"d": tr.d_loss, "g": tr.g_loss, "h": tr.h_loss, "r": tr.r_loss, "v": tr.var_loss,
              "gp": tr.last_gp_loss}
    clocks = sampler.stop() if sampler else {}
    if dv.rank == 0:
        _, bf16_peak, _, peak_kind = load_peaks()
        n_imgs = RH_BATCH * dv.world * args.steps
        step_flops, ed_f, d_f = rehisto_step_flops(RH_BATCH)
        step_s = t_dev / args.steps
        out = {
            "metric": "training images/sec", "value": round(n_imgs / t_dev, 2), "unit": "images/s",
            "n_gpus": dv.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_s * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32 operands (round-to-nearest) / fp32 accumulate + fp32",
            "data": "synthetic",
            "config": {"workload": "ReHistoGAN recoloringTrainer.train step (D phase + G phase: adversarial "
                                   "+ histogram + laplacian reconstruction + variance loss, GP every 4th "
                                   "step, DiffGrad), 256x256, network_capacity=16, batch 16 per GPU, "
                                   "skip_conn_to_GAN, hist 'sampling'",
                       "global_batch": RH_BATCH * dv.world,
                       "parallelism": f"dp{dv.world}" + (" + NCCL grad all-reduce" if dv.world > 1 else ""),
                       "l2_flush": "not needed: the step's working set >> 126 MB L2",
                       "timed_steps": f"trainer.steps {FIRST_STEP}..{FIRST_STEP + args.steps - 1}",
                       "cuda_graphs": bool(tr.cuda_graphs), "peak_mem_gib": round(mem_gb, 2),
                       "final_losses": losses},
            "e2e": {"value": round(n_imgs / t_e2e, 2), "unit": "images/s",
                    "h2d_bytes_per_step": 2 * (host.x.numel() + host.t.numel()) * 4,
                    "d2h_bytes_per_step": 4 * 6},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "whole step: algorithmic conv FLOPs / step time "
                                                      "(a lower bound of the conv kernels' own rate; the "
                                                      "per-layer kernel figures are in the train workload)",
                         "achieved": round(step_flops / step_s / 1e12, 1), "peak": bf16_peak,
                         "unit": "TFLOP/s", "frac": round(step_flops / step_s / 1e12 / bf16_peak, 4),
                         "traffic": None, "peak_kind": peak_kind,
                         "step_algorithmic_conv_tflop": round(step_flops / 1e12, 3)},
        }
        if dv.world == 1:
            step = make_cpu_rehisto_step()
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": round(1.0 / dt, 4), "unit": "images/s",
                                   "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": "1 image (of the 16-image batch) for 1 step, 256x256, capacity 16"}
        emit(out)
    dv.close()


# ------------------------------------------------------- reference / CPU arm --

def cpu_hist_step(x, t, insz=256):
    from oracle import hist_oracle as ho
    return ho.hist_loss_and_grad(x, t, ALPHA, h=H_BINS, insz=insz)


def pick_cpu_threads(fn):
    """all the host threads it can USE: torch's intra-op pool oversubscribes badly on
    many-core hosts for these element-wise + skinny-GEMM ops, so calibrate once."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best, best_dt = 1, float("inf")
    for n in sorted({avail, max(1, avail // 2), 32, 16, 8}):
        if n > avail:
            continue
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best, best_dt = n, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline_hist(sample_images=2, reps=2):
    """the reference's histogram algorithm (torch-CPU restatement) on a bounded sample."""
    x, t = make_inputs(0, B=sample_images)
    pick_cpu_threads(lambda: cpu_hist_step(x[:1], t[:1]))
    cpu_hist_step(x, t)
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_hist_step(x, t)
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(sample_images / dt, 3), "unit": "images/s",
            "cores": torch.get_num_threads(), "kind": "port",
            "what": "RGBuvHistBlock fwd + Hellinger + bwd (the hist workload; the train-step CPU "
                    "baseline is `--impl reference`)",
            "sample": f"{sample_images} of the 32 images per step (256x256, insz=256), {reps} reps "
                      f"after 1 warm-up"}


class CpuTrainer:
    """the reference's train step on the CPU (oracle/train_oracle.py: the restatement pinned
    against the unmodified reference Trainer.train by tests/test_train_oracle.py): D phase with the
    gradient penalty every 4th step, G phase with the histogram loss and the path-length
    regulariser every 32nd step, DiffGrad updates -- the same loss terms and schedule as the GPU
    arm, on a bounded sample of `B` images per step."""

    def __init__(self, B=1):
        from histogan_b200.gan import Discriminator, Generator, HistVectorizer, StyleVectorizer
        from oracle import gan_oracle as go

        def sd_of(m, seed):
            shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
            return {k: v.requires_grad_(True) for k, v in go.seeded_state_dict(shapes, seed).items()}

        with torch.device("meta"):
            mods = (Generator(S, 512, CAPACITY), Discriminator(S, CAPACITY), StyleVectorizer(512, 8),
                    HistVectorizer(H_BINS, 512, 8))
        self.sd_g, self.sd_d, self.sd_s, self.sd_h = [sd_of(m, i + 1) for i, m in enumerate(mods)]
        self.B = B
        _, self.t = make_inputs(0, B=B)
        self.x = torch.rand(B, 3, S, S)
        self.opt_d, self.opt_g, self.pl_mean = {}, {}, 0.0
        self.steps = 0

    def step(self):
        from oracle import train_oracle as to
        k = self.steps
        dr = to.draw_step_inputs(self.B, 5, 512, S, path_penalty=k % 32 == 0)
        d = to.d_phase(self.sd_g, self.sd_d, self.sd_s, self.sd_h, self.x, self.t, dr["d_style"],
                       dr["d_noise"], S, apply_gp=k % 4 == 0)
        names = list(self.sd_d)
        to.diffgrad_step_tensors([self.sd_d[n] for n in names], [d["grads"][n] for n in names], self.opt_d)
        g = to.g_phase(self.sd_g, self.sd_d, self.sd_s, self.sd_h, self.t, dr["g_style"], dr["g_noise"], S,
                       ALPHA, hist_kw=dict(h=H_BINS, insz=150), pl_noise=dr["pl_noise"], pl_mean=self.pl_mean)
        named = [("G." + n, v) for n, v in self.sd_g.items()] + [("S." + n, v) for n, v in self.sd_s.items()] + \
                [("H." + n, v) for n, v in self.sd_h.items()]
        to.diffgrad_step_tensors([v for _, v in named], [g["grads"][n] for n, _ in named], self.opt_g)
        if g["avg_pl"] is not None:
            self.pl_mean = 0.99 * self.pl_mean + 0.01 * g["avg_pl"]
        self.steps += 1


def cpu_train_threads(tr):
    """thread count for the CPU arm, calibrated on the discriminator forward+backward of the
    sample (torch's intra-op pool oversubscribes on many-core hosts)"""
    from oracle import gan_oracle as go

    def probe():
        out = go.discriminator(tr.sd_d, tr.x, S)
        torch.autograd.grad(out.sum(), list(tr.sd_d.values()), allow_unused=True)
    return pick_cpu_threads(probe)


def run_cpu_train(steps, warmup):
    """(images/s, threads, description) of the CPU arm over `steps` timed steps"""
    tr = CpuTrainer(1)
    threads = cpu_train_threads(tr)
    tr.steps = window_start(steps) - warmup
    for _ in range(warmup):
        tr.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step()
    dt = time.perf_counter() - t0
    first = window_start(steps)
    kinds = [("gp+pl" if k % 32 == 0 else "gp" if k % 4 == 0 else "plain") for k in range(first, first + steps)]
    return steps * tr.B / dt, threads, dt / steps, kinds


def cpu_baseline_train(steps=3, warmup=1):
    val, threads, _, kinds = run_cpu_train(steps, warmup)
    return {"value": round(val, 4), "unit": "images/s", "cores": threads, "kind": "port",
            "what": "HistoGAN train step, same loss terms as the GPU arm (D: hinge + gradient penalty; G: "
                    "adversarial + histogram loss + path-length regulariser; DiffGrad updates), reference "
                    "arithmetic (per-sample weights + grouped convs, float64 soft-binning)",
            "sample": f"1 image (of the 32-image batch) per step, {steps} steps ({', '.join(kinds)}) after "
                      f"{warmup} warm-up, 256x256, capacity 16"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "hist":
        x, t = make_inputs(0, B=2)
        pick_cpu_threads(lambda: cpu_hist_step(x[:1], t[:1]))
        step, sample, what = (lambda: cpu_hist_step(x, t)), 2, \
            "RGBuvHistBlock fwd + Hellinger loss + bwd, 256x256, h=64, insz=256"
        warm = 1
    elif args.workload == "rehisto":
        step, sample, what = make_cpu_rehisto_step(), 1, \
            ("ReHistoGAN train step (D + G phase, four loss terms; no GP/optimiser), 256x256, capacity 16; "
             "CPU restatement of the reference")
        warm = 0
    else:
        steps, warm = max(3, min(args.steps, 4)), max(1, min(args.warmup, 1))
        val, threads, per_step, kinds = run_cpu_train(steps, warm)
        val = round(val, 4)
        emit({
            "impl": "reference", "metric": "training images/sec", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(per_step * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 soft-binning, as the reference)", "data": "synthetic",
            "config": {"workload": "HistoGAN Trainer.train step on the host CPU: same loss terms and schedule "
                                   "as the GPU arm (gradient penalty every 4th, path-length every 32nd step, "
                                   "DiffGrad), 256x256, network_capacity=16; oracle/train_oracle.py (pinned "
                                   "against the unmodified reference Trainer.train)",
                       "global_batch": 1, "timed_steps": kinds},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"1 image per step (bounded sample of the 32-image batch), {steps} "
                                       f"steps after {warm} warm-up"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return
    steps = max(1, min(args.steps, 3))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = round(sample * steps / dt, 4)
    out = {
        "impl": "reference", "metric": "images/sec" if args.workload == "hist" else "training images/sec",
        "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": round(dt / steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (f64 soft-binning, as the reference)", "data": "synthetic",
        "config": {"workload": what, "global_batch": sample},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": f"{sample} image(s) per step (bounded sample of the "
                                                   f"32-image batch), {steps} step(s)"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


# ------------------------------------------- reference's own eager GPU path --

def _import_reference_gpu():
    """the UNMODIFIED reference modules from baseline/_ref (staged by oracle/make_baseline_ref.py),
    with the three un-vendored packages stubbed; DiffGrad = a pure-torch restatement (no kernel of
    this repo is on that path)."""
    import types
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "histoGAN", "histoGAN.py")):
        return None

    class TorchDiffGrad(torch.optim.Optimizer):
        def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
            super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

        @torch.no_grad()
        def step(self):
            import math
            for gr in self.param_groups:
                b1, b2 = gr["betas"]
                for p in gr["params"]:
                    if p.grad is None:
                        continue
                    st = self.state[p]
                    if not st:
                        st["step"] = 0
                        st["m"], st["v"], st["g"] = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
                    st["step"] += 1
                    g = p.grad
                    st["m"].mul_(b1).add_(g, alpha=1 - b1)
                    st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
                    xi = torch.sigmoid((st["g"] - g).abs())
                    st["g"] = g.clone()
                    ss = gr["lr"] * math.sqrt(1 - b2 ** st["step"]) / (1 - b1 ** st["step"])
                    p.addcdiv_(st["m"] * xi, st["v"].sqrt().add_(gr["eps"]), value=-ss)

    class _Missing:
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")

    for name, attrs in (("torch_optimizer", {"DiffGrad": TorchDiffGrad}),
                        ("vector_quantize_pytorch", {"VectorQuantize": _Missing}),
                        ("linear_attention_transformer", {"ImageLinearAttention": _Missing})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    # `utils`, `histoGAN`, `histogram_classes` must resolve to the reference's packages
    sys.path.insert(0, ref)
    import importlib
    return importlib.import_module("histoGAN.histoGAN"), importlib.import_module("histogram_classes.RGBuvHistBlock")


def run_reference_gpu(args):
    """CONTEXT arm (not the driver's `--impl reference`): the reference's own eager-PyTorch GPU path
    on the same B200 -- RGBuvHistBlock(device='cuda') at C2 and Trainer.train at C3."""
    mods = _import_reference_gpu()
    if mods is None:
        emit({"impl": "reference-gpu", "unavailable": "baseline/_ref not staged (python -m oracle.make_baseline_ref)"})
        return
    gm, hm = mods
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {"impl": "reference-gpu", "n_gpus": 1, "data": "synthetic",
           "note": "unmodified reference code (baseline/_ref), eager PyTorch + cuDNN/cuBLAS on the same GPU; "
                   "torch_optimizer.DiffGrad replaced by a pure-torch restatement"}
    # ---- C2: histogram block fwd + Hellinger + bwd, 32 x 3 x 256 x 256, insz 256 and 150
    x, t = make_inputs(0, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    for insz in (256, 150):
        blk = hm.RGBuvHistBlock(h=H_BINS, insz=insz, device="cuda")

        def step():
            xg = x.detach().requires_grad_(True)
            hist = blk(F.relu(xg))
            loss = ALPHA * (1 / 2 ** 0.5) * torch.sqrt(torch.sum((torch.sqrt(t) - torch.sqrt(hist)) ** 2)) / t.shape[0]
            loss.backward()
        dt = time_call(step, flush, reps=5)
        out[f"hist_insz{insz}"] = {"ms_per_step": round(dt * 1e3, 3), "images_per_s": round(B_PER_GPU / dt, 2),
                                   "us_per_image": round(dt / B_PER_GPU * 1e6, 2)}
    del x, t, flush
    torch.cuda.empty_cache()
    # ---- C3: Trainer.train, 256^2, capacity 16, batch 32 (smaller if it does not fit)
    out_dir = os.path.join(ROOT, "gpurun_out", "bench_refgpu")
    for B in (32, 16, 8):
        try:
            torch.manual_seed(1234)
            tr = gm.Trainer("ref", os.path.join(out_dir, "results"), os.path.join(out_dir, "models"), image_size=S,
                            network_capacity=CAPACITY, batch_size=B, gradient_accumulate_every=1, hist_insz=150,
                            hist_resizing="interpolation", save_every=10 ** 9)
            xb = torch.rand(B, 3, S, S, device=dev)
            _, tb = make_inputs(0, B=B, device=dev)
            batch = {"images": xb, "histograms": tb}
            tr.loader = iter(lambda: batch, None)
            tr.loader_evaluate = iter(lambda: {"histograms": tb[:4]}, None)
            steps = max(3, min(args.steps, 8))
            first = window_start(steps)
            tr.steps = first - 2
            for _ in range(2):
                tr.train(alpha=ALPHA)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                tr.train(alpha=ALPHA)
            e.record()
            torch.cuda.synchronize()
            dt = s.elapsed_time(e) * 1e-3
            out.update({"metric": "training images/sec", "value": round(B * steps / dt, 2), "unit": "images/s",
                        "steps": steps, "warmup": 2, "ms_per_step": round(dt / steps * 1e3, 2),
                        "higher_is_better": True, "dtype": "fp32 (cuDNN/cuBLAS defaults of torch 2.11)",
                        "config": {"workload": "reference Trainer.train, 256x256, network_capacity=16, "
                                               f"batch {B}, hist_insz=150 interpolation",
                                   "global_batch": B,
                                   "timed_steps": f"trainer.steps {first}..{first + steps - 1}",
                                   "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}})
            break
        except torch.cuda.OutOfMemoryError:
            out.setdefault("oom_at_batch", []).append(B)
            del tr
            torch.cuda.empty_cache()
    emit(out)


_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner, torchrun children) write to fd 1; the contract is ONE
    JSON line on stdout.  Route fd 1 to stderr for the run and keep the real stdout aside."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj)
    if _REAL_STDOUT is not None:
        _REAL_STDOUT.write(line + "\n")
        _REAL_STDOUT.flush()
    else:
        print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="train", choices=["train", "hist", "rehisto"])
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        args.warmup = max(args.warmup, 3)
        {"train": run_train, "hist": run_hist, "rehisto": run_rehisto}[args.workload](args)


if __name__ == "__main__":
    main()
