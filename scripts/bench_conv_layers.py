"""GPU: per-layer CUDA-event times of the three conv primitives (forward, input gradient, weight
gradient) on every G / D layer shape of the bench config (256^2, capacity 16, batch 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from histogan_b200 import conv, ops

B = 32
dev = torch.device("cuda", 0)


def t_us(fn, reps=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print(f"{'layer':28s} {'GF':>6s} | {'fwd us':>8s} {'TF/s':>6s} | {'dgrad us':>8s} {'TF/s':>6s} | {'wgrad us':>8s} {'TF/s':>6s} | HBM floor us")
for net, ci, co, k, s, h in bench.conv_layer_table():
    if co == 3:
        continue
    cip, cop = ops._round_up(ci), ops._round_up(co)
    x = conv.tf32_round(torch.randn(B, cip, h, h, device=dev)).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5).contiguous(memory_format=torch.channels_last)
    oh = h // s
    dy = conv.tf32_round(torch.randn(B, cop, oh, oh, device=dev)).contiguous(memory_format=torch.channels_last)
    wp0, wp1 = conv.pack_weight(w, 0), conv.pack_weight(w, 1)
    pad = k // 2
    f = bench._conv_flops(B, ci, co, k, oh)
    tf = t_us(lambda: conv.conv2d_nhwc(x, wp0, s, pad, cout=cop))
    td = t_us(lambda: ops._raw_grad_input(dy, w, s, pad, (h, h), dy_rounded=True, padded_io=True))
    tw = t_us(lambda: conv.conv2d_wgrad_nhwc(dy, x, k, s, pad))
    floor = (x.numel() + dy.numel()) * 4 / 6.6e12 * 1e6
    tot["fwd"] += tf; tot["dgrad"] += td; tot["wgrad"] += tw
    print(f"{net} {ci:4d}->{co:4d} k{k} s{s} @{h:3d}   {f/1e9:6.1f} | {tf:8.1f} {f/tf/1e6:6.1f} | {td:8.1f} {f/td/1e6:6.1f} | {tw:8.1f} {f/tw/1e6:6.1f} | {floor:6.1f}")
    del x, w, dy, wp0, wp1
print("totals (us):", {k: round(v) for k, v in tot.items()})
