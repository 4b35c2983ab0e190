"""GPU: device time of the image-input convolution kernels (conv_small.cu) and a few element-wise ops at the
benchmark's sizes, CUDA-graph replays over rotating operand sets.  usage: python scripts/bench_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200 import conv, fused
import bench

dev = torch.device("cuda", 0)


def timeit(name, fn, sets, nbytes):
    reps = 4 if sets > 1 else 8                            # launches per graph: every set, several times
    us = bench.graph_time([(lambda i=i: fn(i % sets)) for i in range(sets * reps)]) * 1e6
    print(f"{name:42s} {us:9.1f} us   {nbytes / us / 1e6:7.2f} TB/s (algorithmic)", flush=True)


for B in (32, 64):
    n = 3                                                  # operand sets (> L2 in total)
    imgs = [torch.randn(B, 3, 256, 256, device=dev) for _ in range(n)]
    dys = [torch.randn(B, 16, 256, 256, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(n)]
    for k in (3, 1):
        w = torch.randn(16, 3, k, k, device=dev) / (3 * k * k) ** 0.5
        bias = torch.randn(16, device=dev)
        out_b, in_b = B * 65536 * 16 * 4, B * 65536 * 3 * 4
        timeit(f"small fwd  k{k} B{B}", lambda i: conv.conv_small_fwd(imgs[i], w, 16, bias=bias, lrelu=True, round_tf32=True),
               n, out_b + in_b)
        timeit(f"small dgrad k{k} B{B}", lambda i: conv.conv_small_dgrad(dys[i], w, 3), n, out_b + in_b)
        timeit(f"small wgrad k{k} B{B}", lambda i: conv.conv_small_wgrad(dys[i], imgs[i], (16, 3, k, k)), n, out_b + in_b)
    del imgs, dys
B = 32
for C, S in ((32, 256), (64, 128), (128, 64), (256, 32), (512, 16), (1024, 8), (2048, 4)):
    n = 3
    xs = [torch.randn(B, C, S, S, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(n)]
    style = torch.randn(B, C, device=dev)
    w = torch.randn(3, C, 1, 1, device=dev)
    prev = torch.randn(B, 3, S, S, device=dev)
    with torch.no_grad():
        timeit(f"to_rgb fwd C{C} @{S}", lambda i: fused.to_rgb(xs[i], style, w, prev), n, B * S * S * C * 4)
    del xs
x = torch.randn(32, 12288, device=dev)
w = torch.randn(1024, 12288, device=dev) / 111
b = torch.randn(1024, device=dev)
timeit("linear 12288->1024 B32 (lrelu)", lambda i: fused.grouped_linear([x], [w], [b], fused.LIN_LRELU, slope=0.2), 1, w.numel() * 4)
