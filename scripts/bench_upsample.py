"""GPU micro-benchmark of the 2x up-sample + modulate kernels at the generator's six sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200 import _lib
lib = _lib.load()
B = 32
st = _lib.current_stream_ptr(torch.device("cuda", 0))
SIZES = ((2048, 4), (1024, 8), (512, 16), (256, 32), (128, 64), (64, 128))
if len(sys.argv) > 1:
    SIZES = SIZES[-int(sys.argv[1]):]
for C, H in SIZES:
    x = torch.randn(B, H, H, C, device="cuda"); mod = torch.rand(B, C, device="cuda") + 0.5
    xm = torch.empty(B, 2 * H, 2 * H, C, device="cuda"); dxm = torch.randn_like(xm)
    dx = torch.empty_like(x); gm = torch.empty_like(mod)
    fw = lambda: lib.hg_upsample_modulate_round(_lib.ptr(x), _lib.ptr(mod), _lib.ptr(xm), B, H, H, C, st)
    bw = lambda: lib.hg_upsample_modulate_bwd(_lib.ptr(dxm), _lib.ptr(x), _lib.ptr(mod), _lib.ptr(dx), _lib.ptr(gm), B, H, H, C, st)
    for name, fn, byts in (("fwd", fw, xm.numel() * 4 + x.numel() * 4), ("bwd", bw, xm.numel() * 4 + 2 * x.numel() * 4)):
        for _ in range(3): fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize()
        t = s.elapsed_time(e) / 10 * 1e3
        print(f"{name} C={C:5d} {H:3d}->{2*H:3d}: {t:7.1f} us  {byts / t / 1e3:7.1f} GB/s")
