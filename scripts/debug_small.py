"""GPU: constant-memory 3->16 kernels against the generic ones and torch fp64 at the D-256 test's shapes."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from histogan_b200 import conv
torch.manual_seed(0)
dev = torch.device("cuda", 0)
mode = "generic" if os.environ.get("HG_SMALL_GENERIC") == "1" else "fast"
for k in (3, 1):
    x = torch.rand(2, 3, 256, 256, device=dev)
    w = torch.randn(16, 3, k, k, device=dev) * 0.7
    b = torch.randn(16, device=dev)
    res = torch.randn(2, 16, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(2, 16, 256, 256, device=dev).contiguous(memory_format=torch.channels_last)
    for name, kw in (("bias+lrelu+round", dict(bias=b, lrelu=True, round_tf32=True)), ("plain", dict()),
                     ("bias+res", dict(bias=b, residual=res))):
        y = conv.conv_small_fwd(x, w, 16, **kw)
        ref = F.conv2d(x.double(), w.double(), kw.get("bias").double() if "bias" in kw else None, padding=k // 2)
        if kw.get("lrelu"):
            ref = F.leaky_relu(ref, 0.2)
        if "residual" in kw:
            ref = ref + res.double()
        err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
        print(mode, f"fwd k{k} {name}: max rel err {err:.2e}", "sum", float(y.double().sum()))
    dx = conv.conv_small_dgrad(dy, w, 3)
    dref = torch.nn.grad.conv2d_input((2, 3, 256, 256), w.double(), dy.double(), padding=k // 2)
    print(mode, f"dgrad k{k}: max rel err {(dx.double() - dref).abs().max().item() / dref.abs().max().item():.2e}")
    # two different filters back to back (the constant bank is rewritten between the launches)
    w2 = torch.randn(16, 3, k, k, device=dev)
    ya = conv.conv_small_fwd(x, w, 16); yb = conv.conv_small_fwd(x, w2, 16); ya2 = conv.conv_small_fwd(x, w, 16)
    print(mode, f"k{k} back-to-back filters: first == third {torch.equal(ya, ya2)}, differs from second {not torch.equal(ya, yb)}")
