"""GPU: a few launches of single conv primitives for ncu captures.
usage: python scripts/one_conv.py fwd|wgrad [Cin Cout H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200 import conv
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
ci, co, h = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (32, 32, 256)
B, dev = 32, torch.device("cuda", 0)
x = conv.tf32_round(torch.randn(B, ci, h, h, device=dev)).contiguous(memory_format=torch.channels_last)
w = (torch.randn(co, ci, 3, 3, device=dev) / (ci * 9) ** 0.5).contiguous(memory_format=torch.channels_last)
dy = conv.tf32_round(torch.randn(B, co, h, h, device=dev)).contiguous(memory_format=torch.channels_last)
wp = conv.pack_weight(w, 0)
if what.startswith("small"):
    # the discriminator's image-input layer (conv_small.cu): planar 3-channel image -> 16 channels
    img = torch.randn(B, 3, h, h, device=dev)
    ws = torch.randn(16, 3, 3, 3, device=dev) / 27 ** 0.5
    dys = torch.randn(B, 16, h, h, device=dev).contiguous(memory_format=torch.channels_last)
    for _ in range(4):
        conv.conv_small_fwd(img, ws, 16, lrelu=True, round_tf32=True)
        conv.conv_small_dgrad(dys, ws, 3)
        conv.conv_small_wgrad(dys, img, (16, 3, 3, 3))
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(4):
    if what == "fwd":
        conv.conv2d_nhwc(x, wp, 1, 1, cout=co)
    else:
        conv.conv2d_wgrad_nhwc(dy, x, 3, 1, 1)
torch.cuda.synchronize()
