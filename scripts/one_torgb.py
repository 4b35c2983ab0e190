"""GPU: a few launches of torgb_fwd at 32 channels / 256^2 / batch 32 for an ncu capture"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200 import fused
dev = torch.device("cuda", 0)
B, C, S = 32, 32, 256
x = torch.randn(B, C, S, S, device=dev).contiguous(memory_format=torch.channels_last)
style = torch.randn(B, C, device=dev); w = torch.randn(3, C, 1, 1, device=dev); prev = torch.randn(B, 3, S, S, device=dev)
with torch.no_grad():
    for _ in range(4):
        fused.to_rgb(x, style, w, prev)
torch.cuda.synchronize()
