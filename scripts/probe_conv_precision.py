"""GPU probe: accumulation accuracy of the tcgen05 TF32 conv vs exact arithmetic on the
same TF32-rounded operands (and cuDNN fp32 / TF32 for context)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from histogan_b200 import conv

torch.manual_seed(0)
for (B, Cin, S, Cout, k) in [(4, 64, 16, 64, 3), (4, 256, 16, 256, 3), (2, 1024, 8, 512, 3), (4, 512, 16, 128, 1)]:
    x = conv.tf32_round(torch.randn(B, Cin, S, S)).cuda()
    w = conv.tf32_round(torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5).cuda()
    ref = F.conv2d(x.double(), w.double(), padding=k // 2)
    y = conv.conv2d_nhwc(x, conv.pack_weight(w, 0), 1, k // 2)
    torch.backends.cudnn.allow_tf32 = False
    y32 = F.conv2d(x, w, padding=k // 2)
    torch.backends.cudnn.allow_tf32 = True
    ytf = F.conv2d(x, w, padding=k // 2)
    rel = lambda a: ((a.double() - ref).norm() / ref.norm()).item()
    bias = lambda a: ((a.double() - ref) * ref.sign()).mean().item() / ref.abs().mean().item()
    print(f"Cin={Cin} k={k} K={Cin*k*k}: ours rel {rel(y):.2e} (signed bias {bias(y):+.2e}) | cudnn fp32 {rel(y32):.2e} | cudnn tf32 {rel(ytf):.2e} (bias {bias(ytf):+.2e})")
# unrounded operands: what the hardware does with the low 13 bits
x = torch.randn(4, 256, 16, 16).cuda(); w = (torch.randn(256, 256, 3, 3) / 48).cuda()
ref = F.conv2d(x.double(), w.double(), padding=1)
wp = torch.empty(256, 3, 3, 256, device="cuda"); wp.copy_(w.permute(0, 2, 3, 1))
y = conv.conv2d_nhwc(x, wp.contiguous(), 1, 1)
print("unrounded operands: rel", ((y.double() - ref).norm() / ref.norm()).item(),
      "signed bias", (((y.double() - ref) * ref.sign()).mean() / ref.abs().mean()).item())
