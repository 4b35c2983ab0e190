"""GPU: block 0 of the seeded 256^2 discriminator, stage by stage, against torch fp64 (same operands)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from histogan_b200 import ops, conv
from histogan_b200.gan import Discriminator
from oracle import gan_oracle as go, make_golden_gan as mg
dev = torch.device("cuda", 0)
mode = "generic" if os.environ.get("HG_SMALL_GENERIC") == "1" else "fast"
D = Discriminator(mg.IMAGE_SIZE_L, network_capacity=mg.CAPACITY)
D.load_state_dict(go.seeded_state_dict({k: list(v.shape) for k, v in D.state_dict().items()}, seed=2))
D.to(dev)
x = mg.gan_inputs(mg.IMAGE_SIZE_L, mg.B_L, seed=5)["images"].to(dev)
blk = D.blocks[0]
c1, c2, cr, dn = blk.net[0], blk.net[2], blk.conv_res, blk.downsample


def rel(a, b):
    return ((a.double() - b).norm() / b.norm()).item()


with torch.no_grad():
    for rep in range(2):
        t1 = ops.conv_bias_act(x, c1.weight, c1.bias, None, 1, 1, act=True, x_rounded=True, round_out=True)
        r1 = F.leaky_relu(F.conv2d(x.double(), c1.weight.double(), c1.bias.double(), padding=1), 0.2)
        t2 = ops.conv_bias_act(t1, c2.weight, c2.bias, None, 1, 1, act=True, x_rounded=True, round_out=False)
        r2 = F.leaky_relu(F.conv2d(t1.double(), conv.tf32_round(c2.weight.detach()).double(), c2.bias.double(), padding=1), 0.2)
        y = ops.conv_bias_act(x, cr.weight, cr.bias, t2, 1, 0, act=False, x_rounded=True, round_out=True)
        ry = F.conv2d(x.double(), cr.weight.double(), cr.bias.double()) + t2.double()
        print(mode, rep, f"c1 (3x3 small, rounded out): {rel(t1, r1):.2e}   c2 (tensor core, same operands): {rel(t2, r2):.2e}   "
              f"conv_res (1x1 small + residual, rounded out): {rel(y, ry):.2e}", "| max|w1|", float(c1.weight.abs().max()),
              "max|t1|", float(t1.abs().max()))
    # unrounded outputs of the two small convs right after each other (the constant bank changes in between)
    a = conv.conv_small_fwd(x, c1.weight, 16, bias=c1.bias)
    b = conv.conv_small_fwd(x, cr.weight, 16, bias=cr.bias)
    a2 = conv.conv_small_fwd(x, c1.weight, 16, bias=c1.bias)
    ra = F.conv2d(x.double(), c1.weight.double(), c1.bias.double(), padding=1)
    rb = F.conv2d(x.double(), cr.weight.double(), cr.bias.double())
    print(mode, f"alternating 3x3 / 1x1 / 3x3, unrounded: {rel(a, ra):.2e} {rel(b, rb):.2e} {rel(a2, ra):.2e}")
    logits, _ = D(x)
    g = mg  # golden compare
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gan_discriminator_256.npz"))
    ref = torch.from_numpy(z["logits"]).double().to(dev).reshape(logits.shape)
    print(mode, "logits", logits.flatten().tolist(), "golden", ref.flatten().tolist(), "rel", rel(logits, ref))
