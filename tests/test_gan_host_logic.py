"""CPU: host logic of the G/D modules (activation-side modulation, demodulation as
an output scale, any-order conv autograd, state_dict layout) with the raw conv
kernels replaced by exact torch stand-ins -> must match the reference goldens to
fp32 accuracy.  With TF32-rounded stand-ins the same run measures the rounding
noise floor that the GPU tests' tolerances are derived from."""
import torch

from tests import gan_checks
from tests.emulation import emulated_conv


def test_generator_host_logic_exact():
    with emulated_conv(round_operands=False):
        e, _ = gan_checks.generator_errors("cpu")
    print(e)
    assert max(e.values()) < 2e-4, e


def test_discriminator_host_logic_exact():
    with emulated_conv(round_operands=False):
        e, _ = gan_checks.discriminator_errors("cpu")
    print(e)
    assert max(e.values()) < 2e-4, e


def test_tf32_noise_floor():
    """the same algorithm with TF32-rounded operands (what the tensor cores compute):
    activations stay within 1e-3 of the fp32 reference (north_star); gradients are noisier
    (cancellation): up to ~4e-2 on individual tensors.  DESIGN.md quotes this table."""
    with emulated_conv(round_operands=True):
        eg, _ = gan_checks.generator_errors("cpu")
        ed, _ = gan_checks.discriminator_errors("cpu")
    print("TF32 noise floor G:", {k: f"{v:.2e}" for k, v in eg.items()})
    print("TF32 noise floor D:", {k: f"{v:.2e}" for k, v in ed.items()})
    assert eg["rgb"] < 1e-3 and eg["act_last"] < 1e-3 and ed["logits"] < 1e-3
    assert max(eg.values()) < 1e-1 and max(ed.values()) < 1e-1
