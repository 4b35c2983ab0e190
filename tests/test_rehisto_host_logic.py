"""CPU: the ReHistoGAN modules of histogan_b200 (composed path, conv primitives emulated by
torch) against the reference-made golden vectors -- checks structure / state_dict / autograd
wiring without a GPU."""
import torch

from tests import emulation, rehisto_checks as rc


def test_rehisto_modules_match_reference_golden_fp32_emulation():
    with emulation.emulated_conv(round_operands=False):
        e, _ = rc.g_phase_errors(torch.device("cpu"), rc.oracle_losses)
    print({k: f"{v:.1e}" for k, v in e.items()})
    for k in ("latent", "p1", "p2", "ed_rgb", "generated"):
        assert e[k] < 2e-5, (k, e[k])
    for k in ("d_loss", "hist_loss", "rec_loss", "var_loss"):
        assert e[k] < 5e-5, (k, e[k])
    for k in ("dgen_d", "dgen_hist", "dgen_rec", "dgen_var"):
        assert e[k] < 1e-4, (k, e[k])
    # the seeded 4-level instance-norm encoder amplifies perturbations ~500x (TF32 rounding, 3e-4,
    # moves these gradients by 1e-1: see the GPU test's emulation row); the channels_last weight
    # storage makes torch's CPU kernels differ from the reference run by ~1e-6, hence 5e-3 here --
    # any wiring error shows up as O(1)
    assert e["param_grads_max"] < 5e-3, e["param_grads_max"]


def test_trainer_constructor_mirrors_reference_defaults():
    import inspect
    from histogan_b200.rehistogan import recoloringTrainer, recoloringGAN
    sig = inspect.signature(recoloringTrainer.__init__).parameters
    assert sig["hist_resizing"].default == "sampling" and sig["rec_loss"].default == "laplacian"
    assert sig["variance_loss"].default is True and sig["batch_size"].default == 4
    t = inspect.signature(recoloringTrainer.train).parameters
    assert (t["alpha"].default, t["beta"].default, t["gamma"].default) == (32, 1.5, 4)
    assert inspect.signature(recoloringGAN.__init__).parameters["lr"].default == 1e-4
