"""CPU: host-side bookkeeping of the Trainer's CUDA-graph path -- lazy scalar read-outs and the decision
WHEN a NaN loss raises NanException (histoGAN/histoGAN.py:1003-1006) -- with stand-ins for the in-flight
read-outs (no GPU work)."""
import pytest
import torch

from histogan_b200.trainer import NanException, Trainer


class FakePending:
    """stands in for trainer._PendingScalars: values already 'on the host'"""

    def __init__(self, **vals):
        self.vals = dict(vals)
        self.vals.setdefault("nan", False)
        self.names = [k for k in self.vals if k != "nan"]
        self.fetched = 0

    def get(self):
        self.fetched += 1
        return self.vals


def _trainer(tmp_path, **kw):
    t = Trainer("t", str(tmp_path / "results"), str(tmp_path / "models"), image_size=32, network_capacity=4,
                batch_size=2, save_every=1000, **kw)
    t.reloaded = []
    t.load = lambda num=-1: t.reloaded.append(num)          # no checkpoint I/O in this test
    t.GAN = type("G", (), {"EMA": lambda self: None, "reset_parameter_averaging": lambda self: None})()
    return t


def test_scalar_readouts_are_plain_attributes_until_a_step_is_in_flight(tmp_path):
    t = _trainer(tmp_path)
    assert (t.d_loss, t.g_loss, t.last_gp_loss) == (0, 0, 0)
    t.d_loss = 3.5                                           # the eager path assigns floats (:930)
    assert t.d_loss == 3.5
    p = FakePending(d_loss=1.0, g_loss=2.0, h_loss=0.25)
    t._pending = p
    assert p.fetched == 0
    assert t.g_loss == 2.0 and t._pending is None            # first access fetches and adopts all of them
    assert (t.d_loss, t.h_loss) == (1.0, 0.25) and t.last_gp_loss == 0
    assert p.fetched == 1


def test_gradient_penalty_readout_survives_the_following_steps(tmp_path):
    t = _trainer(tmp_path)
    t.steps = 2504
    gp_step = FakePending(d_loss=1.0, g_loss=2.0, h_loss=0.2, last_gp_loss=7.0)
    t._pending = gp_step
    t._finish_graphed_step(None, gp_step, False)             # step 2504 queued, nothing fetched
    assert gp_step.fetched == 0 and t.steps == 2505
    nxt = FakePending(d_loss=1.5, g_loss=2.5, h_loss=0.3)
    t._pending = nxt
    t._finish_graphed_step(gp_step, nxt, False)              # looks at the PREVIOUS step only
    assert gp_step.fetched == 1 and nxt.fetched == 0
    assert t.last_gp_loss == 7.0 and t.d_loss == 1.5         # :922 keeps the last penalty; d_loss is the newest


@pytest.mark.parametrize("mode", ["deferred", "immediate"])
def test_when_a_nan_loss_raises(mode, tmp_path):
    t = _trainer(tmp_path, nan_check=mode)
    t.steps = 2505
    bad = FakePending(d_loss=float("nan"), g_loss=1.0, h_loss=0.1, nan=True)
    t._pending = bad
    if mode == "immediate":
        with pytest.raises(NanException):
            t._finish_graphed_step(None, bad, False)
        assert t.reloaded == [2]                             # floor(2505 / 1000)
        return
    t._finish_graphed_step(None, bad, False)                 # the NaN step itself returns
    assert t.steps == 2506 and t.reloaded == []
    ok = FakePending(d_loss=1.0, g_loss=1.0, h_loss=0.1)
    t._pending = ok
    with pytest.raises(NanException):
        t._finish_graphed_step(bad, ok, False)               # ... the next call raises
    assert t.reloaded == [2]                                 # the checkpoint of the step that produced the NaN


def test_checkpoint_and_path_length_steps_look_at_their_own_readouts(tmp_path):
    t = _trainer(tmp_path)
    t.save = lambda num: t.reloaded.append(("saved", num))
    t.evaluate = lambda num: None
    t.steps = 3000                                           # checkpoint + evaluation step
    bad = FakePending(d_loss=float("nan"), g_loss=1.0, h_loss=0.1, nan=True)
    t._pending = bad
    with pytest.raises(NanException):
        t._finish_graphed_step(None, bad, False)
    assert t.reloaded == [3]                                 # nothing was saved
    t.steps = 2528                                           # path-length step: the mean needs the value now
    pl = FakePending(d_loss=1.0, g_loss=1.0, h_loss=0.1, avg_pl=4.0)
    t._pending = pl
    t.pl_mean = 0
    t._finish_graphed_step(None, pl, True)
    assert pl.fetched == 1 and t.pl_mean == pytest.approx(0.04)      # EMA(0.99) from 0 (:996-997)
