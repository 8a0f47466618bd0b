"""The reference's OWN training loop (`train_loop` of scripts/train_pytorch.py, executed in place by
tools/reference_train_harness.py) on the CPU around a stand-in for `openpi.models_pytorch.pi0_pytorch`: proves that the
harness supplies exactly what the script imports and that everything else the loop does is the reference's code.  The same
harness runs the loop around the B200 engine in tests/test_zzzz_train_script_gpu.py."""
import logging
import os
import sys
import types

import pytest
import torch

import helpers as H
from kai0_b200 import serving as S
from kai0_b200.model import Observation

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import reference_train_harness as TH  # noqa: E402

KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")


class _StandIn(torch.nn.Module):
    """Same constructor / call surface as PI0Pytorch (pi0_pytorch.py:85,316), trivially small."""

    created: list = []

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.lin = torch.nn.Linear(32, config.action_dim)
        self.calls, self.checkpointing = 0, False
        _StandIn.created.append(self)

    def gradient_checkpointing_enable(self):
        self.checkpointing = True

    def forward(self, observation, actions, noise=None, time=None):
        self.calls += 1
        assert observation.images["base_0_rgb"].device == actions.device and actions.dtype == torch.float32
        return (self.lin(observation.state)[:, None, :].expand_as(actions) - actions) ** 2


def _batches(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        d = {"image": {k: torch.randint(0, 256, (2, 8, 8, 3), dtype=torch.uint8, generator=g) for k in KEYS},
             "image_mask": {k: torch.ones(2, dtype=torch.bool) for k in KEYS}, "state": torch.randn(2, 32, generator=g),
             "tokenized_prompt": torch.zeros(2, 4, dtype=torch.int64), "tokenized_prompt_mask": torch.ones(2, 4, dtype=torch.bool)}
        out.append((Observation.from_dict(d), torch.randn(2, 5, 32, generator=g, dtype=torch.float64)))
    return out


@pytest.mark.skipif(not TH.available(), reason="needs the reference script (/root/reference or the staged copy)")
@pytest.mark.parametrize("prefer_staged", [False, True])
def test_the_references_train_loop_runs_in_place_around_a_stand_in(tmp_path, caplog, prefer_staged):
    if prefer_staged and not all(os.path.isfile(p) for p in TH._STAGED):
        pytest.skip("no staged copy (tools/stage_reference.py)")
    mod = types.ModuleType("stand_in_pi0_pytorch")
    mod.PI0Pytorch = mod.AdvantageEstimator = _StandIn
    _StandIn.created.clear()
    cfg = TH.TrainConfig(checkpoint_dir=tmp_path / "run", model=TH.Pi0Config(action_horizon=5), num_train_steps=4, save_interval=2)
    script = TH.load_script(mod, TH.ListLoader(_batches(2)), prefer_staged)
    stats = {"state": script._normalize.NormStats(mean=torch.zeros(3).numpy(), std=torch.ones(3).numpy())}
    loader = TH.ListLoader(_batches(2), norm_stats=stats, asset_id="agilex")
    with caplog.at_level(logging.INFO):
        script = TH.run(mod, cfg, loader, prefer_staged)
    (model,) = _StandIn.created
    assert model.calls == 4 and model.checkpointing and model.training
    # train_pytorch.py:155: steps on the save interval, and the last one
    assert sorted(os.listdir(tmp_path / "run")) == ["2", "3", "4"]
    assert sorted(os.listdir(tmp_path / "run" / "4")) == ["assets", "metadata.pt", "model.safetensors", "optimizer.pt"]
    assert S.load(tmp_path / "run" / "4" / "assets" / "agilex")["state"].std.tolist() == [1.0, 1.0, 1.0]
    # the script's warm-up / cosine rule (:483-491) with the harness's schedule: peak 1e-3, 2 warm-up steps, decay to 1e-4 by 10
    lrs = [r.getMessage().split("lr=")[1].split()[0] for r in caplog.records if r.getMessage().startswith("step=")]
    assert lrs == ["3.33e-04", "6.67e-04", "1.00e-03", "9.66e-04"]
    meta = torch.load(tmp_path / "run" / "4" / "metadata.pt", weights_only=False)
    assert meta["global_step"] == 4 and meta["config"]["model"]["action_horizon"] == 5
    # resume (:303-318,478-481): the loop restores step 4 and continues to 6
    _StandIn.created.clear()
    cfg2 = TH.TrainConfig(checkpoint_dir=tmp_path / "run", model=TH.Pi0Config(action_horizon=5), num_train_steps=6,
                          save_interval=2, resume=True)
    TH.run(mod, cfg2, loader, prefer_staged)
    (resumed,) = _StandIn.created
    assert resumed.calls == 2 and "6" in os.listdir(tmp_path / "run")
    assert script.get_latest_checkpoint_step(tmp_path / "run") == 6


@pytest.mark.skipif(not TH.available(), reason="needs the reference script (/root/reference or the staged copy)")
def test_fine_tune_and_advantage_estimator_branches_of_the_references_loop(tmp_path):
    """train_pytorch.py:449-460 (`pytorch_weight_path`: safetensors weights loaded before training, strict unless the
    advantage estimator is trained) and :366-368,402-405 (`advantage_estimator=True` builds `AdvantageEstimator`)."""
    import safetensors.torch

    class _Value(_StandIn):
        pass

    mod = types.ModuleType("stand_in_pi0_pytorch")
    mod.PI0Pytorch, mod.AdvantageEstimator = _StandIn, _Value
    pretrained = _StandIn(TH.Pi0Config(action_horizon=5))
    with torch.no_grad():
        pretrained.lin.weight.fill_(0.125)
    os.makedirs(tmp_path / "init")
    safetensors.torch.save_model(pretrained, str(tmp_path / "init" / "model.safetensors"))
    seen = {}
    orig = _StandIn.forward

    def spy(self, observation, actions, noise=None, time=None):
        seen.setdefault("first_weight", self.lin.weight.detach().clone())
        return orig(self, observation, actions, noise, time)

    _StandIn.forward = spy
    try:
        _StandIn.created.clear()
        cfg = TH.TrainConfig(checkpoint_dir=tmp_path / "ft", model=TH.Pi0Config(action_horizon=5), num_train_steps=2,
                             pytorch_weight_path=str(tmp_path / "init"))
        TH.run(mod, cfg, TH.ListLoader(_batches(2)))
        assert type(_StandIn.created[-1]) is _StandIn and torch.all(seen["first_weight"] == 0.125)  # trained FROM the weights
        _StandIn.created.clear()
        cfg = TH.TrainConfig(checkpoint_dir=tmp_path / "adv", model=TH.AdvantageEstimatorConfig(action_horizon=5),
                             num_train_steps=2, advantage_estimator=True)
        TH.run(mod, cfg, TH.ListLoader(_batches(2)))
        assert type(_StandIn.created[-1]) is _Value and _StandIn.created[-1].calls == 2
        # :155 as written: off the save interval only `global_step == num_train_steps - 1` is saved, i.e. step 1 of 2
        assert sorted(os.listdir(tmp_path / "adv")) == ["1"]
    finally:
        _StandIn.forward = orig

