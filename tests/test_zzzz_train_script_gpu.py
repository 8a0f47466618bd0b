"""north_star: "scripts/train_pytorch.py ... call it unchanged".  The reference's OWN `train_loop` (scripts/train_pytorch.py:
295-640, executed in place through tools/reference_train_harness.py; on the GPU box from the byte-for-byte staged copy under
baseline/_ref) runs around `kai0_b200.pi0_pytorch` as `openpi.models_pytorch.pi0_pytorch`: the script seeds, builds
`PI0Pytorch(config).to(device)`, asks for gradient checkpointing, creates `torch.optim.AdamW(model.parameters())`, applies its
learning-rate rule, calls `model(observation, actions)`, `.mean().backward()`, `clip_grad_norm_`, `optim.step()`,
`zero_grad(set_to_none=True)` and `save_checkpoint` -- all its code; only jax / wandb / the config and data-loader packages
it imports at the top are supplied by the harness (listed in its header).  Runs last: it is the one GPU test written after
the round's GPU budget was spent, so the driver's round-end run is its first execution on hardware."""
import os
import sys
import types

import pytest
import torch

import helpers as H
from oracle import pi05_oracle as O

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import reference_train_harness as TH  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
STEPS = 3


def _engine_module():
    """A module object exporting the drop-in classes, recording every instance the script creates and every loss."""
    import kai0_b200.pi0_pytorch as b200

    created = []

    class Recorded(b200.PI0Pytorch):
        def __init__(self, config):
            super().__init__(config)
            self.losses = []
            created.append(self)

        def forward(self, observation, actions, noise=None, time=None):
            out = super().forward(observation, actions, noise, time)
            self.losses.append(float(out.detach().float().mean()))
            return out

    mod = types.ModuleType("kai0_b200_as_openpi_pi0_pytorch")
    mod.PI0Pytorch, mod.AdvantageEstimator = Recorded, b200.AdvantageEstimator
    return mod, created, b200.PI0Pytorch


def _model_config(oc):
    return TH.Pi0Config(dtype="bfloat16", action_dim=oc.action_dim, action_horizon=oc.action_horizon,
                        max_token_len=oc.max_token_len, paligemma_variant=oc.paligemma, action_expert_variant=oc.expert,
                        pi05=True, vit_width=oc.vit_width, vit_depth=oc.vit_depth, vit_mlp_dim=oc.vit_mlp_dim,
                        vit_heads=oc.vit_heads, vit_patch=oc.vit_patch, image_size=oc.image_size, vocab_size=oc.vocab_size,
                        num_images=oc.num_images)


def _loader(oc, n):
    """(Observation, actions) pairs on the HOST, as the reference's data loader yields them (fp32 NCHW images out of
    `Observation.from_dict`, data_loader.py:605-607); the script moves them to the device (:531-533)."""
    from kai0_b200.model import Observation

    out = []
    for i in range(n):
        b = O.synthetic_batch(oc, 2, seed=100 + i, ragged=True)
        obs = Observation(images={k: b["images"][j] for j, k in enumerate(KEYS)},
                          image_masks={k: b["img_masks"][j] for j, k in enumerate(KEYS)},
                          state=torch.zeros(2, 32), tokenized_prompt=b["tokens"], tokenized_prompt_mask=b["token_mask"])
        out.append((obs, b["actions"].to(torch.float64)))  # the script casts actions to float32 itself (:532)
    return TH.ListLoader(out)


@pytest.mark.skipif(not TH.available(), reason="needs the reference's train_pytorch.py (checkout or staged copy)")
def test_the_references_own_train_loop_runs_on_the_engine(tmp_path):
    oc = O.tiny_config()
    mod, created, plain_cls = _engine_module()
    cfg = TH.TrainConfig(checkpoint_dir=tmp_path / "run", model=_model_config(oc), num_train_steps=STEPS, save_interval=1,
                         batch_size=2)
    script = TH.run(mod, cfg, _loader(oc, 2))
    (model,) = created
    assert len(model.losses) == STEPS and all(l == l and abs(l) < 1e6 for l in model.losses), model.losses
    assert model.is_gradient_checkpointing_enabled() and model.training
    print(f"reference train_loop on the engine: losses {['%.4f' % l for l in model.losses]}")
    # train_pytorch.py:155: with save_interval 1 every step after the first is on the schedule
    assert sorted(os.listdir(tmp_path / "run")) == [str(s) for s in range(1, STEPS + 1)]
    dev = next(model.parameters()).device
    assert dev.type == "cuda"

    def restore(step):
        m = plain_cls(_model_config(oc)).to(dev)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)
        only = tmp_path / f"only_{step}"
        os.makedirs(only)
        os.symlink(tmp_path / "run" / str(step), only / str(step))
        assert script.load_checkpoint(m, opt, only, dev) == step  # the reference's own loader
        return m, opt

    last, last_opt = restore(STEPS)
    first, _ = restore(1)
    live = dict(model.named_parameters())
    moved = 0
    for n, p in last.named_parameters():
        assert torch.isfinite(p.float()).all(), n
        assert torch.equal(p, live[n]), n  # the last checkpoint IS the model the loop ended with
        moved += int(not torch.equal(p, dict(first.named_parameters())[n]))
    assert moved > 50  # training moved (nearly) every tensor between step 1 and step 3
    state = last_opt.state_dict()["state"]
    assert len(state) > 50 and all(float(s["step"]) == STEPS for s in state.values())
