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


class _HashTokenizer:
    """Ragged prompts inside the tiny configuration's 24 slots (as tests/test_zzz_serving_gpu.py)."""

    def __init__(self, max_len, vocab):
        self.max_len, self.vocab = max_len, vocab

    def tokenize(self, prompt, state=None):
        import numpy as np

        bins = np.digitize(state, bins=np.linspace(-1, 1, 257)[:-1]) - 1
        n = 7 + (sum(map(ord, prompt)) % (self.max_len - 9))
        ids = [2] + [int((ord(prompt[i % len(prompt)]) * 7 + int(bins[i % len(bins)]) + 3 * i) % (self.vocab - 1)) + 1
                     for i in range(n - 1)]
        pad = self.max_len - n
        return np.asarray(ids + [0] * pad), np.asarray([True] * n + [False] * pad)


def test_the_references_own_create_trained_policy_serves_from_the_engine(tmp_path):
    """north_star: "scripts/serve_policy.py call it unchanged".  The reference's OWN `create_trained_policy`
    (policies/policy_config.py, executed in place by tools/reference_serve_harness.py) loads a checkpoint directory into
    `kai0_b200.pi0_pytorch.PI0Pytorch` through the reference's `load_pytorch`, and the reference's OWN `Policy.infer`
    (policies/policy.py: per-leaf copies, `Observation.from_dict`, `sample_actions`) serves Agilex requests from it; the
    replies must agree with `kai0_b200.serving`'s policy on the same checkpoint (same engine, same inputs up to the image
    format -- fp32 NCHW from the reference's from_dict vs uint8 NHWC -- which the engine converts identically)."""
    import numpy as np

    import reference_serve_harness as RSH
    import kai0_b200.pi0_pytorch as b200
    from kai0_b200 import checkpoint as CK
    from kai0_b200 import serving as S

    if not RSH.available():
        pytest.skip("needs the reference's serving files (checkout or staged copy)")
    oc = O.tiny_config()
    trained, _ = H.build_pair(oc, seed=21, device=None)
    g = np.random.default_rng(11)
    stats = {}
    for key in ("state", "actions"):
        mean = g.normal(0, 0.3, 32)
        q01, q99 = mean - g.uniform(1.0, 2.0, 32), mean + g.uniform(1.0, 2.0, 32)
        for a in (mean, q01, q99):
            a[14:] = 0.0
        stats[key] = S.NormStats(mean=mean, std=np.ones(32), q01=q01, q99=q99)
    step_dir = CK.save_checkpoint(trained, None, 100, tmp_path, norm_stats=stats, asset_id="agilex")
    tok = _HashTokenizer(oc.max_token_len, oc.vocab_size)
    fields = dataclass_fields(_model_config(oc))
    ref = RSH.reference_policy(b200, fields, step_dir, asset_id="agilex", tokenizer=tok, default_prompt="fold the cloth",
                               pytorch_device="cuda", image_size=oc.image_size)
    assert type(ref).__module__ == "openpi.policies.policy" and isinstance(ref._model, b200.PI0Pytorch)
    mine = S.create_trained_policy(b200.PI0Pytorch(_model_config(oc)), step_dir, asset_id="agilex", tokenizer=tok,
                                   default_prompt="fold the cloth", pytorch_device="cuda")
    for (n, a), (_, b) in zip(ref._model.named_parameters(), mine._model.named_parameters()):
        assert torch.equal(a, b) and torch.equal(a.cpu(), dict(trained.named_parameters())[n]), n
    cams = ("top_head", "hand_left", "hand_right")
    prompts = ["fold the cloth", "hang the shirt on the hanger"]
    for i in range(2):
        req = {"images": {c: g.integers(0, 256, (3, 90, 120), dtype=np.uint8) for c in cams},
               "state": g.uniform(-1, 1, 14).astype(np.float32), "prompt": prompts[i]}
        nz = g.normal(size=(oc.action_horizon, oc.action_dim)).astype(np.float32)
        a = ref.infer({**req, "images": dict(req["images"])}, noise=nz)
        b = mine.infer({**req, "images": dict(req["images"])}, noise=nz)
        assert a["actions"].shape == b["actions"].shape == (oc.action_horizon, 14) and np.isfinite(a["actions"]).all()
        err = H.rel_err(torch.from_numpy(np.asarray(b["actions"])), torch.from_numpy(np.asarray(a["actions"])))
        print(f"reference Policy on the engine vs kai0_b200.serving.Policy, request {i}: rel {err:.2e}")
        assert err < 2e-3, (i, err)


def dataclass_fields(cfg):
    import dataclasses

    return {f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg)}

