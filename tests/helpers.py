"""Shared test helpers: build the engine-backed module and the oracle from ONE config + ONE seeded weight set."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pi05_oracle as O  # noqa: E402  (tests are allowed to import the oracle)


def mid_config(**kw):
    """Real head geometry (head_dim 256, GQA 8:1, ViT heads of 72) at small width/depth."""
    base = dict(
        paligemma=O.GemmaCfg(512, 2, 1024, 8, 1, 256),
        expert=O.GemmaCfg(256, 2, 512, 8, 1, 256),
        vit_width=288,
        vit_depth=2,
        vit_mlp_dim=560,
        vit_heads=4,
        vit_patch=14,
        image_size=112,
        vocab_size=1024,
        action_dim=32,
        action_horizon=50,
        max_token_len=40,
        num_images=3,
    )
    base.update(kw)
    return O.OracleConfig(**base)


def engine_config(oc: O.OracleConfig):
    from kai0_b200.pi0_pytorch import Pi05EngineConfig

    return Pi05EngineConfig(
        paligemma_variant=oc.paligemma,
        action_expert_variant=oc.expert,
        action_dim=oc.action_dim,
        action_horizon=oc.action_horizon,
        max_token_len=oc.max_token_len,
        vit_width=oc.vit_width,
        vit_depth=oc.vit_depth,
        vit_mlp_dim=oc.vit_mlp_dim,
        vit_heads=oc.vit_heads,
        vit_patch=oc.vit_patch,
        image_size=oc.image_size,
        vocab_size=oc.vocab_size,
        num_images=oc.num_images,
    )


def build_pair(oc: O.OracleConfig, seed: int = 0, device="cuda", cls=None):
    """(engine module on `device`, oracle param dict on CPU) sharing the same weights."""
    from kai0_b200.pi0_pytorch import PI0Pytorch

    params = O.init_params(oc, seed)
    model = (cls or PI0Pytorch)(engine_config(oc))
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all("lm_head" in m for m in missing), missing
    if device is not None:
        model = model.to(device)
    # the oracle's forward_loss restates the network on un-augmented images; the train-time augmentation
    # (on by default like the reference, pi0_pytorch.py:318) has its own parity tests (tests/test_preprocess_gpu.py)
    model.augment = False
    return model, params


class Obs:
    """Duck-typed observation (what PI0Pytorch reads: preprocessing_pytorch.py:165-173)."""

    def __init__(self, batch, device=None):
        keys = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
        mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
        self.images = {k: mv(batch["images"][i]) for i, k in enumerate(keys)}
        self.image_masks = {k: mv(batch["img_masks"][i]) for i, k in enumerate(keys)}
        self.state = mv(torch.zeros(batch["tokens"].shape[0], 32))
        self.tokenized_prompt = mv(batch["tokens"])
        self.tokenized_prompt_mask = mv(batch["token_mask"])
        self.token_ar_mask = None
        self.token_loss_mask = None


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_err(a, b) -> float:
    return float((a.detach().float().cpu() - b.detach().float().cpu()).abs().max())
