"""The checkpoint directory of the PyTorch training path, as scripts/train_pytorch.py:149-260 writes and reads it:

    <checkpoint_dir>/<global_step>/model.safetensors              safetensors.torch.save_model(model)
                                  /optimizer.pt                   torch.save(optimizer.state_dict())
                                  /metadata.pt                    {"global_step", "config", "timestamp"}
                                  /assets/<asset_id>/norm_stats.json   shared/normalize.save(...)

written into `tmp_<global_step>` first and renamed, so a reader never sees a half-written step.  Same file names, same
containers, same key names: a directory written here is read by the reference's `load_checkpoint` /
`create_trained_policy` and the other way round (parameter names and dtypes: SURVEY §8a "state-dict key families";
optimizer state: positions in `model.parameters()`, which this module registers in the reference's order -- see
pi0_pytorch._registration_key; `FusedClipAdamW.state_dict(format="torch")` for the fused optimiser).
"""
from __future__ import annotations

import dataclasses
import os
import shutil
import time

import torch

from . import serving


def _unwrap(model):
    return model.module if isinstance(model, torch.nn.parallel.DistributedDataParallel) else model


def get_latest_checkpoint_step(checkpoint_dir):
    """train_pytorch.py:276-283: the largest all-digit directory name, None when there is none."""
    steps = [int(d) for d in os.listdir(checkpoint_dir)
             if d.isdigit() and os.path.isdir(os.path.join(checkpoint_dir, d))] if os.path.isdir(checkpoint_dir) else []
    return max(steps) if steps else None


def save_checkpoint(model, optimizer, global_step: int, checkpoint_dir, *, norm_stats=None, asset_id: str | None = None,
                    config=None, is_main: bool = True, torch_optimizer_format: bool = True) -> str | None:
    """train_pytorch.py:149-189 (the caller decides WHEN to save; :155 is the script's schedule).  Returns the final
    directory.  `config`: a dataclass or dict stored under metadata["config"].  A `FusedClipAdamW` is stored in the
    stock-AdamW layout unless `torch_optimizer_format=False`."""
    if not is_main:
        return None
    final = os.path.join(str(checkpoint_dir), f"{global_step}")
    tmp = os.path.join(str(checkpoint_dir), f"tmp_{global_step}")
    if os.path.exists(tmp):
        shutil.rmtree(tmp)
    os.makedirs(tmp)
    import safetensors.torch

    safetensors.torch.save_model(_unwrap(model), os.path.join(tmp, "model.safetensors"))
    if optimizer is not None:
        from .optim import FusedClipAdamW

        if isinstance(optimizer, FusedClipAdamW):
            sd = optimizer.state_dict(format="torch" if torch_optimizer_format else "flat")
        else:
            sd = optimizer.state_dict()
        torch.save(sd, os.path.join(tmp, "optimizer.pt"))
    cfg = dataclasses.asdict(config) if dataclasses.is_dataclass(config) and not isinstance(config, type) else config
    torch.save({"global_step": global_step, "config": cfg, "timestamp": time.time()}, os.path.join(tmp, "metadata.pt"))
    if norm_stats is not None and asset_id is not None:
        serving.save(os.path.join(tmp, "assets", asset_id), norm_stats)
    if os.path.exists(final):
        shutil.rmtree(final)
    os.rename(tmp, final)
    return final


def load_checkpoint(model, optimizer, checkpoint_dir, device) -> int:
    """train_pytorch.py:192-273: restores the latest step in place and returns its global step."""
    latest = get_latest_checkpoint_step(checkpoint_dir)
    if latest is None:
        raise FileNotFoundError(f"No checkpoints found in {checkpoint_dir}")
    ckpt = os.path.join(str(checkpoint_dir), f"{latest}")
    weights = os.path.join(ckpt, "model.safetensors")
    if not os.path.exists(weights):
        raise FileNotFoundError(f"No model checkpoint found at {ckpt}")
    import safetensors.torch

    safetensors.torch.load_model(_unwrap(model), weights, device=str(device))
    if optimizer is not None:
        path = os.path.join(ckpt, "optimizer.pt")
        if not os.path.exists(path):
            raise FileNotFoundError(f"No optimizer checkpoint found at {ckpt}")
        optimizer.load_state_dict(torch.load(path, map_location=device, weights_only=False))
    meta = torch.load(os.path.join(ckpt, "metadata.pt"), map_location=device, weights_only=False)
    return int(meta.get("global_step", latest))


def load_norm_stats(checkpoint_dir, asset_id: str, step: int | None = None):
    """policy_config.py:57-62: the statistics saved WITH the weights (not the config's assets directory)."""
    if step is None and os.path.exists(os.path.join(str(checkpoint_dir), "model.safetensors")):
        base = str(checkpoint_dir)  # already a step directory
    else:
        step = get_latest_checkpoint_step(checkpoint_dir) if step is None else step
        if step is None:
            raise FileNotFoundError(f"No checkpoints found in {checkpoint_dir}")
        base = os.path.join(str(checkpoint_dir), f"{step}")
    return serving.load_norm_stats(os.path.join(base, "assets"), asset_id)
