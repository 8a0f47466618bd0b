"""Run the REFERENCE'S OWN training loop -- `train_loop(config)` of scripts/train_pytorch.py:295-640, executed in place,
unmodified -- around a drop-in for `openpi.models_pytorch.pi0_pytorch` (test infrastructure; needs /root/reference).

north_star asks that "scripts/train_pytorch.py ... call it unchanged".  The script cannot start here as a program (it
imports jax, wandb and the jax-typed config / data-loader packages at the top), so this harness supplies exactly those
imports and nothing else:

  jax                              `jax.tree.map(fn, observation)` (the script's only use, :531): maps over dict values and
                                   dataclass fields, leaves None alone
  wandb                            init / log / finish / Image that do nothing (the run is configured with wandb disabled)
  openpi.models.pi0_config         `Pi0Config` / `AdvantageEstimatorConfig` dataclasses (the `isinstance` checks at :366,402)
  openpi.training.config           a namespace (type annotations only)
  openpi.training.data_loader      `create_data_loader(config, framework=..., shuffle=..., skip_norm_stats=...)` returning the
                                   caller's loader (an iterable of `(observation, actions)` with `.data_config()`)
  openpi.shared.normalize          the reference's real module (norm stats saved with the checkpoint, :173-176)
  openpi.models_pytorch.pi0_pytorch   THE MODULE UNDER TEST (kai0_b200.pi0_pytorch on a B200; a CPU stand-in in the CPU test)

Everything the loop does -- seeding, checkpoint directory handling, model construction and `.to(device)`, gradient
checkpointing request, AdamW over `model.parameters()`, the warm-up / cosine learning-rate rule, forward, `.mean().backward()`,
`clip_grad_norm_`, `optim.step()`, `zero_grad(set_to_none=True)`, logging, `save_checkpoint` -- is the reference's code.
"""
from __future__ import annotations

import dataclasses
import importlib.util
import os
import pathlib
import sys
import types

_TOOLS = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_TOOLS)
if _TOOLS not in sys.path:
    sys.path.insert(0, _TOOLS)
# (script, normalize.py): the read-only checkout of the build container, else the byte-for-byte staged copy under
# baseline/_ref (tools/stage_reference.py; travels to the GPU box, where /root/reference does not exist)
_CHECKOUT = ("/root/reference/scripts/train_pytorch.py", "/root/reference/src/openpi/shared/normalize.py")
_STAGED = (os.path.join(_ROOT, "baseline", "_ref", "scripts", "train_pytorch.py"),
           os.path.join(_ROOT, "baseline", "_ref", "openpi", "shared", "normalize.py"))


def paths(prefer_staged: bool = False):
    order = (_STAGED, _CHECKOUT) if prefer_staged else (_CHECKOUT, _STAGED)
    if os.environ.get("KAI0_REFERENCE_STAGED_ONLY"):
        order = (_STAGED,)
    for cand in order:
        if all(os.path.isfile(p) for p in cand):
            return cand
    return None


def available() -> bool:
    return paths() is not None


@dataclasses.dataclass
class Pi0Config:
    """The fields `PI0Pytorch.__init__` reads (pi0_pytorch.py:85-98) plus the geometry the reference hard-codes elsewhere
    (SigLIP So400m/14, 224 px, PaliGemma vocabulary), spelled out so that tests can shrink it."""

    dtype: str = "bfloat16"
    action_dim: int = 32
    action_horizon: int = 50
    max_token_len: int = 200
    paligemma_variant: object = "gemma_2b"
    action_expert_variant: object = "gemma_300m"
    pi05: bool = True
    vit_width: int = 1152
    vit_depth: int = 27
    vit_mlp_dim: int = 4304
    vit_heads: int = 16
    vit_patch: int = 14
    image_size: int = 224
    vocab_size: int = 257152
    num_images: int = 3


class AdvantageEstimatorConfig(Pi0Config):
    pass


@dataclasses.dataclass
class LRSchedule:  # training/optimizer.py CosineDecaySchedule (the four numbers train_pytorch.py:463-466 reads)
    warmup_steps: int = 2
    peak_lr: float = 1e-3
    decay_steps: int = 10
    decay_lr: float = 1e-4


@dataclasses.dataclass
class Optimizer:  # training/optimizer.py AdamW (train_pytorch.py:469-475,557)
    b1: float = 0.9
    b2: float = 0.95
    eps: float = 1e-8
    weight_decay: float = 1e-10
    clip_gradient_norm: float = 1.0


@dataclasses.dataclass
class TrainConfig:
    """The TrainConfig fields `train_loop` touches (training/config.py)."""

    checkpoint_dir: pathlib.Path
    model: Pi0Config
    seed: int = 42
    batch_size: int = 2
    num_train_steps: int = 3
    log_interval: int = 1
    save_interval: int = 1000
    resume: bool = False
    overwrite: bool = False
    wandb_enabled: bool = False
    skip_norm_stats: bool = True
    advantage_estimator: bool = False
    pytorch_training_precision: str = "bfloat16"
    pytorch_weight_path: str | None = None
    project_name: str = "kai0_b200"
    exp_name: str = "harness"
    lr_schedule: LRSchedule = dataclasses.field(default_factory=LRSchedule)
    optimizer: Optimizer = dataclasses.field(default_factory=Optimizer)


class ListLoader:
    """What `create_data_loader(..., framework="pytorch")` hands the loop: iterable of (observation, actions), `len`, and
    `data_config()` (training/data_loader.py:566-612)."""

    def __init__(self, batches, norm_stats=None, asset_id=None):
        self._batches = list(batches)
        self._dc = types.SimpleNamespace(norm_stats=norm_stats, asset_id=asset_id)

    def __iter__(self):
        return iter(self._batches)

    def __len__(self):
        return len(self._batches)

    def data_config(self):
        return self._dc


def _tree_map(fn, tree):
    if tree is None:
        return None
    if isinstance(tree, dict):
        return {k: _tree_map(fn, v) for k, v in tree.items()}
    if dataclasses.is_dataclass(tree) and not isinstance(tree, type):
        return type(tree)(**{f.name: _tree_map(fn, getattr(tree, f.name)) for f in dataclasses.fields(tree)})
    return fn(tree)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


_normalize_cache: dict = {}


def _load_normalize(path: str):
    """The reference's shared/normalize.py executed in place (numpydantic stand-in: tools/reference_serving_loader.py)."""
    import reference_serving_loader as RSL

    if path in _normalize_cache:  # one NormStats class per file, whatever number of times the script is loaded
        return _normalize_cache[path]
    saved = sys.modules.get("numpydantic")
    sys.modules["numpydantic"] = RSL.numpydantic_stub()
    try:
        spec = importlib.util.spec_from_file_location("_kai0_reference_normalize", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("numpydantic", None)
        else:
            sys.modules["numpydantic"] = saved
    _normalize_cache[path] = mod
    return mod


def load_script(pi0_module, loader: ListLoader, prefer_staged: bool = False):
    """scripts/train_pytorch.py executed in place as a fresh module whose imports resolve as described in the header."""
    script_path, normalize_path = paths(prefer_staged)
    normalize = _load_normalize(normalize_path)
    wandb = _Stub("wandb")
    wandb.run = types.SimpleNamespace(id="offline")
    jax = types.ModuleType("jax")
    jax.tree = types.SimpleNamespace(map=_tree_map)
    cfg_mod = types.ModuleType("openpi.models.pi0_config")
    cfg_mod.Pi0Config, cfg_mod.AdvantageEstimatorConfig = Pi0Config, AdvantageEstimatorConfig
    data_mod = types.ModuleType("openpi.training.data_loader")
    data_mod.create_data_loader = lambda config, framework="pytorch", shuffle=True, skip_norm_stats=False: loader
    tr_cfg = _Stub("openpi.training.config")
    tr_cfg.TrainConfig = TrainConfig
    training = types.ModuleType("openpi.training")
    training.__path__ = []
    training.config, training.data_loader = tr_cfg, data_mod
    mp = types.ModuleType("openpi.models_pytorch")
    mp.__path__ = []
    mp.pi0_pytorch = pi0_module
    models = types.ModuleType("openpi.models")
    models.__path__ = []
    models.pi0_config = cfg_mod
    shared = types.ModuleType("openpi.shared")
    shared.__path__ = []
    shared.normalize = normalize
    openpi = types.ModuleType("openpi")
    openpi.__path__ = []
    openpi.models, openpi.models_pytorch, openpi.training, openpi.shared = models, mp, training, shared
    mods = {"jax": jax, "wandb": wandb, "openpi": openpi, "openpi.models": models, "openpi.models.pi0_config": cfg_mod,
            "openpi.models_pytorch": mp, "openpi.models_pytorch.pi0_pytorch": pi0_module, "openpi.training": training,
            "openpi.training.config": tr_cfg, "openpi.training.data_loader": data_mod, "openpi.shared": shared,
            "openpi.shared.normalize": normalize}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("_kai0_reference_train_pytorch_harness", script_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def run(pi0_module, config: TrainConfig, loader: ListLoader, prefer_staged: bool = False):
    """`train_loop(config)` of the reference's script with `pi0_module` as `openpi.models_pytorch.pi0_pytorch`.  Returns
    the script module (for its `load_checkpoint`, ...)."""
    script = load_script(pi0_module, loader, prefer_staged)
    script.train_loop(config)
    return script
