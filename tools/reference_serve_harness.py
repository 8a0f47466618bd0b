"""Build a policy with the REFERENCE'S OWN `create_trained_policy` (src/openpi/policies/policy_config.py:16-94, executed in
place) around a drop-in for `openpi.models_pytorch.pi0_pytorch` (test infrastructure; needs the reference files, from the
checkout or the staged copy under baseline/_ref).

north_star asks that "scripts/serve_policy.py call it unchanged".  serve_policy.py is `create_trained_policy(config, dir)` +
the websocket server; the function, the `Policy` it returns (policies/policy.py), `Observation.from_dict` and
`BaseModelConfig.load_pytorch` (models/model.py:276-280), the transforms, the Agilex data config and the norm-stats loader
all run here as the reference wrote them.  What the harness supplies is the TrainConfig object the function reads (the real
one lives in the jax/tyro-typed config registry): `.model` (the model fields + the reference's own `load_pytorch` bound to
it), `.data.create(...)` returning the Agilex data config assembled from the reference's own transform classes exactly as
training/config.py:420-452,129-141 does, `.assets_dirs`, `.policy_metadata`.
"""
from __future__ import annotations

import os
import sys
import types

_TOOLS = os.path.dirname(os.path.abspath(__file__))
if _TOOLS not in sys.path:
    sys.path.insert(0, _TOOLS)

import reference_serving_loader as RSL  # noqa: E402


def available() -> bool:
    return RSL.available() and os.path.isfile(os.path.join(RSL.SRC, "policies", "policy_config.py"))


def reference_policy(pi0_module, model_fields: dict, checkpoint_dir, *, asset_id: str, tokenizer, default_prompt=None,
                     sample_kwargs=None, pytorch_device=None, image_size: int = 224, metadata=None):
    """`create_trained_policy(train_config, checkpoint_dir, ...)` of the reference with `pi0_module` as
    `openpi.models_pytorch.pi0_pytorch`.  `model_fields`: what `PI0Pytorch(config=...)` reads.  `tokenizer`: an object with
    `tokenize(prompt, state)` (the reference's `PaligemmaTokenizer` needs the bucket-hosted SentencePiece model)."""
    pc, policy_mod, model_mod = RSL.load_policy_config()
    R = RSL.load()
    T = R.transforms
    model_mod.pi0_pytorch = pi0_module  # `from openpi.models_pytorch import pi0_pytorch` of models/model.py:20

    class ModelConfig(types.SimpleNamespace):
        load_pytorch = model_mod.BaseModelConfig.load_pytorch  # the reference's own method (models/model.py:276-280)

    mcfg = ModelConfig(**model_fields)
    mask = T.make_bool_mask(6, -1, 6, -1)  # training/config.py:436-441
    data_transforms = T.Group(
        inputs=[R.agilex_policy.AgilexInputs(action_dim=mcfg.action_dim, model_type=R.ModelType.PI05)],
        outputs=[R.agilex_policy.AgilexOutputs()],
    ).push(inputs=[T.DeltaActions(mask)], outputs=[T.AbsoluteActions(mask)])
    model_transforms = T.Group(inputs=[  # training/config.py:129-141 (ResizeImages(224, 224) there)
        T.InjectDefaultPrompt(default_prompt), T.ResizeImages(image_size, image_size),
        T.TokenizePrompt(tokenizer, discrete_state_input=True), T.PadStatesAndActions(mcfg.action_dim)])
    data_config = types.SimpleNamespace(asset_id=asset_id, use_quantile_norm=True, data_transforms=data_transforms,
                                        model_transforms=model_transforms)
    train_config = types.SimpleNamespace(model=mcfg, data=types.SimpleNamespace(create=lambda assets_dirs, model: data_config),
                                         assets_dirs=None, policy_metadata=metadata)
    import pathlib

    return pc.create_trained_policy(train_config, pathlib.Path(str(checkpoint_dir)), default_prompt=default_prompt,
                                    sample_kwargs=sample_kwargs, pytorch_device=pytorch_device)
