"""Shared pieces of the reference pin (test infrastructure): the configuration, the seeded weights and the seeded inputs
on which the REFERENCE'S OWN PyTorch model was run on the CPU of the build container (tools/make_golden_reference.py),
so that the oracle — and, through the committed outputs, the B200 engine — can be compared with the reference itself.

Configuration "pin": everything the reference hard-codes is kept (SigLIP-So400m geometry 1152 / 16 heads / mlp 4304 /
patch 14 / 224 px, projector to 2048, vocabulary 257152, action dim 32, horizon 50, GQA 8:1 with head_dim 256); only
the depths and MLP widths are small so a CPU pass takes seconds: PaliGemma 2048 wide x 2 layers (mlp 512), action
expert 256 wide x 2 layers (mlp 512), SigLIP truncated to 2 layers, prompt length 48.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PG = (2048, 2, 512, 8, 1, 256)   # width, depth, mlp_dim, heads, kv heads, head_dim
EX = (256, 2, 512, 8, 1, 256)
VIT_LAYERS = 2
MAX_TOKEN_LEN = 48
BATCH = 2
WEIGHT_SEED = 20250923
KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
_FP32_CACHE = None


def oracle_config():
    from oracle import pi05_oracle as O

    return O.OracleConfig(paligemma=O.GemmaCfg(*PG), expert=O.GemmaCfg(*EX), vit_depth=VIT_LAYERS,
                          max_token_len=MAX_TOKEN_LEN)


def pin_weights(specs: dict, dtype_map: bool = True) -> dict:
    """Seeded weights for `specs` = {name: (shape, dtype)} in sorted-name order (independent of dict order).  Every
    tensor is drawn in fp32 and cast; norm weights and the (reference-zero-initialised) adaRMS dense layers get non-zero
    values so those paths are exercised.  dtype_map=False keeps everything fp32 (the reference's precision="float32")."""
    global _FP32_CACHE
    key = tuple((n, tuple(specs[n][0])) for n in sorted(specs))
    if _FP32_CACHE is not None and _FP32_CACHE[0] == key:  # the fp32 draw is shared by both precisions of one process
        return {n: (v.to(specs[n][1]) if dtype_map else v.clone()) for n, v in _FP32_CACHE[1].items()}
    g = torch.Generator().manual_seed(WEIGHT_SEED)
    out = {}
    for name in sorted(specs):
        shape, dt = specs[name]
        x = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith("layer_norm1.weight") or name.endswith("layer_norm2.weight") or name.endswith("post_layernorm.weight"):
            x = 1.0 + 0.1 * x
        elif name.endswith("layernorm.weight") or name.endswith("model.norm.weight") or name.endswith("language_model.norm.weight"):
            x = 0.1 * x  # GemmaRMSNorm multiplies by (1 + w)
        elif name.startswith(("action_", "time_mlp")):
            x = x * (0.05 if name.endswith("weight") else 0.02)
        else:
            x = 0.02 * x
        out[name] = x
    _FP32_CACHE = (key, out)
    return {n: (v.to(specs[n][1]) if dtype_map else v.clone()) for n, v in out.items()}


def pin_inputs():
    """B = 2 synthetic observations (ragged prompts, one masked camera), actions, injected noise and time."""
    from oracle import pi05_oracle as O

    oc = oracle_config()
    b = O.synthetic_batch(oc, BATCH, seed=77, ragged=True)
    b["img_masks"][1][0] = False
    return b
