"""Golden vectors for the preprocessing row, produced by the REFERENCE'S OWN code:
/root/reference/src/openpi/models_pytorch/preprocessing_pytorch.py is executed here with `jax` and
`openpi.shared.array_typing` stubbed (image_tools.py needs them only for its JAX twin of resize_with_pad).
Runs in the build container only (the GPU box has no /root/reference); the outputs are committed under tests/golden/.

    python tools/make_golden_preprocess.py
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/openpi"


class _Anything:
    """Stands in for jax / array_typing names that image_tools.py touches at import time only."""

    def __getattr__(self, name):
        return _Anything()

    def __getitem__(self, item):
        return self

    def __or__(self, other):
        return self

    def __ror__(self, other):
        return self

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # decorator use
        return _Anything()


def load_reference_preprocessing():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    jax = mod("jax", jit=lambda f=None, **k: f if f is not None else (lambda g: g), image=_Anything(), Array=_Anything())
    mod("jax.numpy", **{n: _Anything() for n in ("uint8", "float32", "round", "pad")})
    jax.numpy = sys.modules["jax.numpy"]
    at = mod("openpi.shared.array_typing", typecheck=lambda f: f, UInt8=_Anything(), Float=_Anything(), Array=_Anything())
    openpi = mod("openpi")
    shared = mod("openpi.shared", array_typing=at)
    openpi.shared = shared

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    shared.image_tools = load("openpi.shared.image_tools", os.path.join(REF, "shared", "image_tools.py"))
    return load("ref_preprocessing_pytorch", os.path.join(REF, "models_pytorch", "preprocessing_pytorch.py"))


class Obs:
    def __init__(self, images, B):
        self.images = images
        self.image_masks = {}
        self.state = torch.zeros(B, 32)
        self.tokenized_prompt = torch.zeros(B, 4, dtype=torch.int64)
        self.tokenized_prompt_mask = torch.ones(B, 4, dtype=torch.bool)
        self.token_ar_mask = None
        self.token_loss_mask = None


KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")


def make_inputs(seed, B, shapes, layout):
    g = torch.Generator().manual_seed(seed)
    images = {}
    for k, (h, w) in zip(KEYS, shapes):
        u8 = torch.randint(0, 256, (B, h, w, 3), generator=g, dtype=torch.uint8)
        x = u8.to(torch.float32) / 255.0 * 2.0 - 1.0
        images[k] = x.permute(0, 3, 1, 2).contiguous() if layout == "nchw" else x
    return images


CASES = [
    # name, seed, batch, per-key (h, w), layout, target resolution, train
    ("train56", 11, 2, [(56, 56)] * 3, "nchw", (56, 56), True),
    ("train56_resize", 12, 2, [(64, 48), (40, 72), (56, 56)], "nhwc", (56, 56), True),
    ("eval56_resize", 13, 2, [(30, 90), (112, 112), (56, 56)], "nchw", (56, 56), False),
    ("train112", 14, 1, [(112, 112)] * 3, "nchw", (112, 112), True),
]


def main():
    sys.path.insert(0, ROOT)
    from oracle import preprocess_oracle as PO

    ref = load_reference_preprocessing()
    for name, seed, B, shapes, layout, res, train in CASES:
        images = make_inputs(seed, B, shapes, layout)
        # the parameters the reference is about to draw: same seed, same call order (oracle.draw_params)
        torch.manual_seed(1000 + seed)
        params = PO.draw_params(KEYS, *res) if train else None
        torch.manual_seed(1000 + seed)
        out = ref.preprocess_observation_pytorch(Obs({k: v.clone() for k, v in images.items()}, B), train=train,
                                                 image_resolution=res)
        mine = PO.preprocess_images(images, KEYS, train=train, params=params, resolution=res)
        for k in KEYS:
            err = float((out.images[k] - mine[k]).abs().max())
            print(f"{name:16s} {k:18s} ref-vs-oracle max abs err {err:.3e}  shape {tuple(out.images[k].shape)}")
        torch.save({"seed": seed, "batch": B, "shapes": shapes, "layout": layout, "resolution": res, "train": train,
                    "params": params, "outputs": {k: out.images[k].to(torch.float32).contiguous() for k in KEYS}},
                   os.path.join(ROOT, "tests", "golden", f"preprocess_{name}.pt"))


if __name__ == "__main__":
    main()
