"""Run the REFERENCE'S OWN PyTorch model on the CPU of the build container (test infrastructure; needs /root/reference).

The reference (`src/openpi/models_pytorch/pi0_pytorch.py`) expects transformers 4.53.2 with three modeling files
overwritten by `transformers_replace/`, and imports the JAX-side `openpi.models.gemma` for its size table.  Here:
  * the three patched modeling files (+ the patched GemmaConfig) are executed from where they lie under /root/reference
    and installed in `sys.modules` UNDER THE NAMES OF THE INSTALLED transformers modules they replace, so that
    `from transformers import GemmaForCausalLM, PaliGemmaForConditionalGeneration` resolves to the patched classes
    (the installed transformers 5.5 lacks one name they import, `utils.LossKwargs`: stubbed);
  * `openpi.models.gemma.get_config` is a stub returning the size record the caller registered (no jax);
  * `openpi.shared.image_tools` is loaded with `jax` stubbed (its torch half is what the PyTorch path uses);
  * `transformers.models.siglip.check` (version guard) is replaced by a stub that says "installed";
  * `torch.compile` is a no-op while the model is constructed (the reference wraps sample_actions in it).
Nothing of the reference is copied: the files are executed in place.  Used by tools/make_golden_reference.py and
tests/test_reference_cpu.py to pin oracle/pi05_oracle.py to the reference itself.
"""
from __future__ import annotations

import dataclasses
import importlib
import importlib.util
import os
import sys
import types
import typing

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Where the reference's `openpi` package lies: the read-only checkout of the build container, else the offline install
# of the UNMODIFIED package under baseline/_ref (git-ignored; made by tools/stage_reference.py, which is what
# `pip install --target baseline/_ref /root/reference` would have produced had the `hatchling` build backend been in
# the wheelhouse) -- that copy travels to the GPU box with the snapshot, /root/reference does not.
_CANDIDATES = ("/root/reference/src/openpi", os.path.join(_ROOT, "baseline", "_ref", "openpi"))
REF = next((c for c in _CANDIDATES if os.path.isdir(os.path.join(c, "models_pytorch"))), _CANDIDATES[0])
TR = os.path.join(REF, "models_pytorch", "transformers_replace", "models")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models_pytorch"))


@dataclasses.dataclass
class SizeRecord:  # what openpi.models.gemma.get_config returns (src/openpi/models/gemma.py:42-56)
    width: int
    depth: int
    mlp_dim: int
    num_heads: int
    num_kv_heads: int
    head_dim: int


_VARIANTS: dict = {
    "gemma_2b": SizeRecord(2048, 18, 16384, 8, 1, 256),
    "gemma_300m": SizeRecord(1024, 18, 4096, 8, 1, 256),
}
_loaded = None


def register_variant(name: str, rec: SizeRecord) -> None:
    _VARIANTS[name] = rec


def _exec_as(name: str, path: str, package: str | None = None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _Anything:
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
            return a[0]
        return _Anything()


def load(vision_layers: int | None = None):
    """Returns the reference's `pi0_pytorch` module (cached).  `vision_layers` truncates the SigLIP tower the reference
    builds from the HF default PaliGemma config (27 layers) so CPU runs stay short; everything else is the reference's."""
    global _loaded
    if _loaded is not None:
        return _loaded
    import transformers
    import transformers.utils as tu

    if not hasattr(tu, "LossKwargs"):
        class LossKwargs(typing.TypedDict, total=False):
            num_items_in_batch: int

        tu.LossKwargs = LossKwargs
    import transformers.cache_utils as cu

    if not hasattr(cu.DynamicCache, "__getitem__"):
        # 4.53.2: cache[layer_idx] -> (keys, values); the patched attention reads the prefix cache that way
        # (modeling_gemma.py:308-310).  5.5 keeps the same tensors in cache.layers[i].keys / .values.
        cu.DynamicCache.__getitem__ = lambda self, i: (self.layers[i].keys, self.layers[i].values)
    if not hasattr(cu, "HybridCache"):  # only named in an isinstance() branch the pi0.5 path never takes
        cu.HybridCache = type("HybridCache", (), {})
    import transformers.modeling_rope_utils as ru

    if "default" not in ru.ROPE_INIT_FUNCTIONS:
        # transformers 4.53.2 `_compute_default_rope_parameters` (the pinned dependency's published algorithm; 5.5
        # moved it): inv_freq_i = base^(-2i/dim), attention factor 1
        def _default_rope(config, device=None, seq_len=None, **_):
            base = getattr(config, "rope_theta", None) or 10000.0
            dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
            inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / dim))
            return inv, 1.0

        ru.ROPE_INIT_FUNCTIONS["default"] = _default_rope
    # patched HF files, under the installed module names
    import transformers.models.gemma  # noqa: F401  (packages must exist before their submodules are replaced)
    import transformers.models.paligemma  # noqa: F401
    import transformers.models.siglip  # noqa: F401

    g = _exec_as("transformers.models.gemma.modeling_gemma", os.path.join(TR, "gemma", "modeling_gemma.py"),
                 "transformers.models.gemma")
    s = _exec_as("transformers.models.siglip.modeling_siglip", os.path.join(TR, "siglip", "modeling_siglip.py"),
                 "transformers.models.siglip")
    pg = _exec_as("transformers.models.paligemma.modeling_paligemma",
                  os.path.join(TR, "paligemma", "modeling_paligemma.py"), "transformers.models.paligemma")
    sys.modules["transformers.models.gemma"].modeling_gemma = g
    sys.modules["transformers.models.siglip"].modeling_siglip = s
    sys.modules["transformers.models.paligemma"].modeling_paligemma = pg
    chk = types.ModuleType("transformers.models.siglip.check")
    chk.check_whether_transformers_replace_is_installed_correctly = lambda: True
    sys.modules["transformers.models.siglip.check"] = chk
    sys.modules["transformers.models.siglip"].check = chk
    # transformers 5.5 wants the tie map as {target: source}; 4.53.2 (and the patched files) declare a list of targets
    g.GemmaForCausalLM._tied_weights_keys = {"lm_head.weight": "model.embed_tokens.weight"}
    pg.PaliGemmaForConditionalGeneration._tied_weights_keys = {"lm_head.weight": "model.language_model.embed_tokens.weight"}
    for cls in (g.GemmaForSequenceClassification, g.GemmaForTokenClassification):
        if isinstance(getattr(cls, "_tied_weights_keys", None), list):
            cls._tied_weights_keys = {}
    # the lazily exported top-level names must be the patched classes
    transformers.GemmaForCausalLM = g.GemmaForCausalLM
    transformers.PaliGemmaForConditionalGeneration = pg.PaliGemmaForConditionalGeneration
    from transformers.models.auto import CONFIG_MAPPING, MODEL_MAPPING  # noqa: F401

    gcfg = CONFIG_MAPPING["gemma"]  # the installed GemmaConfig; the patched modeling file reads the adaRMS fields with getattr
    pcfg_cls = CONFIG_MAPPING["paligemma"]

    class _CfgMap(dict):
        """CONFIG_MAPPING as gemma_pytorch.py uses it.  "gemma": the installed config class, with the two adaRMS fields
        of the reference's patched GemmaConfig (configuration_gemma.py) attached as attributes; "paligemma": the stock
        config, optionally with a truncated vision tower."""

        def __getitem__(self, key):
            if key == "gemma":
                def make_gemma(**kw):
                    use_adarms = kw.pop("use_adarms", False)
                    cond = kw.pop("adarms_cond_dim", None)
                    kw.pop("torch_dtype", None)
                    c = gcfg(**kw)
                    c.use_adarms = use_adarms
                    c.adarms_cond_dim = cond
                    return c

                return make_gemma
            if key == "paligemma":
                def make():
                    c = pcfg_cls()
                    for attr, default in (("pad_token_id", None),):  # PretrainedConfig defaults of 4.53.2 that 5.5 dropped
                        if not hasattr(c, attr):
                            setattr(c, attr, default)
                    if vision_layers is not None:
                        c.vision_config.num_hidden_layers = vision_layers
                    return c

                return make
            return CONFIG_MAPPING[key]

    # openpi package skeleton without jax
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("jax", jit=lambda f=None, **k: f if f is not None else (lambda g_: g_), image=_Anything(), Array=_Anything())
    mod("jax.numpy", **{n: _Anything() for n in ("uint8", "float32", "round", "pad")})
    sys.modules["jax"].numpy = sys.modules["jax.numpy"]
    at = mod("openpi.shared.array_typing", typecheck=lambda f: f, UInt8=_Anything(), Float=_Anything(), Array=_Anything())
    openpi = mod("openpi")
    openpi.__path__ = []
    shared = mod("openpi.shared", array_typing=at)
    shared.__path__ = []
    openpi.shared = shared
    models = mod("openpi.models")
    models.__path__ = []
    openpi.models = models
    models.gemma = mod("openpi.models.gemma", get_config=lambda variant: _VARIANTS[variant])
    mp = mod("openpi.models_pytorch")
    mp.__path__ = []
    openpi.models_pytorch = mp
    shared.image_tools = _exec_as("openpi.shared.image_tools", os.path.join(REF, "shared", "image_tools.py"))
    mp.preprocessing_pytorch = _exec_as("openpi.models_pytorch.preprocessing_pytorch",
                                        os.path.join(REF, "models_pytorch", "preprocessing_pytorch.py"))
    gp = _exec_as("openpi.models_pytorch.gemma_pytorch", os.path.join(REF, "models_pytorch", "gemma_pytorch.py"))
    gp.CONFIG_MAPPING = _CfgMap()
    mp.gemma_pytorch = gp
    real_compile = torch.compile
    torch.compile = lambda f=None, **k: f if f is not None else (lambda g_: g_)
    try:
        p0 = _exec_as("openpi.models_pytorch.pi0_pytorch", os.path.join(REF, "models_pytorch", "pi0_pytorch.py"))
    finally:
        pass
    p0._real_compile = real_compile  # construction of PI0Pytorch also calls torch.compile: keep the no-op installed
    mp.pi0_pytorch = p0
    _loaded = p0
    return p0


def restore_torch_compile():
    if _loaded is not None and hasattr(_loaded, "_real_compile"):
        torch.compile = _loaded._real_compile
