"""Drop-in mirror of the reference's `openpi.models_pytorch.pi0_pytorch` surface on top of libpi05.so.

Reference class: `PI0Pytorch` (src/openpi/models_pytorch/pi0_pytorch.py:84-461).  Same constructor argument, same
`forward(observation, actions, noise=None, time=None)` / `sample_actions(device, observation, noise=None,
num_steps=10)` signatures and return shapes, same state_dict keys / shapes / dtypes (gemma_pytorch.py:63-83), real
`.grad`s, `gradient_checkpointing_enable()` (a no-op: the engine keeps activations resident, nothing is recomputed)
and `paligemma_with_expert.to_bfloat16_for_selected_params()`.  All arithmetic happens in the CUDA engine; this file
only owns memory (flat per-dtype parameter/gradient arenas the `nn.Parameter`s are views of), RNG draws
(noise / time, exactly where the reference draws them) and the autograd plumbing.

There is no CPU path: calling `forward` / `sample_actions` without the built library on an sm_100 GPU raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import logging
import math
from collections import OrderedDict

import torch
from torch import Tensor, nn

from . import _lib

logger = logging.getLogger("kai0_b200")

IMAGE_KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")  # preprocessing_pytorch.py:11-15
IMAGE_RESOLUTION = (224, 224)
PALIGEMMA_VOCAB_SIZE = 257_152  # models/gemma.py:40


@dataclasses.dataclass(frozen=True)
class GemmaVariant:
    """models/gemma.py:43-51 (without LoRA: full fine-tune is the path in scope)."""

    width: int
    depth: int
    mlp_dim: int
    num_heads: int
    num_kv_heads: int
    head_dim: int


def get_gemma_config(variant) -> GemmaVariant:
    """models/gemma.py:58-110.  Accepts a variant name or an object with the six fields (duck-typed)."""
    if not isinstance(variant, str):
        return GemmaVariant(
            variant.width, variant.depth, variant.mlp_dim, variant.num_heads, variant.num_kv_heads, variant.head_dim
        )
    if variant == "dummy":
        return GemmaVariant(64, 4, 128, 8, 1, 16)
    if variant == "gemma_300m":
        return GemmaVariant(1024, 18, 4096, 8, 1, 256)
    if variant == "gemma_2b":
        return GemmaVariant(2048, 18, 16_384, 8, 1, 256)
    raise ValueError(f"Unknown variant: {variant}")


@dataclasses.dataclass
class Pi05EngineConfig:
    """What PI0Pytorch.__init__ reads from its config (pi0_pytorch.py:85-98,307,380) plus engine-side knobs.
    Any object with these attribute names works (the reference's `Pi0Config` does)."""

    pi05: bool = True
    paligemma_variant: object = "gemma_2b"
    action_expert_variant: object = "gemma_300m"
    dtype: str = "bfloat16"
    action_dim: int = 32
    action_horizon: int = 50
    max_token_len: int = 200
    # vision tower (gemma_pytorch.py:38-41 + SiglipVisionConfig defaults of the PaliGemma config)
    vit_width: int = 1152
    vit_depth: int = 27
    vit_mlp_dim: int = 4304
    vit_heads: int = 16
    vit_patch: int = 14
    image_size: int = 224
    vocab_size: int = PALIGEMMA_VOCAB_SIZE
    num_images: int = 3


def _cfg_get(config, name, default):
    return getattr(config, name, default)


# ------------------------------------------------------------------------------------------------------------
# parameter table: reference names, shapes, dtypes, in ARENA ORDER (q|k|v and gate|up adjacent: the engine
# treats them as single fused weights)
# ------------------------------------------------------------------------------------------------------------
_PWE = "paligemma_with_expert."
_VT = _PWE + "paligemma.model.vision_tower.vision_model."
_LM = _PWE + "paligemma.model.language_model."
_EX = _PWE + "gemma_expert.model."
_KEEP_F32 = (  # gemma_pytorch.py:72-79
    "vision_tower.vision_model.embeddings.patch_embedding.weight",
    "vision_tower.vision_model.embeddings.patch_embedding.bias",
    "vision_tower.vision_model.embeddings.position_embedding.weight",
    "input_layernorm",
    "post_attention_layernorm",
    "model.norm",
)


def parameter_table(cfg: Pi05EngineConfig, pg: GemmaVariant, ex: GemmaVariant, value_head: bool = False):
    """[(name, shape, dtype, init_kind)] for every parameter of the reference module tree."""
    out = []

    def add(name, shape, kind):
        dt = torch.bfloat16 if name.startswith(_PWE) else torch.float32
        if name.startswith(_PWE) and any(k in name for k in _KEEP_F32):
            dt = torch.float32
        out.append((name, tuple(int(s) for s in shape), dt, kind))

    W, p = cfg.vit_width, cfg.vit_patch
    T = (cfg.image_size // p) ** 2
    add(_VT + "embeddings.patch_embedding.weight", (W, 3, p, p), "linear")
    add(_VT + "embeddings.patch_embedding.bias", (W,), "zeros")
    add(_VT + "embeddings.position_embedding.weight", (T, W), "embed")
    for i in range(cfg.vit_depth):
        L = f"{_VT}encoder.layers.{i}."
        add(L + "layer_norm1.weight", (W,), "ones")
        add(L + "layer_norm1.bias", (W,), "zeros")
        add(L + "layer_norm2.weight", (W,), "ones")
        add(L + "layer_norm2.bias", (W,), "zeros")
        for pj in ("q_proj", "k_proj", "v_proj"):
            add(L + f"self_attn.{pj}.weight", (W, W), "linear")
        for pj in ("q_proj", "k_proj", "v_proj"):
            add(L + f"self_attn.{pj}.bias", (W,), "zeros")
        add(L + "self_attn.out_proj.weight", (W, W), "linear")
        add(L + "self_attn.out_proj.bias", (W,), "zeros")
        add(L + "mlp.fc1.weight", (cfg.vit_mlp_dim, W), "linear")
        add(L + "mlp.fc1.bias", (cfg.vit_mlp_dim,), "zeros")
        add(L + "mlp.fc2.weight", (W, cfg.vit_mlp_dim), "linear")
        add(L + "mlp.fc2.bias", (W,), "zeros")
    add(_VT + "post_layernorm.weight", (W,), "ones")
    add(_VT + "post_layernorm.bias", (W,), "zeros")
    D = pg.width
    add(_PWE + "paligemma.model.multi_modal_projector.linear.weight", (D, W), "linear")
    add(_PWE + "paligemma.model.multi_modal_projector.linear.bias", (D,), "zeros")
    add(_LM + "embed_tokens.weight", (cfg.vocab_size, D), "embed")
    for prefix, g, ada in ((_LM, pg, False), (_EX, ex, True)):
        for i in range(g.depth):
            L = f"{prefix}layers.{i}."
            add(L + "self_attn.q_proj.weight", (g.num_heads * g.head_dim, g.width), "linear")
            add(L + "self_attn.k_proj.weight", (g.num_kv_heads * g.head_dim, g.width), "linear")
            add(L + "self_attn.v_proj.weight", (g.num_kv_heads * g.head_dim, g.width), "linear")
            add(L + "self_attn.o_proj.weight", (g.width, g.num_heads * g.head_dim), "linear")
            add(L + "mlp.gate_proj.weight", (g.mlp_dim, g.width), "linear")
            add(L + "mlp.up_proj.weight", (g.mlp_dim, g.width), "linear")
            add(L + "mlp.down_proj.weight", (g.width, g.mlp_dim), "linear")
            for nm in ("input_layernorm", "post_attention_layernorm"):
                if ada:
                    add(L + nm + ".dense.weight", (3 * g.width, g.width), "zeros")  # modeling_gemma.py:59-61
                    add(L + nm + ".dense.bias", (3 * g.width,), "zeros")
                else:
                    add(L + nm + ".weight", (g.width,), "zeros")  # modeling_gemma.py:63
        if ada:
            add(prefix + "norm.dense.weight", (3 * g.width, g.width), "zeros")
            add(prefix + "norm.dense.bias", (3 * g.width,), "zeros")
        else:
            add(prefix + "norm.weight", (g.width,), "zeros")
    E = ex.width
    add("action_in_proj.weight", (E, cfg.action_dim), "linear")
    add("action_in_proj.bias", (E,), "zeros")
    add("action_out_proj.weight", (cfg.action_dim, E), "linear")
    add("action_out_proj.bias", (cfg.action_dim,), "zeros")
    add("time_mlp_in.weight", (E, E), "linear")
    add("time_mlp_in.bias", (E,), "zeros")
    add("time_mlp_out.weight", (E, E), "linear")
    add("time_mlp_out.bias", (E,), "zeros")
    if value_head:
        add("value_head.0.weight", (E, E), "linear")
        add("value_head.0.bias", (E,), "zeros")
        add("value_head.2.weight", (E, E), "linear")
        add("value_head.2.bias", (E,), "zeros")
        add("value_head.4.weight", (1, E), "linear")
        add("value_head.4.bias", (1,), "zeros")
    # the expert's never-used lm_head (kept so checkpoints round-trip).  LAST in the bf16 arena so that
    # flat_parameters() can expose "everything that trains" as one contiguous prefix.
    add(_PWE + "gemma_expert.lm_head.weight", (cfg.vocab_size, ex.width), "embed")
    return out


_UNUSED = (_PWE + "gemma_expert.lm_head.weight",)  # parameters the path never reads (no gradient)

# Order in which the reference's module tree REGISTERS its parameters, i.e. the order of `model.parameters()` /
# `named_parameters()` of the reference `PI0Pytorch` (pi0_pytorch.py:92-109; gemma_pytorch.py:57-59;
# modeling_paligemma.py:138-149,389-393; modeling_siglip.py:212-231 embeddings, :435-442 encoder layer = layer_norm1,
# self_attn, layer_norm2, mlp, :348-369 attention = k_proj, v_proj, q_proj, out_proj, :420-426 mlp = fc1, fc2;
# modeling_gemma.py:332-342 decoder layer = self_attn (q, k, v, o :256-280), mlp (gate, up, down :113-121),
# input_layernorm, post_attention_layernorm; then the model's final norm and the causal-LM head).  It matters because
# `torch.optim.AdamW.state_dict()` -- the `optimizer.pt` of a checkpoint directory (train_pytorch.py:170,236-243) --
# keys its per-parameter state by POSITION in `model.parameters()`: with the same registration order a checkpoint
# written by the reference resumes on this module and vice versa.  (The arenas keep their own layout: `parameter_table`.)
_REG_TOP = ("paligemma_with_expert", "action_in_proj", "action_out_proj", "time_mlp_in", "time_mlp_out", "value_head")
_REG_VIT_LAYER = ("layer_norm1", "self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj", "self_attn.out_proj",
                  "layer_norm2", "mlp.fc1", "mlp.fc2")
_REG_GEMMA_LAYER = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj",
                    "mlp.up_proj", "mlp.down_proj", "input_layernorm", "post_attention_layernorm")


def _registration_key(name: str):
    """Sort key reproducing the reference's `named_parameters()` order (see _REG_* above)."""
    wb = 1 if name.endswith(".bias") else 0
    top = name.split(".", 1)[0]
    if top != "paligemma_with_expert":
        # nn.Linear / nn.Sequential heads: weight then bias, Sequential children by index
        idx = int(name.split(".")[1]) if top == "value_head" else 0
        return (_REG_TOP.index(top), idx, wb)

    def in_layer(rest: str, stems):
        i, tail = rest.split(".", 1)
        stem = next(s for s in stems if tail.startswith(s + "."))
        return (int(i), stems.index(stem), wb)

    if name.startswith(_VT):
        rest = name[len(_VT):]
        if rest.startswith("embeddings.patch_embedding."):
            k = (0, 0, 0, wb)
        elif rest.startswith("embeddings.position_embedding."):
            k = (0, 1, 0, 0)
        elif rest.startswith("encoder.layers."):
            k = (1, *in_layer(rest[len("encoder.layers."):], _REG_VIT_LAYER))
        else:  # post_layernorm
            k = (2, 0, 0, wb)
        return (0, 0, 0, *k)
    if name.startswith(_PWE + "paligemma.model.multi_modal_projector."):
        return (0, 0, 1, wb)
    for which, prefix in ((0, _LM), (1, _EX)):
        if name.startswith(prefix):
            rest = name[len(prefix):]
            if rest.startswith("embed_tokens."):
                k = (0, 0, 0, 0)
            elif rest.startswith("layers."):
                k = (1, *in_layer(rest[len("layers."):], _REG_GEMMA_LAYER))
            else:  # final norm
                k = (2, 0, 0, wb)
            # paligemma: model.language_model is the third child of paligemma.model; expert: its own model
            return (0, 0, 2, *k) if which == 0 else (0, 1, 0, *k)
    if name == _PWE + "paligemma.lm_head.weight":
        return (0, 0, 3)
    if name == _PWE + "gemma_expert.lm_head.weight":
        return (0, 1, 1)
    raise KeyError(name)


class _Node(nn.Module):
    """Parameter-only skeleton module: gives parameters the reference's dotted state_dict paths."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("skeleton module of the pi0.5 engine: call the top-level PI0Pytorch instead")

    # `layers` containers index like the reference's nn.ModuleList (children are named "0", "1", ...)
    def __getitem__(self, i):
        return self._modules[str(int(i))]

    def __len__(self):
        return sum(1 for k in self._modules if k.isdigit())

    def _apply(self, fn, recurse=True):
        """The parameters below this node are views of the top-level module's flat arenas, which only PI0Pytorch._apply
        moves (it never recurses into the skeleton).  Reaching this method means `.to()` / `.cuda()` / `.half()` was called
        on a SUB-module: that would re-point its parameters away from the arenas the engine reads."""
        raise RuntimeError(
            "kai0_b200: move / cast the whole PI0Pytorch module, not a sub-module: its parameters are views of the "
            "top-level module's flat arenas (the reference's mixed dtype map is fixed: gemma_pytorch.py:63-83)")


class _PaliGemmaWithExpert(_Node):
    def to_bfloat16_for_selected_params(self, precision: str = "bfloat16"):
        """gemma_pytorch.py:63-83.  The engine's parameters are created with exactly that dtype map, so this only
        validates the request."""
        if precision != "bfloat16":
            raise ValueError(f"the B200 engine implements the bfloat16 dtype map only (got {precision!r})")
        return self


def _attach(root: nn.Module, dotted: str, param: nn.Parameter, top_cls=_Node):
    parts = dotted.split(".")
    mod = root
    for i, name in enumerate(parts[:-1]):
        if name not in mod._modules:
            cls = _PaliGemmaWithExpert if (i == 0 and name == "paligemma_with_expert") else _Node
            mod.add_module(name, cls())
        mod = mod._modules[name]
    mod.register_parameter(parts[-1], param)


class _PatchRows:
    """The preprocessed images of one batch, already laid out as the patch-embedding GEMM operand (pi05_batch.patch_rows)."""

    def __init__(self, rows: Tensor, num_images: int):
        self.rows = rows
        self.num_images = num_images

    def __len__(self):
        return self.num_images


class _EngineFunction(torch.autograd.Function):
    """One autograd node for the whole network: forward = pi05_forward, backward = pi05_backward.

    The engine keeps ONE activation stash per training engine, so a node's backward is only valid while its forward is
    still the engine's most recent training forward: every training forward stamps a generation number, backward checks
    it and raises instead of silently using another batch's stash (the reference's autograd graph would keep both)."""

    @staticmethod
    def forward(ctx, model, batch, actions, noise, time, *params):
        ctx.model = model
        loss = model._engine_forward(batch, actions, noise, time)
        ctx.engine = model._engine
        ctx.generation = model._train_generation
        return loss

    @staticmethod
    def backward(ctx, dloss):
        model = ctx.model
        if ctx.engine is not model._train_engine_handle() or ctx.generation != model._train_generation:
            raise RuntimeError(
                "PI0Pytorch (B200 engine): backward() of a forward whose activation stash is gone - another training "
                "forward ran on this model (or the engine was re-created by .to() / a different batch size) before this "
                "backward.  The engine keeps one stash: call backward() before the next training forward "
                "(no_grad / eval forwards, sample_actions and sample_values use a separate engine and are fine).")
        params = [p for _, p in model._grad_params]
        if not model.direct_grads:
            # autograd path (DDP hooks, plain .grad): the engine OVERWRITES its flat gradient arena on every backward.
            # A .grad that is still a view of the arena (autograd steals the tensor we return when .grad was None) would
            # be overwritten in place and then doubled by AccumulateGrad: detach such gradients from the arena first,
            # and hand out copies whenever something is already accumulated (zero_grad(set_to_none=False), gradient
            # accumulation).  The reference loop (set_to_none=True, one backward per step) never takes this branch.
            accumulating = False
            for p in params:
                if p.grad is not None:
                    accumulating = True
                    if model._aliases_grad_arena(p.grad):
                        p.grad = p.grad.detach().clone()
        grads = model._engine_backward(dloss.contiguous())
        if model.direct_grads:
            # engine-owned gradients: .grad of every parameter is (a view of) the flat gradient arena; autograd only
            # sees the anchor.  Skips ~700 per-parameter AccumulateGrad copies per step.  Overwrite semantics.
            if model._flat_params is not None:
                model._sync_flat_param_grads()  # the caller optimises the two flat tensors
            for (_, p), g in zip(model._grad_params, grads):
                p.grad = g
            return (None, None, None, None, None, torch.zeros_like(model._grad_anchor))
        return (None, None, None, None, None, *[(g.clone() if accumulating else g) for g in grads])


class PI0Pytorch(nn.Module):
    """B200-native pi0.5 model with the reference's surface (pi0_pytorch.py:84-461)."""

    _value_head = False

    def __init__(self, config, *, max_batch: int | None = None, init_weights: bool = True):
        super().__init__()
        self.config = config
        self.pi05 = bool(_cfg_get(config, "pi05", True))
        if not self.pi05:
            raise ValueError("the B200 engine implements the pi0.5 branch (config.pi05=True) only")
        if _cfg_get(config, "dtype", "bfloat16") != "bfloat16":
            raise ValueError("the B200 engine implements dtype='bfloat16' only")
        d = Pi05EngineConfig()
        self.ecfg = Pi05EngineConfig(
            pi05=True,
            paligemma_variant=_cfg_get(config, "paligemma_variant", d.paligemma_variant),
            action_expert_variant=_cfg_get(config, "action_expert_variant", d.action_expert_variant),
            action_dim=_cfg_get(config, "action_dim", d.action_dim),
            action_horizon=_cfg_get(config, "action_horizon", d.action_horizon),
            max_token_len=_cfg_get(config, "max_token_len", d.max_token_len) or d.max_token_len,
            vit_width=_cfg_get(config, "vit_width", d.vit_width),
            vit_depth=_cfg_get(config, "vit_depth", d.vit_depth),
            vit_mlp_dim=_cfg_get(config, "vit_mlp_dim", d.vit_mlp_dim),
            vit_heads=_cfg_get(config, "vit_heads", d.vit_heads),
            vit_patch=_cfg_get(config, "vit_patch", d.vit_patch),
            image_size=_cfg_get(config, "image_size", d.image_size),
            vocab_size=_cfg_get(config, "vocab_size", d.vocab_size),
            num_images=_cfg_get(config, "num_images", d.num_images),
        )
        self.pg = get_gemma_config(self.ecfg.paligemma_variant)
        self.ex = get_gemma_config(self.ecfg.action_expert_variant)
        self._max_batch_hint = max_batch
        self.gradient_checkpointing_enabled = False
        # train-time image augmentation (preprocessing_pytorch.py:52-142).  The reference's forward() always calls its
        # preprocessing with train=True (pi0_pytorch.py:318), so this is on by default; parity tests that compare
        # against un-augmented oracles switch it off or inject the drawn parameters (_augment_params_override).
        self.augment = True
        # preprocessing writes straight into the patch-embedding GEMM operand (row f2); False: fp32 NCHW images + the fp32
        # im2col convolution on CUDA cores (round-1 path, kept as the cross-check in tests/test_preprocess_gpu.py)
        self.use_patch_rows = True
        # Prompt padding removal: slots that are padding in EVERY sample of the batch (left-aligned masks, as the reference's
        # tokenizer produces: models/tokenizer.py:35-38) are dropped before the engine call (pi05_batch.token_len).  Outputs
        # and gradients are unchanged: a padded slot is a masked key (probability exactly 0) and its own row feeds nothing.
        # Off automatically while taps are recorded (parity tests compare full-length intermediates).
        self.skip_prompt_padding = True
        self._token_len_cache = []
        self._augment_params_override = None
        self._pre_scratch = None

        # ---- flat arenas + parameter views
        table = parameter_table(self.ecfg, self.pg, self.ex, self._value_head)
        self._table = table
        n_bf16 = sum(math.prod(s) for _, s, dt, _ in table if dt == torch.bfloat16)
        n_f32 = sum(math.prod(s) for _, s, dt, _ in table if dt == torch.float32)
        # 8-element alignment of every tensor inside an arena (TMA needs 16B-aligned bases)
        self._offsets = OrderedDict()
        off = {torch.bfloat16: 0, torch.float32: 0}
        for name, shape, dt, _ in table:
            n = math.prod(shape)
            self._offsets[name] = (dt, off[dt], n, shape)
            off[dt] += (n + 7) // 8 * 8
        self._flat = {
            torch.bfloat16: torch.zeros(off[torch.bfloat16], dtype=torch.bfloat16),
            torch.float32: torch.zeros(off[torch.float32], dtype=torch.float32),
        }
        del n_bf16, n_f32
        self._flat_grad = {torch.bfloat16: None, torch.float32: None}
        # registered in the REFERENCE'S order (see _registration_key), independent of the arena layout above
        specs = {name: (shape, dt) for name, shape, dt, _ in table}
        for name in sorted(specs, key=_registration_key):
            shape, dt = specs[name]
            _, o, n, _ = self._offsets[name]
            p = nn.Parameter(self._flat[dt][o : o + n].view(shape), requires_grad=True)
            _attach(self, name, p)
        # tied head (modeling_paligemma.py:389-393): same Parameter object under the second name
        self.paligemma_with_expert.paligemma.add_module("lm_head", _Node())
        self.paligemma_with_expert.paligemma.lm_head.register_parameter(
            "weight", self.paligemma_with_expert.paligemma.model.language_model.embed_tokens.weight
        )
        if init_weights:
            self.reset_parameters()
        self.check_inputs = True  # validate token ids on the host (one sync); benchmarks turn it off
        # False (default): gradients flow through autograd per parameter (DDP hooks work, as train_pytorch.py:441 needs).
        # True: backward assigns .grad = view of the flat gradient arena directly (single GPU or enable_flat_allreduce).
        self.direct_grads = False
        self._grad_anchor = None
        self._flat_params = None
        self._flat_used_bf16 = 0
        self.use_cuda_graph = True  # sample_actions replays one captured CUDA graph per (batch, num_steps)
        self._graphs = {}

        self._engine = None        # the ACTIVE engine handle (one of self._engines)
        self._engine_key = None    # (device index, train, max batch, num_images) of the active engine
        self._workspace = None
        self._engines = OrderedDict()  # (device index, train, num_images) -> dict(handle, workspace, max_batch, graphs); LRU order
        self._train_generation = 0
        self._dp_group = None
        self._dp_comm = None
        self._dp_overlap = False
        self._dp_engine = False
        self._dp_average = "in_place"
        self._dp_world = 1
        self._dp_max_ctas = 0
        self.grad_scale = 1.0  # what an optimiser must multiply .grad by (1/world with average="optimizer")
        self._keep = None
        # parameters the reference's autograd leaves WITHOUT a gradient (nothing on the loss path reads the last layer's
        # prefix-stream output or the final prefix norm, pi0_pytorch.py:350-358): .grad stays None for them, so a stock
        # optimiser creates no state and applies no weight decay, exactly as with the reference module
        last = f"{_LM}layers.{self.pg.depth - 1}."
        self._dead_grad_names = frozenset(
            [last + "self_attn.o_proj.weight", last + "mlp.gate_proj.weight", last + "mlp.up_proj.weight",
             last + "mlp.down_proj.weight", last + "post_attention_layernorm.weight", _LM + "norm.weight"]
        ) if self.pg.depth > 0 else frozenset([_LM + "norm.weight"])

    # ------------------------------------------------------------------ init / housekeeping
    @torch.no_grad()
    def reset_parameters(self, seed: int | None = None):
        """Seeded from-scratch initialisation in the reference's DISTRIBUTIONS (not its RNG stream: a checkpoint is what
        makes two runs comparable, train_pytorch.py:450-460).  Transformer weights and embeddings N(0, 0.02) (HF
        `initializer_range`), LayerNorm ones / zeros, RMSNorm weight and adaRMS dense zeros (modeling_gemma.py:59-63),
        and for the fp32 heads outside the HF stack (`action_in/out_proj`, `time_mlp_*`, `value_head.*`: plain nn.Linear in
        pi0_pytorch.py:100-109) torch's nn.Linear default: weight and bias U(-1/sqrt(fan_in), +1/sqrt(fan_in)).  The
        SigLIP patch convolution / position embedding also get N(0, 0.02) (HF uses lecun-normal / width-scaled normal
        there; immaterial once `safetensors.load_model` has run)."""
        dev = self._device()
        g = None
        if seed is not None:
            g = torch.Generator(device=dev).manual_seed(seed)
        params = dict(self.named_parameters())
        for name, shape, dt, kind in self._table:
            p = params[name]
            if not name.startswith(_PWE):  # plain nn.Linear heads: kaiming_uniform_(a=sqrt(5)) + uniform bias
                w_shape = shape if name.endswith(".weight") else params[name[: -len("bias")] + "weight"].shape
                bound = 1.0 / math.sqrt(math.prod(w_shape[1:]))
                p.uniform_(-bound, bound, generator=g)
            elif kind == "zeros":
                p.zero_()
            elif kind == "ones":
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02, generator=g)  # drawn on the parameter's own device

    def _apply(self, fn, recurse=True):
        """Keep the parameters views of the flat arenas across .to()/.cuda(): move the arenas, re-point the views."""
        new_flat = {}
        for dt, t in self._flat.items():
            nt = fn(t)
            if nt.dtype != dt:
                raise TypeError(
                    "PI0Pytorch keeps the reference's mixed dtype map (bf16 weights, fp32 norms/heads); "
                    f"casting the whole module to {nt.dtype} is not supported"
                )
            new_flat[dt] = nt
        moved = any(new_flat[dt] is not self._flat[dt] for dt in new_flat)
        if moved:
            self._flat = new_flat
            self._flat_grad = {torch.bfloat16: None, torch.float32: None}
            self._flat_params = None
            params = dict(self.named_parameters())
            for name, (dt, o, n, shape) in self._offsets.items():
                p = params[name]
                p.data = self._flat[dt][o : o + n].view(shape)
                p.grad = None
            self._destroy_engine()
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        if assign:
            raise ValueError("kai0_b200: load_state_dict(assign=True) would replace the parameters by tensors outside the "
                             "flat arenas the engine reads; load with assign=False (values are copied into the arenas)")
        return super().load_state_dict(state_dict, strict=strict)

    def state_dict(self, *args, **kwargs):
        """Checkpoints must not see the flat arenas: `safetensors.torch.save_model` (train_pytorch.py:167) refuses
        tensors that are partial views of a larger storage.  Values are therefore returned as independent copies
        (the tied `paligemma.lm_head.weight` / `embed_tokens.weight` pair shares ONE copy, as in the reference)."""
        sd = super().state_dict(*args, **kwargs)
        if kwargs.get("keep_vars", False):
            return sd
        seen = {}
        for k in list(sd.keys()):
            t = sd[k]
            key = (t.data_ptr(), t.numel(), t.dtype)
            if key not in seen:
                seen[key] = t.detach().clone()
            sd[k] = seen[key]
        return sd

    def gradient_checkpointing_enable(self):
        """pi0_pytorch.py:126-133.  Accepted for script compatibility; the engine never recomputes (180 GB HBM)."""
        self.gradient_checkpointing_enabled = True
        logger.info("gradient checkpointing requested: no-op on the B200 engine (activations stay resident)")

    def gradient_checkpointing_disable(self):
        self.gradient_checkpointing_enabled = False

    def is_gradient_checkpointing_enabled(self):
        return self.gradient_checkpointing_enabled

    def flat_parameters(self):
        """Opt-in fast path for the caller-side optimiser step: the two flat arenas as TWO
        nn.Parameters (bf16 and fp32) aliasing exactly the memory of all trainable named parameters (the unused expert
        lm_head sits after the bf16 prefix and is excluded).  Their `.grad` are the flat gradient arenas, kept in sync
        after every backward when `direct_grads` is on.  AdamW / clip_grad_norm_ over these two tensors is
        element-for-element what the per-parameter calls compute (both are element-wise / one global norm), in 2
        launches instead of ~1400.  Use either these or `parameters()` with an optimiser, not both."""
        if self._flat_params is None:
            dt_b, dt_f = torch.bfloat16, torch.float32
            _, o, _, _ = self._offsets[_UNUSED[0]]
            self._flat_params = [nn.Parameter(self._flat[dt_b][:o]), nn.Parameter(self._flat[dt_f])]
            self._flat_used_bf16 = o
        return self._flat_params

    def _sync_flat_param_grads(self):
        if self._flat_params is None:
            return
        gb, gf = self._flat_grad[torch.bfloat16], self._flat_grad[torch.float32]
        self._flat_params[0].grad = gb[: self._flat_used_bf16]
        self._flat_params[1].grad = gf

    def enable_flat_allreduce(self, process_group=None, *, overlap: bool = False, average: str = "in_place",
                              max_ctas: int | None = None):
        """Engine-owned data parallelism, INSTEAD of wrapping the module in DistributedDataParallel
        (train_pytorch.py:440-447): the gradient arenas are all-reduced over the group and averaged, as DDP's bucketed
        all-reduce does.  On a CUDA module whose group runs on NCCL the ENGINE issues the collectives from inside
        pi05_backward on a communicator of its own (pi05_set_grad_exchange; the unique id travels through
        torch.distributed); otherwise (gloo / CPU tests) two torch.distributed all-reduces follow backward.

        overlap=False (default): one exchange at the end of backward at NCCL's full speed.  overlap=True: chunk by chunk
        while backward is still running, on a communicator restricted to `max_ctas` SMs (default 4, PI05_NCCL_MAX_CTAS) --
        measured slower on power-capped B200s (see include/pi05.h), kept as an option.
        average="in_place": .grad holds the average (any optimiser works);  "optimizer": .grad holds the SUM and
        `kai0_b200.optim.FusedClipAdamW` applies 1 / world on the fly (saves a 14 GB read-modify-write per step)."""
        import torch.distributed as dist

        if average not in ("in_place", "optimizer"):
            raise ValueError("average must be 'in_place' or 'optimizer'")
        self._dp_group = process_group if process_group is not None else dist.group.WORLD
        self._dp_world = dist.get_world_size(self._dp_group)
        self._dp_engine = self._device().type == "cuda" and "nccl" in str(dist.get_backend(self._dp_group))
        if overlap and not self._dp_engine:
            raise RuntimeError("enable_flat_allreduce(overlap=True) needs a CUDA module and an NCCL process group")
        self._dp_overlap = bool(overlap)
        self._dp_average = average
        self.grad_scale = 1.0 / self._dp_world if average == "optimizer" else 1.0
        self.direct_grads = True  # no DDP hooks to feed, so gradients can be handed over without autograd copies
        if self._dp_engine and self._dp_comm is None:
            import os

            dev = self._device()
            if max_ctas is None:
                # overlapped: at most 4 CTAs; at the end of backward: at least 32 (negative = minCTAs, include/pi05.h)
                max_ctas = (int(os.environ.get("PI05_NCCL_MAX_CTAS", "4")) if overlap
                            else -int(os.environ.get("PI05_NCCL_MIN_CTAS", "32")))
            rank = dist.get_rank(self._dp_group)
            uid = C.create_string_buffer(128)
            l = _lib.lib()
            if rank == 0:
                _lib.check(l.pi05_nccl_unique_id(uid), "pi05_nccl_unique_id")
            box = [uid.raw]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self._dp_group, 0), group=self._dp_group)
            comm = C.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(l.pi05_nccl_comm_create(C.c_char_p(box[0]), self._dp_world, rank, int(max_ctas), C.byref(comm)),
                           "pi05_nccl_comm_create")
            self._dp_comm = comm
            self._dp_max_ctas = int(max_ctas)
        for ent in self._engines.values():
            self._apply_exchange(ent["handle"])

    def _apply_exchange(self, handle):
        if self._dp_group is None or not self._dp_engine or self._dp_comm is None:
            return
        _lib.check(_lib.lib().pi05_set_grad_exchange(handle, self._dp_comm, self._dp_world,
                                                     1 if self._dp_average == "in_place" else 0,
                                                     1 if self._dp_overlap else 0), "pi05_set_grad_exchange")

    def exchange_description(self) -> str:
        if self._dp_group is None:
            return "none (single rank)"
        if not self._dp_engine:
            return "two torch.distributed all-reduces (bf16 arena, fp32 arena) after backward, then 1/world"
        calls, nbytes = C.c_int64(), C.c_int64()
        h = self._train_engine_handle()
        if h is not None:
            _lib.lib().pi05_grad_exchange_stats(h, C.byref(calls), C.byref(nbytes))
        how = ("per gradient group, overlapped with backward on a high-priority stream" if self._dp_overlap
               else "of the two gradient arenas at the end of pi05_backward")
        ctas = f"maxCTAs {self._dp_max_ctas}" if self._dp_max_ctas > 0 else f"minCTAs {-self._dp_max_ctas}"
        return (f"engine-issued ncclAllReduce(sum) {how} (own communicator, {ctas}; last backward: {calls.value} collectives, {nbytes.value / 1e9:.2f} GB); "
                f"average {self._dp_average}")

    # ------------------------------------------------------------------ engine lifecycle
    def _device(self):
        return self._flat[torch.bfloat16].device

    def _destroy_engine(self):
        """Destroys EVERY cached engine (parameters moved / module deleted)."""
        for ent in self._engines.values():
            _lib.lib().pi05_destroy(ent["handle"])
        self._engines.clear()
        self._engine = None
        self._engine_key = None
        self._workspace = None
        self._graphs = {}

    def _evict_engine(self, key):
        ent = self._engines.pop(key)
        if self._engine is ent["handle"]:
            self._engine, self._engine_key, self._workspace, self._graphs = None, None, None, {}
        _lib.lib().pi05_destroy(ent["handle"])
        ent["workspace"] = None

    def _train_engine_handle(self):
        """Handle of the training engine whose stash the last training forward filled (None if it was evicted)."""
        ent = self._engines.get(getattr(self, "_train_key", None))
        return ent["handle"] if ent is not None else None

    def __del__(self):
        try:
            self._destroy_engine()
        except Exception:  # noqa: BLE001
            pass

    def _ensure_engine(self, batch: int, train: bool, num_images: int | None = None, rtc: bool = False):
        """Makes the engine for (device, train, num_images) the active one, creating it on first use.  Engines are kept
        (an AdvantageEstimator alternating 3- and 6-image calls, or a validation forward between training steps, does not
        re-plan 100+ GB of workspace each time); when a new workspace does not fit in free device memory the least
        recently used other engines are destroyed first."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                "PI0Pytorch (B200 engine) has no CPU path: move the module to an sm_100 CUDA device first"
            )
        ni = int(num_images or self.ecfg.num_images)
        key = (dev.index, bool(train), ni, bool(rtc) and not train)
        ent = self._engines.get(key)
        if ent is not None and ent["max_batch"] < batch:
            self._evict_engine(key)  # grown batch: re-plan this one
            ent = None
        if ent is None:
            ent = self._create_engine(dev, key, max(batch, self._max_batch_hint or 0))
            self._engines[key] = ent
        self._engines.move_to_end(key)
        self._engine = ent["handle"]
        self._engine_ent = ent
        self._workspace = ent["workspace"]
        self._graphs = ent["graphs"]
        self._engine_key = (dev.index, bool(train), ent["max_batch"], ni)

    def _create_engine(self, dev, key, need_b):
        _, train, ni, rtc = key
        l = _lib.lib()
        c = _lib.Config()
        for dst, src in ((c.paligemma, self.pg), (c.expert, self.ex)):
            dst.width, dst.depth, dst.mlp_dim = src.width, src.depth, src.mlp_dim
            dst.num_heads, dst.num_kv_heads, dst.head_dim = src.num_heads, src.num_kv_heads, src.head_dim
        e = self.ecfg
        c.vit_width, c.vit_depth, c.vit_mlp_dim, c.vit_heads = e.vit_width, e.vit_depth, e.vit_mlp_dim, e.vit_heads
        c.vit_patch, c.image_size, c.vocab_size = e.vit_patch, e.image_size, e.vocab_size
        c.action_dim, c.action_horizon, c.max_token_len = e.action_dim, e.action_horizon, e.max_token_len
        c.num_images, c.max_batch, c.train = ni, need_b, 1 if train else 0
        c.value_head = 1 if self._value_head else 0
        c.rtc = 1 if rtc else 0
        nbytes = l.pi05_workspace_bytes(C.byref(c))
        if nbytes == 0:
            raise RuntimeError(f"pi05_workspace_bytes: {_lib.last_error()}")
        with torch.cuda.device(dev):
            need = nbytes + 256
            if train:
                need += sum(t.numel() * t.element_size() for dt, t in self._flat.items() if self._flat_grad[dt] is None)
            # LRU eviction until the new workspace fits (free memory incl. what torch's caching allocator can give back)
            while self._engines:
                free, _ = torch.cuda.mem_get_info(dev)
                cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
                if free + cached >= need + (1 << 30):
                    break
                self._evict_engine(next(iter(self._engines)))
                torch.cuda.empty_cache()
            workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            base = (workspace.data_ptr() + 255) // 256 * 256
            handle = C.c_void_p()
            _lib.check(
                l.pi05_create(C.byref(c), dev.index or 0, C.c_void_p(base), C.c_size_t(nbytes), C.byref(handle)),
                "pi05_create",
            )
        if train:
            for dt in (torch.bfloat16, torch.float32):
                if self._flat_grad[dt] is None:
                    self._flat_grad[dt] = torch.zeros_like(self._flat[dt])
        ent = {"handle": handle, "workspace": workspace, "max_batch": need_b, "graphs": {}}
        self._bind(handle, with_grads=train)
        if train:
            self._apply_exchange(handle)
        return ent

    def _aliases_grad_arena(self, t: Tensor) -> bool:
        for g in self._flat_grad.values():
            if g is not None and g.device == t.device:
                lo = g.data_ptr()
                if lo <= t.data_ptr() < lo + g.numel() * g.element_size():
                    return True
        return False

    def _bind(self, handle, with_grads: bool):
        names, arr = [], (_lib.Param * len(self._offsets))()
        for i, (name, (dt, o, n, _)) in enumerate(self._offsets.items()):
            esz = 2 if dt == torch.bfloat16 else 4
            bname = name.encode()
            names.append(bname)
            arr[i].name = bname
            arr[i].dtype = 1 if dt == torch.bfloat16 else 0
            arr[i].numel = n
            arr[i].data = self._flat[dt].data_ptr() + o * esz
            g = self._flat_grad[dt] if with_grads else None
            arr[i].grad = (g.data_ptr() + o * esz) if g is not None else None
        self._keep = names
        _lib.check(_lib.lib().pi05_bind_params(handle, arr, len(self._offsets)), "pi05_bind_params")

    def set_taps(self, enabled: bool):
        self._taps = bool(enabled)

    def get_tap(self, name: str) -> Tensor:
        """Debug/parity hook: a named intermediate of the last call (tests only)."""
        n, dt = C.c_int64(), C.c_int32()
        l = _lib.lib()
        _lib.check(l.pi05_get_tap(self._engine, name.encode(), None, C.byref(n), C.byref(dt), None), "pi05_get_tap")
        tdt = {0: torch.float32, 1: torch.bfloat16, 2: torch.int32, 3: torch.uint8}[dt.value]
        out = torch.empty(n.value, dtype=tdt, device=self._device())
        stream = C.c_void_p(torch.cuda.current_stream(self._device()).cuda_stream)
        _lib.check(
            l.pi05_get_tap(self._engine, name.encode(), C.c_void_p(out.data_ptr()), C.byref(n), C.byref(dt), stream),
            "pi05_get_tap",
        )
        return out

    # ------------------------------------------------------------------ inputs
    _apply_aug = True  # AdvantageEstimator never augments (pi0_pytorch.py:488-489)

    def _draw_augment_params(self, keys, S, dev):
        """The random numbers of preprocessing_pytorch.py:60-137, drawn with the same torch calls in the same order
        (one draw per batch and key, on the images' device) and kept on the device: fp32 [len(keys), 6] =
        {start_h, start_w, angle_deg, brightness, contrast, saturation}."""
        if self._augment_params_override is not None:
            return self._augment_params_override.to(dev, torch.float32).contiguous()
        p = torch.zeros(len(keys), 6, dtype=torch.float32, device=dev)
        mx = S - int(S * 0.95)
        for i, key in enumerate(keys):
            if "wrist" not in key:
                if mx > 0:
                    p[i, 0:1] = torch.randint(0, mx + 1, (1,), device=dev)
                    p[i, 1:2] = torch.randint(0, mx + 1, (1,), device=dev)
                p[i, 2:3] = torch.rand(1, device=dev) * 10 - 5
            p[i, 3:4] = 0.7 + torch.rand(1, device=dev) * 0.6
            p[i, 4:5] = 0.6 + torch.rand(1, device=dev) * 0.8
            p[i, 5:6] = 0.5 + torch.rand(1, device=dev) * 1.0
        return p

    def _preprocess_observation(self, observation, *, train=True, rows=False, engine_train=None, engine_rtc=False):
        """preprocessing_pytorch.py:20-173 on the device, one launch group per image key: layout sniffing, resize-with-pad
        to image_size, train-time augmentation, default masks.  Images may be fp32 in [-1, 1] (what Observation.from_dict
        hands over) or uint8 -- then from_dict's `x / 255 * 2 - 1` (models/model.py:129-133) is taken inside the kernel.
        rows=False: returns the images stacked in the engine's layout, fp32 [num_images, B, 3, S, S] (pi05_preprocess_image).
        rows=True (what forward / sample_actions use, `use_patch_rows`): the kernels write straight into the operand of the
        patch-embedding GEMM (pi05_preprocess_patches) and a `_PatchRows` handle is returned."""
        images = getattr(observation, "images")
        keys = self._image_keys(images)
        state = observation.state
        batch_shape = state.shape[:-1]
        S = self.ecfg.image_size
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                "PI0Pytorch (B200 engine) has no CPU path: move the module to an sm_100 CUDA device first"
            )
        B = int(state.shape[0])
        aug = bool(train and self.augment and self._apply_aug)
        l = _lib.lib()
        patch = self.ecfg.vit_patch
        if rows:
            kp = int(l.pi05_patch_row_kp(patch))
            T = (S // patch) ** 2
            # Owned by the active engine (train and inference engines have their own, so a validation forward cannot
            # overwrite the rows a pending backward still reads) and zero-filled ONCE: the kernels never touch the padding
            # columns [3 p^2, Kp) of each block, which must read as 0 in the GEMM.
            ent = {}
            if engine_train is not None:  # the engine that will consume the rows becomes the active one first
                self._ensure_engine(B, train=bool(engine_train), num_images=len(keys), rtc=engine_rtc)
                ent = self._engine_ent
            out = ent.get("rows")
            if out is None or out.shape != (len(keys) * B * T, 3 * kp):
                out = torch.zeros((len(keys) * B * T, 3 * kp), dtype=torch.bfloat16, device=dev)
                ent["rows"] = out
        else:
            out = torch.empty((len(keys), B, 3, S, S), dtype=torch.float32, device=dev)
        need = l.pi05_preprocess_scratch_floats(B, S)
        if self._pre_scratch is None or self._pre_scratch.numel() < need or self._pre_scratch.device != dev:
            self._pre_scratch = torch.empty(need, dtype=torch.float32, device=dev)
        params = self._draw_augment_params(keys, S, dev) if aug else None
        out_masks = []
        masks = getattr(observation, "image_masks", {}) or {}
        stream = self._stream()
        keep = []
        for i, key in enumerate(keys):
            img = images[key]
            if img.dim() != 4 or img.shape[0] != B:
                raise ValueError(f"image {key} must be [B,3,H,W] or [B,H,W,3] with B={B}, got {tuple(img.shape)}")
            channels_first = img.shape[1] == 3  # preprocessing_pytorch.py:42 (same sniffing rule)
            if not channels_first and img.shape[-1] != 3:
                raise ValueError(f"image {key} has neither 3 channels first nor last: {tuple(img.shape)}")
            h, w = (img.shape[2], img.shape[3]) if channels_first else (img.shape[1], img.shape[2])
            if img.dtype == torch.uint8:
                if not rows:
                    # the fp32-image entry point mirrors the reference's function, which only ever sees from_dict's output
                    img = img.to(dev).to(torch.float32) / 255.0 * 2.0 - 1.0
            elif img.dtype != torch.float32:
                raise ValueError(f"image {key} must be uint8 or float32 in [-1, 1] (Observation.from_dict), got {img.dtype}")
            img = img.to(dev).contiguous()
            keep.append(img)
            p_i = C.c_void_p(params[i].data_ptr()) if aug else None
            if rows:
                _lib.check(
                    l.pi05_preprocess_patches(
                        C.c_void_p(img.data_ptr()), 3 if img.dtype == torch.uint8 else 0, int(h), int(w),
                        0 if channels_first else 1, B, S, patch, 1 if aug else 0, 0 if "wrist" in key else 1, p_i,
                        C.c_void_p(self._pre_scratch.data_ptr()),
                        C.c_void_p(out.data_ptr() + i * B * T * 3 * kp * 2), stream,
                    ),
                    "pi05_preprocess_patches",
                )
            else:
                _lib.check(
                    l.pi05_preprocess_image(
                        C.c_void_p(img.data_ptr()), int(h), int(w), 0 if channels_first else 1, B, S, 1 if aug else 0,
                        0 if "wrist" in key else 1, p_i, C.c_void_p(self._pre_scratch.data_ptr()),
                        C.c_void_p(out[i].data_ptr()), stream,
                    ),
                    "pi05_preprocess_image",
                )
            if key in masks:
                out_masks.append(masks[key])
            else:
                out_masks.append(torch.ones(batch_shape, dtype=torch.bool, device=state.device))
        if rows:
            out = _PatchRows(out, len(keys))
        return out, out_masks, observation.tokenized_prompt, observation.tokenized_prompt_mask, state

    def _image_keys(self, images):
        if not set(IMAGE_KEYS).issubset(images):
            raise ValueError(f"images dict missing keys: expected {IMAGE_KEYS}, got {list(images)}")
        return IMAGE_KEYS

    def _effective_token_len(self, lang_masks) -> int:
        """Number of leading prompt slots to keep: the longest valid prompt of the batch, rounded up to 8, when every mask
        row is left-aligned; the full length otherwise.  One small device->host read per NEW mask tensor (cached by
        tensor object + version, so a resident batch costs nothing)."""
        L = int(lang_masks.shape[1])
        if not self.skip_prompt_padding or getattr(self, "_taps", False):
            return L
        # cache by tensor OBJECT (weak reference) + version counter: a pointer would be reused by a later allocation
        import weakref

        for ref, version, value in self._token_len_cache:
            if ref() is lang_masks and version == lang_masks._version:
                return value
        m = lang_masks.to(torch.bool)
        lens = m.sum(dim=1)
        aligned = (m == (torch.arange(L, device=m.device)[None, :] < lens[:, None])).all()
        ok, mx = torch.stack([aligned.to(torch.int64), lens.max().to(torch.int64)]).tolist()
        out = L if not ok else min(L, max(8, (int(mx) + 7) // 8 * 8))
        self._token_len_cache = [e for e in self._token_len_cache if e[0]() is not None][-7:]
        self._token_len_cache.append((weakref.ref(lang_masks), lang_masks._version, out))
        return out

    def _make_batch(self, images, img_masks, lang_tokens, lang_masks):
        dev = self._device()
        if len(images) != self._engine_key[3]:
            raise ValueError(f"expected {self._engine_key[3]} images, got {len(images)}")
        rows = None
        if isinstance(images, _PatchRows):
            rows, imgs = images.rows, images.rows  # kept alive through `keep`
        elif isinstance(images, torch.Tensor):
            imgs = images  # already [num_images, B, 3, S, S] fp32 from _preprocess_observation
        else:
            imgs = torch.stack([i.to(dev, torch.float32) for i in images], dim=0).contiguous()
        masks = torch.stack([m.to(dev) for m in img_masks], dim=0).to(torch.uint8).contiguous()
        if lang_tokens.shape[1] != self.ecfg.max_token_len:
            raise ValueError(f"tokenized_prompt length {lang_tokens.shape[1]} != max_token_len {self.ecfg.max_token_len}")
        if self.check_inputs and (int(lang_tokens.min()) < 0 or int(lang_tokens.max()) >= self.ecfg.vocab_size):
            raise ValueError("token id out of range")
        token_len = self._effective_token_len(lang_masks)
        toks = lang_tokens[:, :token_len].to(dev, torch.int64).contiguous()
        tmask = lang_masks[:, :token_len].to(dev).to(torch.uint8).contiguous()
        B = toks.shape[0]
        b = _lib.Batch()
        b.batch = B
        b.images = imgs.data_ptr() if rows is None else None
        b.patch_rows = rows.data_ptr() if rows is not None else None
        b.image_masks = masks.data_ptr()
        b.tokens = toks.data_ptr()
        b.token_mask = tmask.data_ptr()
        b.token_len = token_len
        return b, (imgs, masks, toks, tmask)

    def sample_noise(self, shape, device):  # pi0_pytorch.py:172-179
        return torch.normal(mean=0.0, std=1.0, size=shape, dtype=torch.float32, device=device)

    def sample_time(self, bsize, device):  # pi0_pytorch.py:45-49,181-184
        alpha_t = torch.as_tensor(1.5, dtype=torch.float32, device=device)
        beta_t = torch.as_tensor(1.0, dtype=torch.float32, device=device)
        time_beta = torch.distributions.Beta(alpha_t, beta_t).sample((bsize,))
        return (time_beta * 0.999 + 0.001).to(dtype=torch.float32, device=device)

    # ------------------------------------------------------------------ engine calls
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device()).cuda_stream)

    def _engine_forward(self, batch_pack, actions, noise, time):
        b, keep = batch_pack
        B = b.batch
        loss = torch.empty((B, self.ecfg.action_horizon, self.ecfg.action_dim), dtype=torch.float32, device=self._device())
        l = _lib.lib()
        l.pi05_set_taps(self._engine, 1 if getattr(self, "_taps", False) else 0)
        _lib.check(
            l.pi05_forward(
                self._engine,
                C.byref(b),
                C.c_void_p(actions.data_ptr()),
                C.c_void_p(noise.data_ptr()),
                C.c_void_p(time.data_ptr()),
                C.c_void_p(loss.data_ptr()),
                self._stream(),
            ),
            "pi05_forward",
        )
        self._engine_ent["keep"] = (keep, actions, noise, time)  # device buffers stay alive until this engine's next call
        return loss

    def _engine_backward(self, dloss):
        handle = self._train_engine_handle()
        _lib.check(_lib.lib().pi05_backward(handle, C.c_void_p(dloss.data_ptr()), self._stream()), "pi05_backward")
        if self._dp_group is not None and not self._dp_engine:
            self._allreduce_flat_grads()
        grads = []
        for name, p in self._grad_params:
            dt, o, n, shape = self._offsets[name]
            grads.append(self._flat_grad[dt][o : o + n].view(shape))
        return grads

    def _allreduce_flat_grads(self):
        """The single exchange step of the data-parallel path: SUM all-reduce of each dtype arena over NVLink, then
        the 1/world average DDP applies (train_pytorch.py:440-447).  Two collectives per step (bf16, fp32)."""
        import torch.distributed as dist

        world = dist.get_world_size(self._dp_group)
        for dt in (torch.bfloat16, torch.float32):
            g = self._flat_grad[dt]
            if g is None:
                continue
            if dt == torch.bfloat16:
                # the unused expert lm_head (0.53 GB of never-written gradient) sits last in the bf16 arena: skip it
                g = g[: self._offsets[_UNUSED[0]][1]]
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self._dp_group)
            if self._dp_average == "in_place":
                g.mul_(1.0 / world)

    # ------------------------------------------------------------------ public surface
    def forward(self, observation, actions, noise=None, time=None) -> Tensor:
        """Training forward (pi0_pytorch.py:316-373): returns the un-reduced loss [B, horizon, action_dim] fp32."""
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        # (the reference draws the augmentation parameters first, then noise, then time: same order here)
        images, img_masks, lang_tokens, lang_masks, _state = self._preprocess_observation(
            observation, train=True, rows=self.use_patch_rows, engine_train=need_grad)
        self._ensure_engine(int(actions.shape[0]), train=need_grad, num_images=len(images))
        dev = self._device()
        actions = actions.to(dev, torch.float32).contiguous()
        if noise is None:
            noise = self.sample_noise(actions.shape, actions.device)
        if time is None:
            time = self.sample_time(actions.shape[0], actions.device)
        noise = noise.to(dev, torch.float32).contiguous()
        time = time.to(dev, torch.float32).contiguous()
        pack = self._make_batch(images, img_masks, lang_tokens, lang_masks)
        return self._run_training_forward(pack, actions, noise, time, need_grad)

    def _run_training_forward(self, pack, actions, noise, time, need_grad):
        dev = self._device()
        if not need_grad:
            return self._engine_forward(pack, actions, noise, time)
        self._train_generation += 1  # invalidates the stash any earlier, not yet back-propagated forward relied on
        self._train_key = (dev.index, True, self._engine_key[3], False)
        named = dict(self.named_parameters())
        # autograd inputs = the parameters that can receive a gradient.  The unused lm_head and the parameters the loss
        # cannot reach (self._dead_grad_names) are NOT inputs, so they are unreachable from the loss exactly as in the
        # reference's graph: .grad stays None, and DistributedDataParallel(find_unused_parameters=True) marks them unused.
        self._grad_params = [(n, named[n]) for n in self._offsets
                             if n not in _UNUSED and n not in self._dead_grad_names and named[n].requires_grad]
        if self.direct_grads:
            if self._grad_anchor is None or self._grad_anchor.device != dev:
                self._grad_anchor = torch.zeros(1, device=dev, requires_grad=True)
            return _EngineFunction.apply(self, pack, actions, noise, time, self._grad_anchor)
        return _EngineFunction.apply(self, pack, actions, noise, time, *[p for _, p in self._grad_params])

    @torch.no_grad()
    def sample_actions(self, device, observation, noise=None, num_steps=10, *, prev_action_chunk=None,
                       inference_delay=None, execute_horizon=None, mask_prefix_delay=False,
                       prefix_attention_schedule="exp", max_guidance_weight=0.5, enable_rtc=True) -> Tensor:
        """Inference (pi0_pytorch.py:375-419): prefix pass + KV cache, then `num_steps` Euler steps.

        The keyword arguments after `num_steps` are the real-time-chunking interface of the reference's JAX model
        (src/openpi/models/pi0_rtc.py:234-251; its PyTorch class has none): with `prev_action_chunk` given and `enable_rtc`,
        every step is guided towards the previous chunk through the vector-Jacobian product of the denoiser
        (pi05_denoise_rtc); without it this is exactly the reference PyTorch sampler."""
        use_rtc = bool(enable_rtc) and prev_action_chunk is not None
        images, img_masks, lang_tokens, lang_masks, state = self._preprocess_observation(
            observation, train=False, rows=self.use_patch_rows, engine_train=False, engine_rtc=use_rtc)
        bsize = state.shape[0]
        dev = self._device()
        if noise is None:
            noise = self.sample_noise((bsize, self.ecfg.action_horizon, self.ecfg.action_dim), dev)
        noise = noise.to(dev, torch.float32).contiguous()
        self._ensure_engine(bsize, train=False, num_images=len(images), rtc=use_rtc)
        b, keep = self._make_batch(images, img_masks, lang_tokens, lang_masks)
        taps = bool(getattr(self, "_taps", False))
        if use_rtc:
            return self._sample_actions_rtc(b, keep, noise, int(num_steps), prev_action_chunk, inference_delay,
                                            execute_horizon, bool(mask_prefix_delay), prefix_attention_schedule,
                                            float(max_guidance_weight), taps)
        if self.use_cuda_graph and not taps:
            return self._sample_actions_graphed(b, keep, noise, int(num_steps))
        l = _lib.lib()
        l.pi05_set_taps(self._engine, 1 if taps else 0)
        _lib.check(l.pi05_prefill(self._engine, C.byref(b), self._stream()), "pi05_prefill")
        out = torch.empty_like(noise)
        _lib.check(
            l.pi05_denoise(self._engine, C.c_void_p(noise.data_ptr()), int(num_steps), C.c_void_p(out.data_ptr()), self._stream()),
            "pi05_denoise",
        )
        del keep
        return out

    @staticmethod
    def rtc_prefix_weights(start: int, end: int, total: int, schedule: str) -> Tensor:
        """pi0_rtc.py:47-61 (`get_prefix_weights`): 1 up to `start`, decaying to 0 at `end`, 0 afterwards."""
        start = min(start, end)
        idx = torch.arange(total, dtype=torch.float32)
        if schedule == "ones":
            w = torch.ones(total)
        elif schedule == "zeros":
            w = (idx < start).to(torch.float32)
        elif schedule in ("linear", "exp"):
            w = torch.clamp((start - 1 - idx) / (end - start + 1) + 1, 0, 1)
            if schedule == "exp":
                w = w * torch.expm1(w) / (math.e - 1)
        else:
            raise ValueError(f"Invalid schedule: {schedule}")
        return torch.where(idx >= end, torch.zeros(()), w)

    @staticmethod
    def rtc_guidance_weights(num_steps: int, max_guidance_weight: float) -> list:
        """pi0_rtc.py:341-347 for every step: time is the fp32 running sum 1, 1 + dt, ... (dt = -1 / num_steps, :256)."""
        dt = torch.tensor(-1.0 / num_steps, dtype=torch.float32)
        time = torch.tensor(1.0, dtype=torch.float32)
        mx = torch.tensor(max_guidance_weight, dtype=torch.float32)
        out = []
        for _ in range(num_steps):
            tau = torch.clamp(1.0 - time, 1e-3, 1.0)
            sq = (1 - tau) ** 2
            inv_r2 = (sq + tau**2) / sq
            c = torch.nan_to_num((1 - tau) / tau, posinf=max_guidance_weight)
            gw = c * inv_r2
            out.append(float(torch.where(torch.isnan(gw), gw, torch.minimum(gw, mx))))
            time = time + dt
        return out

    def _sample_actions_rtc(self, b, keep, noise, num_steps, prev_action_chunk, inference_delay, execute_horizon,
                            mask_prefix_delay, schedule, max_guidance_weight, taps):
        H, A = self.ecfg.action_horizon, self.ecfg.action_dim
        dev = self._device()
        exec_h = int(min(max(execute_horizon if execute_horizon is not None else H, 1), H))  # pi0_rtc.py:305-306
        d = int(min(max(0 if inference_delay is None else inference_delay, 0), H))             # :307-308
        prev = torch.as_tensor(prev_action_chunk, dtype=torch.float32, device=dev)
        if prev.dim() == 2:
            prev = prev[None]
        exec_h = min(exec_h, prev.shape[1])                                                    # :313
        provided_before_pad = prev.shape[-1]
        prev = torch.nan_to_num(prev, nan=0.0, posinf=0.0, neginf=0.0)                         # :317
        if prev.shape[-1] > A:                                                                 # :319-324
            prev = prev[..., :A]
        elif prev.shape[-1] < A:
            prev = torch.cat([prev, torch.zeros(*prev.shape[:-1], A - prev.shape[-1], device=dev)], dim=-1)
        if prev.shape[0] != noise.shape[0] or prev.shape[1] != H:
            raise ValueError(f"prev_action_chunk must be [batch, {H}, <= {A}], got {tuple(prev.shape)}")
        prev = prev.contiguous()
        provided = min(14, provided_before_pad, A)                                             # :326
        dim_mask = (torch.arange(A) < provided).to(torch.float32).to(dev)
        weights = self.rtc_prefix_weights(d, exec_h, H, schedule).to(dev).contiguous()         # :337
        gws = self.rtc_guidance_weights(num_steps, max_guidance_weight)
        garr = (C.c_float * num_steps)(*gws)
        l = _lib.lib()
        l.pi05_set_taps(self._engine, 1 if taps else 0)
        _lib.check(l.pi05_prefill(self._engine, C.byref(b), self._stream()), "pi05_prefill")
        out = torch.empty_like(noise)
        _lib.check(
            l.pi05_denoise_rtc(self._engine, C.c_void_p(noise.data_ptr()), num_steps, C.c_void_p(prev.data_ptr()),
                               C.c_void_p(weights.data_ptr()), C.c_void_p(dim_mask.data_ptr()), garr,
                               d if (mask_prefix_delay and provided > 0) else 0,
                               provided if (mask_prefix_delay and provided > 0) else 0,
                               C.c_void_p(out.data_ptr()), self._stream()),
            "pi05_denoise_rtc",
        )
        self._engine_ent["keep_rtc"] = (keep, prev, weights, dim_mask)  # alive until the stream has consumed them
        return out

    def _sample_actions_graphed(self, b, keep, noise, num_steps):
        """The decode path is ~2400 small launches for one observation; replaying them as ONE CUDA graph removes the
        per-launch host cost.  Static device buffers hold the inputs; the graph is captured once per (batch, steps)."""
        imgs, masks, toks, tmask = keep
        is_rows = bool(b.patch_rows)
        key = (b.batch, num_steps, is_rows, int(b.token_len))
        ent = self._graphs.get(key)
        l = _lib.lib()
        if ent is None:
            st = dict(imgs=imgs.clone(), masks=masks.clone(), toks=toks.clone(), tmask=tmask.clone(), noise=noise.clone(),
                      out=torch.empty_like(noise))
            sb = _lib.Batch()
            sb.batch = b.batch
            sb.images = None if is_rows else st["imgs"].data_ptr()
            sb.patch_rows = st["imgs"].data_ptr() if is_rows else None
            sb.image_masks = st["masks"].data_ptr()
            sb.tokens, sb.token_mask = st["toks"].data_ptr(), st["tmask"].data_ptr()
            sb.token_len = b.token_len
            l.pi05_set_taps(self._engine, 0)

            def run():
                _lib.check(l.pi05_prefill(self._engine, C.byref(sb), self._stream()), "pi05_prefill")
                _lib.check(l.pi05_denoise(self._engine, C.c_void_p(st["noise"].data_ptr()), num_steps,
                                          C.c_void_p(st["out"].data_ptr()), self._stream()), "pi05_denoise")

            side = torch.cuda.Stream(device=self._device())
            side.wait_stream(torch.cuda.current_stream(self._device()))
            with torch.cuda.stream(side):
                run()  # eager warm-up: function attributes, tensor-map cache
            torch.cuda.current_stream(self._device()).wait_stream(side)
            torch.cuda.synchronize(self._device())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            ent = (graph, st, sb)
            self._graphs[key] = ent
        graph, st, _ = ent
        st["imgs"].copy_(imgs)
        st["masks"].copy_(masks)
        st["toks"].copy_(toks)
        st["tmask"].copy_(tmask)
        st["noise"].copy_(noise)
        graph.replay()
        return st["out"].clone()


class AdvantageEstimator(PI0Pytorch):
    """B200-native mirror of `AdvantageEstimator` (pi0_pytorch.py:464-644): the pi0.5 backbone over any number of
    camera/timestep images plus a 3-layer tanh value head on suffix_out[:, 0].  Same engine, `cfg.value_head = 1`."""

    _value_head = True
    _apply_aug = False  # "Not applying aug for policy and reward model training" (pi0_pytorch.py:488-489)

    def __init__(self, config, **kw):
        super().__init__(config, **kw)
        self.loss_value_weight = float(_cfg_get(config, "loss_value_weight", 0.0))  # pi0_pytorch.py:467-468
        self.loss_action_weight = float(_cfg_get(config, "loss_action_weight", 1.0))
        self._adv_progress = None
        self._adv_aux = None

    _PART_ORDER = {"base": 0, "left_wrist": 1, "right_wrist": 2}

    def _image_keys(self, images):
        """preprocessing_pytorch.py:196-204: keys `<part>_<timestep>_rgb` sorted by (timestep, part)."""

        def sort_key(k):
            try:
                part, timestep, _ = k.rsplit("_", 2)
                return (int(timestep), self._PART_ORDER[part])
            except (ValueError, KeyError) as exc:
                raise ValueError(f"image key {k!r} is not of the form <base|left_wrist|right_wrist>_<timestep>_rgb") from exc

        return sorted(images.keys(), key=sort_key)

    def _engine_forward(self, batch_pack, actions, noise, time):
        b, keep = batch_pack
        B = b.batch
        dev = self._device()
        loss = torch.empty((B, self.ecfg.action_horizon), dtype=torch.float32, device=dev)
        self._adv_aux = torch.empty(2, dtype=torch.float32, device=dev)
        l = _lib.lib()
        l.pi05_set_taps(self._engine, 1 if getattr(self, "_taps", False) else 0)
        _lib.check(
            l.pi05_forward_advantage(
                self._engine, C.byref(b), C.c_void_p(actions.data_ptr()), C.c_void_p(noise.data_ptr()),
                C.c_void_p(time.data_ptr()), C.c_void_p(self._adv_progress.data_ptr()),
                C.c_float(self.loss_action_weight), C.c_float(self.loss_value_weight),
                C.c_void_p(loss.data_ptr()), C.c_void_p(self._adv_aux.data_ptr()), self._stream(),
            ),
            "pi05_forward_advantage",
        )
        self._engine_ent["keep"] = (keep, actions, noise, time, self._adv_progress)
        return loss

    def forward(self, observation, actions, noise=None, time=None, return_loss_dict=False):
        """pi0_pytorch.py:499-592: loss [B, horizon] = w_action * mean_d (u_t - v_t)^2 + w_value * (value - progress)^2
        (and the reference's loss_aux_dict when asked).  Augmentation is never applied here (:488-489)."""
        progress = getattr(observation, "progress", None)
        if progress is None:
            raise ValueError("AdvantageEstimator.forward needs observation.progress (pi0_pytorch.py:574)")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        images, img_masks, lang_tokens, lang_masks, _state = self._preprocess_observation(
            observation, train=self.training, rows=self.use_patch_rows, engine_train=need_grad)
        dev = self._device()
        actions = actions.to(dev, torch.float32).contiguous()
        if noise is None:
            noise = self.sample_noise(actions.shape, actions.device)
        if time is None:
            time = self.sample_time(actions.shape[0], actions.device)
        noise = noise.to(dev, torch.float32).contiguous()
        time = time.to(dev, torch.float32).contiguous()
        B = actions.shape[0]
        self._adv_progress = progress.to(dev, torch.float32).reshape(B).contiguous()
        self._ensure_engine(B, train=need_grad, num_images=len(images))
        pack = self._make_batch(images, img_masks, lang_tokens, lang_masks)
        loss = self._run_training_forward(pack, actions, noise, time, need_grad)
        if return_loss_dict:
            aux = self._adv_aux
            return loss, {"loss_action": aux[0], "loss_value": aux[1]}
        return loss

    @torch.no_grad()
    def sample_values(self, device, observation) -> Tensor:
        """pi0_pytorch.py:596-644: value (progress) of the current observation, [B, 1] fp32."""
        images, img_masks, lang_tokens, lang_masks, state = self._preprocess_observation(
            observation, train=False, rows=self.use_patch_rows, engine_train=False)
        bsize = state.shape[0]
        dev = self._device()
        noise = self.sample_noise((bsize, self.ecfg.action_horizon, self.ecfg.action_dim), dev).contiguous()
        time = self.sample_time(bsize, dev).contiguous()
        return self._sample_values(images, img_masks, lang_tokens, lang_masks, noise, time)

    def _sample_values(self, images, img_masks, lang_tokens, lang_masks, noise, time):
        bsize = noise.shape[0]
        self._ensure_engine(bsize, train=False, num_images=len(images))
        b, keep = self._make_batch(images, img_masks, lang_tokens, lang_masks)
        out = torch.empty(bsize, dtype=torch.float32, device=self._device())
        l = _lib.lib()
        l.pi05_set_taps(self._engine, 1 if getattr(self, "_taps", False) else 0)
        _lib.check(
            l.pi05_forward_value(self._engine, C.byref(b), C.c_void_p(noise.data_ptr()), C.c_void_p(time.data_ptr()),
                                 C.c_void_p(out.data_ptr()), self._stream()),
            "pi05_forward_value",
        )
        del keep
        return out.view(bsize, 1)


# stage_advantage/annotation/evaluator.py:31 imports this name (its value model: sample_values over whole episodes);
# the stock reference module does not define it, the evaluator's call sites are AdvantageEstimator's surface.
PI0Pytorch_Custom = AdvantageEstimator
