"""CPU oracle for the pi0.5 hot path of OpenDriveLab/kai0 — TEST INFRASTRUCTURE ONLY.

This file is a plain-torch restatement of the reference's PyTorch arithmetic, op for op and rounding point for
rounding point.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it; the product path (kai0_b200/) never does.

PARITY PINNED to the reference itself: the reference ships no test, golden vector or fixture for `models_pytorch`
(SURVEY.md §8c), but its own PI0Pytorch can be executed on the CPU of the build container — the patched transformers
files are loaded in place over the installed transformers and jax is stubbed (tools/reference_loader.py).
tools/make_golden_reference.py ran it in both of its precisions on a seeded configuration with the reference's
geometry (tools/reference_pin.py) and committed the outputs (tests/golden/reference_pin.pt); this oracle reproduces
them to 1.5e-7 relative (loss) / 9e-8 (action chunk) in float32 and to the bf16 noise floor (1.1e-3 / 3.7e-4) under the
bfloat16 dtype map (tests/test_reference_pin_cpu.py, which also re-runs the reference bit-exactly when the checkout is
present).  Additional pins (tests/test_oracle_cpu.py): stock transformers-5.5 `SiglipVisionModel` / `GemmaModel` for
the un-patched math, internal consistency (KV-cache decode == joint forward), mask known-answers.
The same fixture pins the backward pass (gradient of loss.mean() w.r.t. every parameter, reference autograd vs
torch.autograd through this file: 1.1e-6 in float32) and the AdvantageEstimator (6 images, value head, weighted loss,
sample_values, gradients: 4.9e-8 / 5.6e-9 / 1.1e-6 in float32).

Every function cites the reference lines it follows (paths relative to /root/reference/src/openpi/):
  P  = models_pytorch/pi0_pytorch.py
  G  = models_pytorch/gemma_pytorch.py
  MG = models_pytorch/transformers_replace/models/gemma/modeling_gemma.py
  MS = models_pytorch/transformers_replace/models/siglip/modeling_siglip.py
  MP = models_pytorch/transformers_replace/models/paligemma/modeling_paligemma.py
"""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F

MASK_VALUE = -2.3819763e38  # P:159


# ------------------------------------------------------------------------------------------------------------
# configuration (models/gemma.py:58-110, models/pi0_config.py:19-40, G:24-55)
# ------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class GemmaCfg:
    width: int
    depth: int
    mlp_dim: int
    num_heads: int
    num_kv_heads: int
    head_dim: int


GEMMA_2B = GemmaCfg(2048, 18, 16384, 8, 1, 256)
GEMMA_300M = GemmaCfg(1024, 18, 4096, 8, 1, 256)


@dataclasses.dataclass(frozen=True)
class OracleConfig:
    paligemma: GemmaCfg = GEMMA_2B
    expert: GemmaCfg = GEMMA_300M
    vit_width: int = 1152
    vit_depth: int = 27
    vit_mlp_dim: int = 4304
    vit_heads: int = 16
    vit_patch: int = 14
    image_size: int = 224
    vocab_size: int = 257152
    action_dim: int = 32
    action_horizon: int = 50
    max_token_len: int = 200
    num_images: int = 3
    value_head: bool = False
    rms_eps: float = 1e-6
    ln_eps: float = 1e-6
    rope_theta: float = 10000.0

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.vit_patch) ** 2


def tiny_config(**kw) -> OracleConfig:
    """A shrunken architecture with every structural feature of the real one (GQA 8:1, adaRMS expert, ViT with a
    head_dim that is not a multiple of 16, ragged token counts) for tests that must run in seconds."""
    base = dict(
        paligemma=GemmaCfg(256, 2, 512, 8, 1, 32),
        expert=GemmaCfg(128, 2, 256, 8, 1, 32),
        vit_width=144,
        vit_depth=2,
        vit_mlp_dim=272,
        vit_heads=2,
        vit_patch=14,
        image_size=56,
        vocab_size=512,
        action_dim=32,
        action_horizon=10,
        max_token_len=24,
        num_images=3,
    )
    base.update(kw)
    return OracleConfig(**base)


# ------------------------------------------------------------------------------------------------------------
# parameters: reference state_dict names, reference dtype map (G:63-83)
# ------------------------------------------------------------------------------------------------------------
_PWE = "paligemma_with_expert."
_VT = _PWE + "paligemma.model.vision_tower.vision_model."
_LM = _PWE + "paligemma.model.language_model."
_EX = _PWE + "gemma_expert.model."
_KEEP_F32 = (
    "vision_tower.vision_model.embeddings.patch_embedding.weight",
    "vision_tower.vision_model.embeddings.patch_embedding.bias",
    "vision_tower.vision_model.embeddings.position_embedding.weight",
    "input_layernorm",
    "post_attention_layernorm",
    "model.norm",
)


def param_specs(cfg: OracleConfig) -> dict[str, tuple[tuple[int, ...], torch.dtype]]:
    """name -> (shape, dtype) for every tensor the hot path reads (the unused `lm_head`s are listed by
    kai0_b200.pi0_pytorch, which owns the module tree; the oracle does not need them)."""
    s: dict[str, tuple[tuple[int, ...], torch.dtype]] = {}

    def add(name, shape):
        full = name
        dt = torch.bfloat16 if full.startswith(_PWE) else torch.float32
        if full.startswith(_PWE) and any(k in full for k in _KEEP_F32):
            dt = torch.float32
        s[full] = (tuple(shape), dt)

    W, P = cfg.vit_width, cfg.vit_patch
    add(_VT + "embeddings.patch_embedding.weight", (W, 3, P, P))
    add(_VT + "embeddings.patch_embedding.bias", (W,))
    add(_VT + "embeddings.position_embedding.weight", (cfg.num_patches, W))
    for i in range(cfg.vit_depth):
        L = f"{_VT}encoder.layers.{i}."
        for ln in ("layer_norm1", "layer_norm2"):
            add(L + ln + ".weight", (W,))
            add(L + ln + ".bias", (W,))
        for pj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            add(L + f"self_attn.{pj}.weight", (W, W))
            add(L + f"self_attn.{pj}.bias", (W,))
        add(L + "mlp.fc1.weight", (cfg.vit_mlp_dim, W))
        add(L + "mlp.fc1.bias", (cfg.vit_mlp_dim,))
        add(L + "mlp.fc2.weight", (W, cfg.vit_mlp_dim))
        add(L + "mlp.fc2.bias", (W,))
    add(_VT + "post_layernorm.weight", (W,))
    add(_VT + "post_layernorm.bias", (W,))
    D = cfg.paligemma.width
    add(_PWE + "paligemma.model.multi_modal_projector.linear.weight", (D, W))
    add(_PWE + "paligemma.model.multi_modal_projector.linear.bias", (D,))
    add(_LM + "embed_tokens.weight", (cfg.vocab_size, D))
    for prefix, g, ada in ((_LM, cfg.paligemma, False), (_EX, cfg.expert, True)):
        for i in range(g.depth):
            L = f"{prefix}layers.{i}."
            add(L + "self_attn.q_proj.weight", (g.num_heads * g.head_dim, g.width))
            add(L + "self_attn.k_proj.weight", (g.num_kv_heads * g.head_dim, g.width))
            add(L + "self_attn.v_proj.weight", (g.num_kv_heads * g.head_dim, g.width))
            add(L + "self_attn.o_proj.weight", (g.width, g.num_heads * g.head_dim))
            add(L + "mlp.gate_proj.weight", (g.mlp_dim, g.width))
            add(L + "mlp.up_proj.weight", (g.mlp_dim, g.width))
            add(L + "mlp.down_proj.weight", (g.width, g.mlp_dim))
            for nm in ("input_layernorm", "post_attention_layernorm"):
                if ada:
                    add(L + nm + ".dense.weight", (3 * g.width, g.width))
                    add(L + nm + ".dense.bias", (3 * g.width,))
                else:
                    add(L + nm + ".weight", (g.width,))
        if ada:
            add(prefix + "norm.dense.weight", (3 * g.width, g.width))
            add(prefix + "norm.dense.bias", (3 * g.width,))
        else:
            add(prefix + "norm.weight", (g.width,))
    E = cfg.expert.width
    add("action_in_proj.weight", (E, cfg.action_dim))
    add("action_in_proj.bias", (E,))
    add("action_out_proj.weight", (cfg.action_dim, E))
    add("action_out_proj.bias", (cfg.action_dim,))
    add("time_mlp_in.weight", (E, E))
    add("time_mlp_in.bias", (E,))
    add("time_mlp_out.weight", (E, E))
    add("time_mlp_out.bias", (E,))
    if cfg.value_head:  # P:473-481 (nn.Sequential indices 0, 2, 4)
        add("value_head.0.weight", (E, E))
        add("value_head.0.bias", (E,))
        add("value_head.2.weight", (E, E))
        add("value_head.2.bias", (E,))
        add("value_head.4.weight", (1, E))
        add("value_head.4.bias", (1,))
    return s


def init_params(cfg: OracleConfig, seed: int = 0, *, exercise_all: bool = True) -> dict[str, torch.Tensor]:
    """Seeded synthetic weights (no checkpoint is reachable).  With exercise_all=True the tensors the reference
    zero-initialises (RMSNorm weight, adaRMS dense: MG:59-63) get non-zero values so those paths are exercised."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, (shape, dt) in param_specs(cfg).items():
        if name.endswith("layer_norm1.weight") or name.endswith("layer_norm2.weight") or name.endswith(
            "post_layernorm.weight"
        ):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "layernorm.weight" in name or name.endswith("model.norm.weight") or name.endswith(
            "language_model.norm.weight"
        ):
            t = 0.1 * torch.randn(shape, generator=g) if exercise_all else torch.zeros(shape)
        elif "dense.weight" in name:
            t = 0.02 * torch.randn(shape, generator=g) if exercise_all else torch.zeros(shape)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name.endswith("position_embedding.weight"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name.endswith("embed_tokens.weight"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        out[name] = t.to(dt)
    return out


# ------------------------------------------------------------------------------------------------------------
# small pieces
# ------------------------------------------------------------------------------------------------------------
def create_sinusoidal_pos_embedding(time: torch.Tensor, dimension: int, min_period: float, max_period: float):
    """P:25-42 — computed in float64 (get_safe_dtype keeps float64 on CPU too, P:14-22)."""
    if dimension % 2 != 0:
        raise ValueError(f"dimension ({dimension}) must be divisible by 2")
    if time.ndim != 1:
        raise ValueError("The time tensor is expected to be of shape `(batch_size, )`.")
    fraction = torch.linspace(0.0, 1.0, dimension // 2, dtype=torch.float64, device=time.device)
    period = min_period * (max_period / min_period) ** fraction
    scaling_factor = 1.0 / period * 2 * math.pi
    sin_input = scaling_factor[None, :] * time[:, None]
    return torch.cat([torch.sin(sin_input), torch.cos(sin_input)], dim=1)


def make_att_2d_masks(pad_masks, att_masks):
    """P:52-81."""
    if att_masks.ndim != 2:
        raise ValueError(att_masks.ndim)
    if pad_masks.ndim != 2:
        raise ValueError(pad_masks.ndim)
    cumsum = torch.cumsum(att_masks, dim=1)
    att_2d_masks = cumsum[:, None, :] <= cumsum[:, :, None]
    pad_2d_masks = pad_masks[:, None, :] * pad_masks[:, :, None]
    return att_2d_masks & pad_2d_masks


def prepare_attention_masks_4d(att_2d_masks):
    """P:156-159."""
    return torch.where(att_2d_masks[:, None, :, :], 0.0, MASK_VALUE)


def gemma_rmsnorm(x, weight=None, dense_w=None, dense_b=None, cond=None, eps=1e-6):
    """MG:49-104 (GemmaRMSNorm.forward): returns (normed, gate)."""
    dtype = x.dtype
    var = torch.mean(torch.square(x.float()), dim=-1, keepdim=True)  # MG:68
    normed = x * torch.rsqrt(var + eps)  # MG:70 (promotes to fp32)
    if cond is None or dense_w is None:
        normed = normed * (1.0 + weight.float())  # MG:80
        return normed.to(dtype), None
    modulation = F.linear(cond, dense_w, dense_b)  # MG:88
    if x.dim() == 3:
        modulation = modulation.unsqueeze(1)
    scale, shift, gate = torch.chunk(modulation, 3, dim=-1)  # MG:93
    normed = normed * (1 + scale.to(torch.float32)) + shift.to(torch.float32)  # MG:102
    return normed.to(dtype), gate.to(dtype)


def gated_residual(x, y, gate):
    """MG:209-227."""
    if gate is None:
        return x + y
    return x + y * gate


def rope_cos_sin(position_ids, head_dim, theta, dtype):
    """MG:129-160 (GemmaRotaryEmbedding, rope_type default): fp32 angles, cos/sin cast to `dtype`."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(dtype=torch.float) / head_dim))
    inv_freq_expanded = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
    position_ids_expanded = position_ids[:, None, :].float()
    freqs = (inv_freq_expanded.float() @ position_ids_expanded.float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    """MG:163-167."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2 :]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, unsqueeze_dim=1):
    """MG:170-194."""
    cos = cos.unsqueeze(unsqueeze_dim)
    sin = sin.unsqueeze(unsqueeze_dim)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def repeat_kv(h, n_rep):
    """MG:197-206."""
    b, kvh, s, d = h.shape
    if n_rep == 1:
        return h
    return h[:, :, None, :, :].expand(b, kvh, n_rep, s, d).reshape(b, kvh * n_rep, s, d)


def eager_attention(query, key, value, attention_mask, scaling, n_rep):
    """MG:230-253 (Gemma) and MS:325-345 (SigLIP, n_rep = 1, mask None)."""
    key_states = repeat_kv(key, n_rep)
    value_states = repeat_kv(value, n_rep)
    attn_weights = torch.matmul(query, key_states.transpose(2, 3)) * scaling
    if attention_mask is not None:
        attn_weights = attn_weights + attention_mask[:, :, :, : key_states.shape[-2]]
    attn_weights = F.softmax(attn_weights, dim=-1, dtype=torch.float32).to(query.dtype)
    attn_output = torch.matmul(attn_weights, value_states)
    return attn_output.transpose(1, 2).contiguous()


def gemma_mlp(x, wg, wu, wd):
    """MG:113-126 with act = gelu_pytorch_tanh (G:33,51)."""
    return F.linear(F.gelu(F.linear(x, wg), approximate="tanh") * F.linear(x, wu), wd)


# ------------------------------------------------------------------------------------------------------------
# SigLIP vision tower + projector
# ------------------------------------------------------------------------------------------------------------
def siglip_embed_image(p, cfg: OracleConfig, img: torch.Tensor, taps=None, tap_prefix=""):
    """G:85-86 -> MP:232-247 -> MS:763-796.  img: [B,3,H,W] fp32 in [-1,1] -> [B, num_patches, D] bf16."""
    W = cfg.vit_width
    pw = p[_VT + "embeddings.patch_embedding.weight"]
    x = F.conv2d(img.to(pw.dtype), pw, p[_VT + "embeddings.patch_embedding.bias"], stride=cfg.vit_patch)  # MS:273-274
    x = x.flatten(2).transpose(1, 2)  # MS:275
    x = x + p[_VT + "embeddings.position_embedding.weight"][None]  # MS:280
    qdt = p[_VT + "encoder.layers.0.self_attn.q_proj.weight"].dtype
    if qdt == torch.bfloat16:
        x = x.to(torch.bfloat16)  # MS:777-778
    if taps is not None:
        taps[tap_prefix + "vit_embed"] = x
    H = cfg.vit_heads
    hd = W // H
    scale = hd**-0.5
    B, T, _ = x.shape
    for i in range(cfg.vit_depth):
        L = f"{_VT}encoder.layers.{i}."
        res = x
        h = F.layer_norm(x, (W,), p[L + "layer_norm1.weight"], p[L + "layer_norm1.bias"], cfg.ln_eps)  # MS:466
        q = F.linear(h, p[L + "self_attn.q_proj.weight"], p[L + "self_attn.q_proj.bias"])
        k = F.linear(h, p[L + "self_attn.k_proj.weight"], p[L + "self_attn.k_proj.bias"])
        v = F.linear(h, p[L + "self_attn.v_proj.weight"], p[L + "self_attn.v_proj.bias"])
        q = q.view(B, T, H, hd).transpose(1, 2)
        k = k.view(B, T, H, hd).transpose(1, 2)
        v = v.view(B, T, H, hd).transpose(1, 2)
        a = eager_attention(q, k, v, None, scale, 1).reshape(B, T, W)  # MS:325-345,399-411
        a = F.linear(a, p[L + "self_attn.out_proj.weight"], p[L + "self_attn.out_proj.bias"])
        x = res + a  # MS:472
        res = x
        h = F.layer_norm(x, (W,), p[L + "layer_norm2.weight"], p[L + "layer_norm2.bias"], cfg.ln_eps)
        h = F.linear(h, p[L + "mlp.fc1.weight"], p[L + "mlp.fc1.bias"])
        h = F.gelu(h, approximate="tanh")  # gelu_pytorch_tanh (HF SiglipVisionConfig default)
        h = F.linear(h, p[L + "mlp.fc2.weight"], p[L + "mlp.fc2.bias"])
        x = res + h  # MS:477
        if taps is not None:
            taps[f"{tap_prefix}vit_layer{i}"] = x
    x = F.layer_norm(x, (W,), p[_VT + "post_layernorm.weight"], p[_VT + "post_layernorm.bias"], cfg.ln_eps)  # MS:787
    x = F.linear(
        x,
        p[_PWE + "paligemma.model.multi_modal_projector.linear.weight"],
        p[_PWE + "paligemma.model.multi_modal_projector.linear.bias"],
    )  # MP:96-99 (no 1/sqrt(d) rescale, MP:244-247)
    return x


# ------------------------------------------------------------------------------------------------------------
# prefix / suffix embedding
# ------------------------------------------------------------------------------------------------------------
def embed_prefix(p, cfg, images, img_masks, lang_tokens, lang_masks, taps=None):
    """P:186-235."""
    embs, pad_masks, att_masks = [], [], []
    for n, (img, img_mask) in enumerate(zip(images, img_masks, strict=True)):
        img_emb = siglip_embed_image(p, cfg, img, taps, f"img{n}_")
        bsize, num_img_embs = img_emb.shape[:2]
        embs.append(img_emb)
        pad_masks.append(img_mask[:, None].expand(bsize, num_img_embs))
        att_masks += [0] * num_img_embs
    lang_emb = F.embedding(lang_tokens, p[_LM + "embed_tokens.weight"])  # G:88-89
    lang_emb = lang_emb * math.sqrt(lang_emb.shape[-1])  # P:213-216
    embs.append(lang_emb)
    pad_masks.append(lang_masks)
    att_masks += [0] * lang_emb.shape[1]
    embs = torch.cat(embs, dim=1)
    pad_masks = torch.cat(pad_masks, dim=1)
    att_masks = torch.tensor(att_masks, dtype=torch.bool, device=pad_masks.device)
    att_masks = att_masks[None, :].expand(pad_masks.shape[0], len(att_masks))
    return embs, pad_masks, att_masks


def embed_suffix(p, cfg, noisy_actions, timestep):
    """P:237-314, pi05 branch (no state token)."""
    E = cfg.expert.width
    time_emb = create_sinusoidal_pos_embedding(timestep, E, min_period=4e-3, max_period=4.0)
    time_emb = time_emb.type(dtype=timestep.dtype)  # P:267
    action_emb = F.linear(noisy_actions, p["action_in_proj.weight"], p["action_in_proj.bias"])  # P:270-273
    x = F.linear(time_emb, p["time_mlp_in.weight"], p["time_mlp_in.bias"])  # P:289-293
    x = F.silu(x)
    x = F.linear(x, p["time_mlp_out.weight"], p["time_mlp_out.bias"])
    adarms_cond = F.silu(x)
    bsize, n = action_emb.shape[:2]
    pad_masks = torch.ones(bsize, n, dtype=torch.bool, device=timestep.device)
    att = [1] + [0] * (cfg.action_horizon - 1)  # P:307
    att_masks = torch.tensor(att, dtype=action_emb.dtype, device=action_emb.device)[None, :].expand(bsize, len(att))
    return action_emb, pad_masks, att_masks, adarms_cond


# ------------------------------------------------------------------------------------------------------------
# the joint transformer (G:126-279) and the single-stream branches (G:102-125 -> MG:446-555)
# ------------------------------------------------------------------------------------------------------------
def _norm(p, prefix, name, x, cond, ada, eps):
    if ada:
        return gemma_rmsnorm(x, None, p[prefix + name + ".dense.weight"], p[prefix + name + ".dense.bias"], cond, eps)
    return gemma_rmsnorm(x, p[prefix + name + ".weight"], eps=eps)


def joint_forward(p, cfg, prefix_embs, suffix_embs, attention_mask, position_ids, adarms_cond, taps=None):
    """G:158-238 for every layer, then the final norms G:262-275.  Returns (prefix_out, suffix_out)."""
    streams = [(_LM, cfg.paligemma, False), (_EX, cfg.expert, True)]
    xs = [prefix_embs, suffix_embs]
    conds = [None, adarms_cond]
    hd = cfg.paligemma.head_dim
    for li in range(cfg.paligemma.depth):
        qs, ks, vs, gates = [], [], [], []
        for (pre, g, ada), x, c in zip(streams, xs, conds):
            L = f"{pre}layers.{li}."
            h, gate = _norm(p, L, "input_layernorm", x, c, ada, cfg.rms_eps)  # G:167
            gates.append(gate)
            shp = (*h.shape[:-1], -1, g.head_dim)
            qs.append(F.linear(h, p[L + "self_attn.q_proj.weight"]).view(shp).transpose(1, 2))  # G:172-174
            ks.append(F.linear(h, p[L + "self_attn.k_proj.weight"]).view(shp).transpose(1, 2))
            vs.append(F.linear(h, p[L + "self_attn.v_proj.weight"]).view(shp).transpose(1, 2))
        q = torch.cat(qs, dim=2)
        k = torch.cat(ks, dim=2)
        v = torch.cat(vs, dim=2)
        cos, sin = rope_cos_sin(position_ids, hd, cfg.rope_theta, q.dtype)  # G:185-192
        q, k = apply_rotary_pos_emb(q, k, cos, sin)  # G:193-195
        n_rep = cfg.paligemma.num_heads // cfg.paligemma.num_kv_heads
        att = eager_attention(q, k, v, attention_mask, hd**-0.5, n_rep)  # G:201-208
        att = att.reshape(q.shape[0], -1, cfg.paligemma.num_heads * hd)  # G:211
        if taps is not None:
            taps[f"layer{li}_att"] = att
        outs = []
        start = 0
        for (pre, g, ada), x, c, gate in zip(streams, xs, conds, gates):
            L = f"{pre}layers.{li}."
            end = start + x.shape[1]
            wo = p[L + "self_attn.o_proj.weight"]
            a = att if att.dtype == wo.dtype else att.to(wo.dtype)  # G:220-221
            o = F.linear(a[:, start:end], wo)  # G:222
            o = gated_residual(x, o, gate)  # G:225
            after_first = o.clone()
            o, gate2 = _norm(p, L, "post_attention_layernorm", o, c, ada, cfg.rms_eps)  # G:227
            if p[L + "mlp.up_proj.weight"].dtype == torch.bfloat16:
                o = o.to(torch.bfloat16)  # G:229-230
            o = gemma_mlp(o, p[L + "mlp.gate_proj.weight"], p[L + "mlp.up_proj.weight"], p[L + "mlp.down_proj.weight"])
            o = gated_residual(after_first, o, gate2)  # G:234
            outs.append(o)
            start = end
        xs = outs
        if taps is not None:
            taps[f"layer{li}_prefix"] = xs[0]
            taps[f"layer{li}_suffix"] = xs[1]
    outs = []
    for (pre, g, ada), x, c in zip(streams, xs, conds):
        o, _ = _norm(p, pre, "norm", x, c, ada, cfg.rms_eps)  # G:262-267
        outs.append(o)
    return outs[0], outs[1]


def single_stream_forward(p, cfg, which, inputs_embeds, attention_mask, position_ids, past_kv=None, use_cache=False,
                          adarms_cond=None):
    """GemmaModel.forward MG:446-555 with GemmaDecoderLayer MG:344-384 and GemmaAttention MG:282-329.
    which = "prefix" (PaliGemma LM, fills the cache: G:102-113) or "suffix" (expert, reads the cache: G:114-125).
    Returns (last_hidden_state, cache) where cache is a list of (K, V) per layer (post-RoPE K: MG:303-307)."""
    pre, g, ada = (_LM, cfg.paligemma, False) if which == "prefix" else (_EX, cfg.expert, True)
    h = inputs_embeds
    if p[pre + "layers.0.self_attn.q_proj.weight"].dtype == torch.bfloat16:
        h = h.to(torch.bfloat16)  # MG:506-507
    cos, sin = rope_cos_sin(position_ids, g.head_dim, cfg.rope_theta, h.dtype)  # MG:510
    cache = []
    n_rep = g.num_heads // g.num_kv_heads
    for li in range(g.depth):
        L = f"{pre}layers.{li}."
        res = h
        x, gate = _norm(p, L, "input_layernorm", h, adarms_cond, ada, cfg.rms_eps)  # MG:358
        shp = (*x.shape[:-1], -1, g.head_dim)
        q = F.linear(x, p[L + "self_attn.q_proj.weight"]).view(shp).transpose(1, 2)
        k = F.linear(x, p[L + "self_attn.k_proj.weight"]).view(shp).transpose(1, 2)
        v = F.linear(x, p[L + "self_attn.v_proj.weight"]).view(shp).transpose(1, 2)
        q, k = apply_rotary_pos_emb(q, k, cos, sin)  # MG:300
        if past_kv is not None and not use_cache:
            k = torch.cat([past_kv[li][0], k], dim=2)  # MG:309-310
            v = torch.cat([past_kv[li][1], v], dim=2)
        if use_cache:
            cache.append((k, v))
        a = eager_attention(q, k, v, attention_mask, g.head_dim**-0.5, n_rep)
        a = a.reshape(*x.shape[:-1], -1).contiguous()  # MG:327
        a = F.linear(a, p[L + "self_attn.o_proj.weight"])  # MG:328
        h = gated_residual(res, a, gate)  # MG:372
        res = h
        x, gate = _norm(p, L, "post_attention_layernorm", h, adarms_cond, ada, cfg.rms_eps)  # MG:376
        x = gemma_mlp(x, p[L + "mlp.gate_proj.weight"], p[L + "mlp.up_proj.weight"], p[L + "mlp.down_proj.weight"])
        h = gated_residual(res, x, gate)  # MG:378
    h, _ = _norm(p, pre, "norm", h, adarms_cond, ada, cfg.rms_eps)  # MG:544
    return h, cache


# ------------------------------------------------------------------------------------------------------------
# model surface
# ------------------------------------------------------------------------------------------------------------
def _is_bf16(p):
    return p[_LM + "layers.0.self_attn.q_proj.weight"].dtype == torch.bfloat16


def model_v_t(p, cfg: OracleConfig, images, img_masks, lang_tokens, lang_masks, x_t, time, taps=None):
    """The network part of P:316-373: (obs, x_t, t) -> v_t [B,H,A] fp32 (and suffix_out for the value head)."""
    prefix_embs, prefix_pad, prefix_att = embed_prefix(p, cfg, images, img_masks, lang_tokens, lang_masks, taps)
    suffix_embs, suffix_pad, suffix_att, cond = embed_suffix(p, cfg, x_t, time)
    if _is_bf16(p):
        suffix_embs = suffix_embs.to(torch.bfloat16)  # P:332-337
        prefix_embs = prefix_embs.to(torch.bfloat16)
    if taps is not None:
        taps["prefix_embs"] = prefix_embs
        taps["suffix_embs"] = suffix_embs
        taps["adarms_cond"] = cond
    pad_masks = torch.cat([prefix_pad, suffix_pad], dim=1)
    att_masks = torch.cat([prefix_att, suffix_att], dim=1)
    att_2d = make_att_2d_masks(pad_masks, att_masks)  # P:342
    position_ids = torch.cumsum(pad_masks, dim=1) - 1  # P:343
    mask4d = prepare_attention_masks_4d(att_2d)  # P:346
    if taps is not None:  # the index work of P:219-235,342-346, for bit-exact comparison
        taps["pad_masks"] = pad_masks
        taps["position_ids"] = position_ids
        taps["att_2d_masks"] = att_2d
    prefix_out, suffix_out = joint_forward(p, cfg, prefix_embs, suffix_embs, mask4d, position_ids, cond, taps)
    if taps is not None:
        taps["prefix_out"] = prefix_out
        taps["suffix_out"] = suffix_out
    so = suffix_out[:, -cfg.action_horizon :].to(torch.float32)  # P:364-365
    v_t = F.linear(so, p["action_out_proj.weight"], p["action_out_proj.bias"])  # P:368-371
    return v_t, suffix_out


def forward_loss(p, cfg: OracleConfig, images, img_masks, lang_tokens, lang_masks, actions, noise, time, taps=None):
    """PI0Pytorch.forward, P:316-373, with noise/time injected (RNG parity is out of scope).
    Preprocessing (P:318) is the identity for 224x224 inputs with train-time augmentation disabled."""
    t = time[:, None, None]
    x_t = t * noise + (1 - t) * actions  # P:326-327
    u_t = noise - actions  # P:328
    v_t, _ = model_v_t(p, cfg, images, img_masks, lang_tokens, lang_masks, x_t, time, taps)
    if taps is not None:
        taps["v_t"] = v_t
    return F.mse_loss(u_t, v_t, reduction="none")  # P:373


def value_head(p, suffix_out):
    """P:571-572, 640-642: tanh(MLP3(suffix_out[:, 0]))."""
    x = suffix_out[:, 0, :].to(torch.float32)
    x = F.silu(F.linear(x, p["value_head.0.weight"], p["value_head.0.bias"]))
    x = F.silu(F.linear(x, p["value_head.2.weight"], p["value_head.2.bias"]))
    return torch.tanh(F.linear(x, p["value_head.4.weight"], p["value_head.4.bias"]))


def advantage_forward_loss(p, cfg, images, img_masks, lang_tokens, lang_masks, actions, noise, time, progress,
                           loss_action_weight=1.0, loss_value_weight=0.0):
    """AdvantageEstimator.forward, P:499-592. Returns loss [B, H]."""
    t = time[:, None, None]
    x_t = t * noise + (1 - t) * actions
    u_t = noise - actions
    v_t, suffix_out = model_v_t(p, cfg, images, img_masks, lang_tokens, lang_masks, x_t, time)
    loss_action = F.mse_loss(u_t, v_t, reduction="none").mean(dim=-1)  # P:564
    loss = loss_action * loss_action_weight
    value_pred = value_head(p, suffix_out)  # P:571-572
    tgt = torch.clamp(progress.float(), -1.0, 1.0).unsqueeze(1)  # P:574-576
    value_loss = F.mse_loss(value_pred, tgt, reduction="none").to(loss.dtype) * loss_value_weight
    return loss + value_loss  # P:587


def prefill(p, cfg, images, img_masks, lang_tokens, lang_masks):
    """P:383-399: prefix pass that fills the KV cache. Returns (prefix_pad_masks, cache)."""
    prefix_embs, prefix_pad, prefix_att = embed_prefix(p, cfg, images, img_masks, lang_tokens, lang_masks)
    att_2d = make_att_2d_masks(prefix_pad, prefix_att)
    position_ids = torch.cumsum(prefix_pad, dim=1) - 1
    mask4d = prepare_attention_masks_4d(att_2d)
    _, cache = single_stream_forward(p, cfg, "prefix", prefix_embs, mask4d, position_ids, use_cache=True)
    return prefix_pad, cache


def denoise_step(p, cfg, prefix_pad_masks, cache, x_t, timestep):
    """P:421-461."""
    suffix_embs, suffix_pad, suffix_att, cond = embed_suffix(p, cfg, x_t, timestep)
    suffix_len = suffix_pad.shape[1]
    bsz, prefix_len = prefix_pad_masks.shape
    prefix_pad_2d = prefix_pad_masks[:, None, :].expand(bsz, suffix_len, prefix_len)
    suffix_att_2d = make_att_2d_masks(suffix_pad, suffix_att)
    full = torch.cat([prefix_pad_2d, suffix_att_2d], dim=2)
    prefix_offsets = torch.sum(prefix_pad_masks, dim=-1)[:, None]
    position_ids = prefix_offsets + torch.cumsum(suffix_pad, dim=1) - 1
    mask4d = prepare_attention_masks_4d(full)
    out, _ = single_stream_forward(p, cfg, "suffix", suffix_embs, mask4d, position_ids, past_kv=cache,
                                   use_cache=False, adarms_cond=cond)
    so = out[:, -cfg.action_horizon :].to(torch.float32)
    return F.linear(so, p["action_out_proj.weight"], p["action_out_proj.bias"])


@torch.no_grad()
def sample_actions(p, cfg, images, img_masks, lang_tokens, lang_masks, noise, num_steps=10):
    """PI0Pytorch.sample_actions, P:375-419 (time is an fp32 running sum; loop while time >= -dt/2)."""
    bsize = noise.shape[0]
    prefix_pad, cache = prefill(p, cfg, images, img_masks, lang_tokens, lang_masks)
    dt = torch.tensor(-1.0 / num_steps, dtype=torch.float32)
    x_t = noise
    time = torch.tensor(1.0, dtype=torch.float32)
    while time >= -dt / 2:
        v_t = denoise_step(p, cfg, prefix_pad, cache, x_t, time.expand(bsize))
        x_t = x_t + dt * v_t
        time = time + dt
    return x_t


def decode_times(num_steps=10) -> list[float]:
    """The fp32 running-sum timestep sequence of P:401-418 (1.0, 0.9, 0.8, 0.70000005, ...)."""
    dt = torch.tensor(-1.0 / num_steps, dtype=torch.float32)
    time = torch.tensor(1.0, dtype=torch.float32)
    out = []
    while time >= -dt / 2:
        out.append(float(time))
        time = time + dt
    return out


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------------
def synthetic_batch(cfg: OracleConfig, batch: int, seed: int = 1234, valid_tokens: int | None = None,
                    ragged: bool = False):
    """uint8 images -> fp32 NCHW in [-1,1] exactly as Observation.from_dict (models/model.py:129-133)."""
    g = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    images = []
    for _ in range(cfg.num_images):
        u8 = torch.randint(0, 256, (batch, S, S, 3), generator=g, dtype=torch.uint8)
        images.append(u8.to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0)
    img_masks = [torch.ones(batch, dtype=torch.bool) for _ in range(cfg.num_images)]
    L = cfg.max_token_len
    nv = valid_tokens if valid_tokens is not None else max(1, int(L * 0.48))
    tokens = torch.zeros(batch, L, dtype=torch.int64)
    mask = torch.zeros(batch, L, dtype=torch.bool)
    for b in range(batch):
        n = nv if not ragged else max(1, nv - 3 * b)
        tokens[b, :n] = torch.randint(0, cfg.vocab_size, (n,), generator=g)
        mask[b, :n] = True
    actions = torch.randn(batch, cfg.action_horizon, cfg.action_dim, generator=g)
    actions[..., 14:] = 0.0
    noise = torch.randn(batch, cfg.action_horizon, cfg.action_dim, generator=torch.Generator().manual_seed(4321))
    beta = torch.distributions.Beta(torch.tensor(1.5), torch.tensor(1.0))
    torch.manual_seed(4321)
    time = (beta.sample((batch,)) * 0.999 + 0.001).to(torch.float32)
    return dict(images=images, img_masks=img_masks, tokens=tokens, token_mask=mask, actions=actions, noise=noise,
                time=time)
