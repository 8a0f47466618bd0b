"""CPU tests that PIN the oracle (the reference ships no test or golden vector for models_pytorch: SURVEY.md §8c):
  * against stock transformers `SiglipVisionModel` / `GemmaModel` for the un-patched arithmetic,
  * through internal consistency (KV-cache decode == joint forward on the suffix rows),
  * against the reference's documented mask semantics (pi0_pytorch.py:52-81 docstring examples),
  * against the committed golden fixtures (tests/golden, produced by tools/make_golden.py).
"""
import os

import pytest
import torch

import helpers as H
from oracle import pi05_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _fp32(params):
    return {k: v.float() for k, v in params.items()}


def test_att_2d_masks_known_answers():
    # [[1 1 1 1 1 1]] -> pure causal
    pad = torch.ones(1, 6, dtype=torch.bool)
    m = O.make_att_2d_masks(pad, torch.ones(1, 6, dtype=torch.int64))
    assert torch.equal(m[0], torch.tril(torch.ones(6, 6, dtype=torch.bool)))
    # [[0 0 0 1 1 1]] -> prefix-LM: first 3 see each other, last 3 causal over everything before
    m = O.make_att_2d_masks(pad, torch.tensor([[0, 0, 0, 1, 1, 1]]))
    exp = torch.tensor([[1, 1, 1, 0, 0, 0]] * 3 + [[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 0], [1, 1, 1, 1, 1, 1]],
                       dtype=torch.bool)
    assert torch.equal(m[0], exp)
    # padding removes rows and columns
    pad2 = torch.tensor([[1, 1, 0, 1]], dtype=torch.bool)
    m = O.make_att_2d_masks(pad2, torch.zeros(1, 4, dtype=torch.int64))
    assert not m[0, 2].any() and not m[0, :, 2].any() and m[0, 0, 3]
    with pytest.raises(ValueError):
        O.make_att_2d_masks(pad[0], torch.ones(1, 6))


def test_pi05_block_mask_structure():
    """prefix x prefix (minus padded columns) and suffix x everything: the structure the engine hard-codes."""
    oc = O.tiny_config()
    B, P, A = 2, oc.num_images * oc.num_patches + oc.max_token_len, oc.action_horizon
    pad = torch.ones(B, P + A, dtype=torch.bool)
    pad[1, P - 5:P] = False
    att = torch.cat([torch.zeros(B, P), torch.tensor([1.0] + [0.0] * (A - 1)).expand(B, A)], dim=1)
    m = O.make_att_2d_masks(pad, att)
    assert not m[:, :P, P:].any()  # prefix never sees the suffix
    assert m[0, P:, :].all()  # suffix sees all valid keys
    assert not m[1, P:, P - 5:P].any() and m[1, P:, :P - 5].all()
    pos = torch.cumsum(pad, dim=1) - 1
    assert pos[1, P] == P - 5  # suffix positions start at the number of valid prefix tokens


def test_decode_time_sequence():
    ts = O.decode_times(10)
    assert len(ts) == 10 and ts[0] == 1.0
    t = torch.tensor(1.0)
    dt = torch.tensor(-0.1)
    for v in ts:
        assert v == float(t)
        t = t + dt


def test_time_embedding_matches_closed_form():
    t = torch.tensor([0.25, 1.0], dtype=torch.float32)
    e = O.create_sinusoidal_pos_embedding(t, 8, 4e-3, 4.0)
    assert e.dtype == torch.float64 and e.shape == (2, 8)
    import math

    period = 4e-3 * (4.0 / 4e-3) ** (torch.arange(4, dtype=torch.float64) / 3)
    ref = torch.sin(2 * math.pi * t.double()[:, None] / period[None])
    assert torch.allclose(e[:, :4], ref, atol=1e-12)
    with pytest.raises(ValueError):
        O.create_sinusoidal_pos_embedding(t, 7, 4e-3, 4.0)


def test_siglip_matches_stock_transformers():
    transformers = pytest.importorskip("transformers")
    oc = O.tiny_config()
    p = _fp32(O.init_params(oc, 3))
    cfg = transformers.SiglipVisionConfig(
        hidden_size=oc.vit_width, intermediate_size=oc.vit_mlp_dim, num_hidden_layers=oc.vit_depth,
        num_attention_heads=oc.vit_heads, image_size=oc.image_size, patch_size=oc.vit_patch,
        hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6, attn_implementation="eager")
    cfg.vision_use_head = False
    hf = transformers.SiglipVisionModel(cfg).eval()
    pre = "paligemma_with_expert.paligemma.model.vision_tower."
    sd = {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("head" in m for m in missing), missing
    img = O.synthetic_batch(oc, 2)["images"][0]
    with torch.no_grad():
        ref = hf(pixel_values=img).last_hidden_state
        ref = torch.nn.functional.linear(ref, p["paligemma_with_expert.paligemma.model.multi_modal_projector.linear.weight"],
                                         p["paligemma_with_expert.paligemma.model.multi_modal_projector.linear.bias"])
        got = O.siglip_embed_image(p, oc, img)
    assert H.rel_err(got, ref) < 2e-5


def test_gemma_prefix_matches_stock_transformers():
    transformers = pytest.importorskip("transformers")
    oc = O.tiny_config()
    p = _fp32(O.init_params(oc, 4))
    g = oc.paligemma
    cfg = transformers.GemmaConfig(
        vocab_size=oc.vocab_size, hidden_size=g.width, intermediate_size=g.mlp_dim, num_hidden_layers=g.depth,
        num_attention_heads=g.num_heads, num_key_value_heads=g.num_kv_heads, head_dim=g.head_dim,
        hidden_activation="gelu_pytorch_tanh", hidden_act="gelu_pytorch_tanh", rms_norm_eps=1e-6, rope_theta=10000.0,
        attention_bias=False, attn_implementation="eager")
    hf = transformers.GemmaModel(cfg).eval()
    pre = "paligemma_with_expert.paligemma.model.language_model."
    sd = {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    B, T = 2, 12
    x = torch.randn(B, T, g.width, generator=torch.Generator().manual_seed(0))
    # causal attention with right-padding on the second sequence: expressible both as the reference's
    # (pad_masks, att_masks) pair and as stock HF's 2-D attention_mask
    pad = torch.ones(B, T, dtype=torch.bool)
    pad[1, T - 3:] = False
    pos = torch.cumsum(pad, dim=1) - 1
    mask4d = O.prepare_attention_masks_4d(O.make_att_2d_masks(pad, torch.ones(B, T, dtype=torch.int64)))
    with torch.no_grad():
        got, _ = O.single_stream_forward(p, oc, "prefix", x, mask4d, pos)
        # transformers 5.x applies Gemma's sqrt(hidden) normaliser inside the token embedding only, so inputs_embeds
        # pass through unscaled exactly as in the reference's patched file (modeling_gemma.py:515-516)
        ref = hf(inputs_embeds=x, attention_mask=pad.to(torch.int64), position_ids=pos).last_hidden_state
    assert H.rel_err(got[pad], ref[pad]) < 1e-5


def test_kv_cache_decode_equals_joint_forward():
    """sample_actions' per-step v_t (cache path, gemma_pytorch.py:102-125) == joint forward on the suffix rows."""
    oc = O.tiny_config()
    p = _fp32(O.init_params(oc, 5))
    b = O.synthetic_batch(oc, 2, ragged=True)
    with torch.no_grad():
        v_joint, _ = O.model_v_t(p, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"], b["time"])
        pad, cache = O.prefill(p, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"])
        v_cache = O.denoise_step(p, oc, pad, cache, b["noise"], b["time"])
    assert H.rel_err(v_cache, v_joint) < 1e-4


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_oracle_reproduces_golden(name):
    gold = torch.load(os.path.join(GOLD, f"{name}_b2.pt"))
    oc = O.tiny_config() if name == "tiny" else H.mid_config()
    params = O.init_params(oc, gold["weight_seed"])
    b = O.synthetic_batch(oc, gold["batch"], ragged=True)
    chk = torch.stack([i.double().sum() for i in b["images"]])
    assert torch.equal(chk, gold["image_checksum"]), "synthetic inputs changed"
    assert torch.equal(b["tokens"], gold["tokens"]) and torch.equal(b["noise"], gold["noise"])
    masks = list(gold["img_masks"])
    with torch.no_grad():
        loss = O.forward_loss(params, oc, b["images"], masks, b["tokens"], b["token_mask"], b["actions"], b["noise"],
                              b["time"])
    # bf16 CPU kernels may block differently across hosts: compare at the bf16 noise floor, not bit-exactly
    assert H.rel_err(loss, gold["loss"]) < 3e-3


def test_bf16_noise_floor_of_the_oracle_itself():
    """Two mathematically equivalent evaluations of the oracle (bf16 linears vs the same linears computed in fp32 and
    rounded once) differ by the bf16 noise floor; the GPU parity thresholds in test_engine_gpu.py sit at ~2x this."""
    oc = H.mid_config()
    params = O.init_params(oc, 0)
    b = O.synthetic_batch(oc, 2, ragged=True)
    args = (b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"], b["noise"], b["time"])
    with torch.no_grad():
        t1 = {}
        O.forward_loss(params, oc, *args, t1)
        orig = torch.nn.functional.linear

        def linear_fp32_once(x, w, bias=None):
            if x.dtype == torch.bfloat16:
                y = orig(x.float(), w.float(), None if bias is None else bias.float())
                return y.to(torch.bfloat16)
            return orig(x, w, bias)

        torch.nn.functional.linear = linear_fp32_once
        O.F.linear = linear_fp32_once
        try:
            t2 = {}
            O.forward_loss(params, oc, *args, t2)
        finally:
            torch.nn.functional.linear = orig
            O.F.linear = orig
    floor_v = H.rel_err(t2["v_t"], t1["v_t"])
    floor_s = H.rel_err(t2["suffix_out"], t1["suffix_out"])
    print(f"oracle self-noise: v_t {floor_v:.2e} suffix_out {floor_s:.2e}")
    assert floor_v < 1e-2 and floor_s < 1e-2


def test_advantage_loss_reduces_to_the_plain_loss_and_broadcasts_the_value_term():
    """AdvantageEstimator.forward (pi0_pytorch.py:560-587): with w_value = 0, w_action = 1 the loss is the plain flow
    loss averaged over the action dimension; the value term is one number per sample broadcast over the horizon, with
    the progress target clamped to [-1, 1]."""
    oc = O.tiny_config(value_head=True)
    p = O.init_params(oc, seed=3)
    b = O.synthetic_batch(oc, 2)
    args = (b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"], b["noise"], b["time"])
    with torch.no_grad():
        plain = O.forward_loss(p, oc, *args).mean(dim=-1)
        prog = torch.tensor([0.4, -3.0])
        la = O.advantage_forward_loss(p, oc, *args, prog, loss_action_weight=1.0, loss_value_weight=0.0)
        both = O.advantage_forward_loss(p, oc, *args, prog, loss_action_weight=1.0, loss_value_weight=2.0)
        clamped = O.advantage_forward_loss(p, oc, *args, torch.tensor([0.4, -1.0]), loss_action_weight=1.0,
                                           loss_value_weight=2.0)
    assert la.shape == (2, oc.action_horizon) and torch.allclose(la, plain, rtol=0, atol=0)
    extra = both - la
    assert torch.allclose(extra, extra[:, :1].expand_as(extra)) and bool((extra > 0).all())
    assert torch.equal(both, clamped)  # progress -3 is clamped to -1
