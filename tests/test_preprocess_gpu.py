"""GPU parity of the device preprocessing (pi05_preprocess_image through the
reference-facing class) against (1) golden outputs of the reference's own preprocess_observation_pytorch and
(2) the CPU oracle at the benchmark's 224x224 resolution.

Tolerance: fp32 images in [-1, 1]; the kernel evaluates the same bilinear / grid_sample / colour formulas with a
different fused-multiply-add contraction and CUDA sinf/cosf, so values agree to ~1e-5 absolute (asserted: 1e-4);
the eval path without resize is a pure copy and must be bit-exact."""
import os

import pytest
import torch

import helpers as H
from oracle import pi05_oracle as O
from oracle import preprocess_oracle as PO
from test_preprocess_cpu import GOLD, KEYS, golden_cases, make_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-4


class _Obs:
    def __init__(self, images, B, device="cuda"):
        self.images = {k: v.to(device) for k, v in images.items()}
        self.image_masks = {}
        self.state = torch.zeros(B, 32, device=device)
        self.tokenized_prompt = torch.zeros(B, 4, dtype=torch.int64, device=device)
        self.tokenized_prompt_mask = torch.ones(B, 4, dtype=torch.bool, device=device)


def _model(image_size):
    oc = O.tiny_config(image_size=image_size) if image_size != 224 else None
    if oc is None:
        oc = O.tiny_config(image_size=224)
    model, _ = H.build_pair(oc, seed=0)
    model.augment = True
    return model


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[11:-3])
def test_device_preprocessing_matches_reference_goldens(path):
    g = torch.load(path)
    res = tuple(g["resolution"])
    model = _model(res[0])
    images = make_inputs(g["seed"], g["batch"], g["shapes"], g["layout"])
    model._augment_params_override = g["params"]
    out, masks, *_ = model._preprocess_observation(_Obs(images, g["batch"]), train=g["train"])
    torch.cuda.synchronize()
    assert out.shape == (3, g["batch"], 3, res[0], res[1]) and len(masks) == 3 and all(bool(m.all()) for m in masks)
    for i, k in enumerate(KEYS):
        ref = g["outputs"][k]
        if g["layout"] == "nhwc":
            ref = ref.permute(0, 3, 1, 2)  # the engine always consumes NCHW
        assert H.max_err(out[i], ref) < TOL, (k, H.max_err(out[i], ref))


def test_full_resolution_augmentation_and_resize_match_oracle():
    """224x224 target (BASELINE configs): one camera needs resize-with-pad from 480x640, all are augmented; also the
    rotation-skipped branch (|angle| <= 0.1) and a crop at the far corner."""
    model = _model(224)
    B = 2
    shapes = [(480, 640), (224, 224), (224, 224)]
    images = make_inputs(21, B, shapes, "nchw")
    params = torch.tensor([[12.0, 12.0, 0.05, 0.9, 1.3, 0.6],
                           [0.0, 0.0, 0.0, 1.25, 0.7, 1.4],
                           [0.0, 0.0, 0.0, 0.75, 1.0, 1.0]])
    for angle in (0.05, -4.5):
        params[0, PO.ANGLE] = angle
        model._augment_params_override = params
        out, *_ = model._preprocess_observation(_Obs(images, B), train=True)
        ref = PO.preprocess_images(images, KEYS, train=True, params=params, resolution=(224, 224))
        for i, k in enumerate(KEYS):
            assert H.max_err(out[i], ref[k]) < TOL, (angle, k)
    # eval path, no resize: bit-exact copy into the stacked layout; NHWC input gives the same result as NCHW
    same = make_inputs(22, B, [(224, 224)] * 3, "nchw")
    out, *_ = model._preprocess_observation(_Obs(same, B), train=False)
    for i, k in enumerate(KEYS):
        assert torch.equal(out[i].cpu(), same[k])
    nhwc = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in same.items()}
    out2, *_ = model._preprocess_observation(_Obs(nhwc, B), train=False)
    assert torch.equal(out, out2)


def test_draws_follow_the_reference_call_order_on_the_device():
    """Same seed -> same parameters as the reference's sequence of torch.randint / torch.rand calls on that device."""
    model = _model(56)
    torch.manual_seed(77)
    mine = model._draw_augment_params(KEYS, 56, torch.device("cuda"))
    torch.manual_seed(77)
    ref = PO.draw_params(KEYS, 56, 56, device="cuda")
    assert torch.equal(mine.cpu(), ref)


def test_training_forward_with_augmentation_matches_oracle_on_augmented_images():
    """End to end: loss of the engine with augmentation on == oracle loss on the oracle-augmented images."""
    oc = O.tiny_config()
    model, params = H.build_pair(oc, seed=4)
    model.augment = True
    batch = O.synthetic_batch(oc, 2)
    torch.manual_seed(3)
    p = PO.draw_params(KEYS, oc.image_size, oc.image_size)
    model._augment_params_override = p
    imgs = PO.preprocess_images(dict(zip(KEYS, batch["images"])), KEYS, train=True, params=p,
                                resolution=(oc.image_size, oc.image_size))
    ref = O.forward_loss(params, oc, [imgs[k] for k in KEYS], batch["img_masks"], batch["tokens"], batch["token_mask"],
                         batch["actions"], batch["noise"], batch["time"])
    model.eval()
    with torch.no_grad():
        loss = model(H.Obs(batch, "cuda"), batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda())
    assert H.rel_err(loss, ref) < 4e-3
    base = O.forward_loss(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                          batch["actions"], batch["noise"], batch["time"])
    assert H.rel_err(base, ref) > 1e-3  # the augmentation does change the result


def _unfold_rows(rows, n_keys, B, S, patch, kp):
    """bf16 patch rows [n_keys*B*T, 3*kp] -> the three blocks as fp32 images [n_keys, B, 3, S, S] (inverse of the im2col
    layout of modeling_siglip.py:220-226: column = c * p * p + (y % p) * p + (x % p))."""
    Pn = S // patch
    k = 3 * patch * patch
    r = rows.float().view(n_keys, B, Pn, Pn, 3, kp)
    out = []
    for blk in range(3):
        x = r[..., blk, :k].reshape(n_keys, B, Pn, Pn, 3, patch, patch)  # [n, b, py, px, c, iy, ix]
        out.append(x.permute(0, 1, 4, 2, 5, 3, 6).reshape(n_keys, B, 3, S, S))
    pad = r[..., :, k:]
    return out, pad


@pytest.mark.parametrize("train", [False, True])
def test_patch_row_output_is_the_split_im2col_of_the_fp32_path(train):
    """Row f2: pi05_preprocess_patches writes the SAME pixels pi05_preprocess_image writes, straight into the patch-GEMM
    operand: block 0 = block 2 = bf16(v) bit-exactly, block 0 + block 1 = v to 2^-16, padding columns zero; and uint8
    input (from_dict folded into the kernel) gives bit-identical rows to from_dict's fp32 output."""
    model = _model(224)
    B, S, patch = 2, 224, 14
    shapes = [(480, 640), (224, 224), (224, 224)]
    g = torch.Generator().manual_seed(31)
    u8 = {k: torch.randint(0, 256, (B, *shp, 3), generator=g, dtype=torch.uint8) for k, shp in zip(KEYS, shapes)}
    f32 = {k: v.to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0 for k, v in u8.items()}  # Observation.from_dict
    params = torch.tensor([[7.0, 3.0, -4.5, 0.9, 1.3, 0.6], [0.0, 0.0, 0.0, 1.25, 0.7, 1.4], [0.0, 0.0, 0.0, 0.75, 1.0, 1.0]])
    model._augment_params_override = params
    img, *_ = model._preprocess_observation(_Obs(f32, B), train=train)
    rows_f, *_ = model._preprocess_observation(_Obs(f32, B), train=train, rows=True)
    rows_u, *_ = model._preprocess_observation(_Obs(u8, B), train=train, rows=True)
    torch.cuda.synchronize()
    from kai0_b200 import _lib

    kp = int(_lib.lib().pi05_patch_row_kp(patch))
    assert rows_f.rows.shape == (3 * B * (S // patch) ** 2, 3 * kp) and len(rows_f) == 3
    assert torch.equal(rows_f.rows, rows_u.rows)
    (hi, lo, hi2), pad = _unfold_rows(rows_f.rows, 3, B, S, patch, kp)
    assert torch.equal(hi, img.to(torch.bfloat16).float())
    assert torch.equal(hi2, hi)
    assert float((hi + lo - img).abs().max()) <= 2.0 ** -16
    assert float(pad.abs().max()) == 0.0


def test_patch_embedding_on_split_rows_matches_the_fp32_convolution():
    """vit_embed through the tcgen05 GEMM on [hi | lo | hi] x [Whi | Whi | Wlo] against (a) the engine's own fp32 im2col
    convolution and (b) the oracle's fp32 Conv2d: the products are exact to 2^-16, so after the rounding to bf16 the two
    differ only where a value sits on a rounding boundary."""
    oc = H.mid_config()
    model, params = H.build_pair(oc, seed=8)
    batch = O.synthetic_batch(oc, 2, seed=9)
    obs = H.Obs(batch, "cuda")
    model.set_taps(True)
    model.eval()
    taps = {}
    with torch.no_grad():
        O.forward_loss(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                       batch["actions"], batch["noise"], batch["time"], taps)
    ref = torch.stack([taps[f"img{n}_vit_embed"] for n in range(oc.num_images)]).float()
    got = {}
    for mode in (True, False):
        model.use_patch_rows = mode
        with torch.no_grad():
            model(obs, batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda())
        got[mode] = model.get_tap("vit_embed").float().cpu().view(ref.shape)
    e_rows, e_conv, e_between = H.rel_err(got[True], ref), H.rel_err(got[False], ref), H.rel_err(got[True], got[False])
    print(f"\n[patch] vit_embed vs oracle: split-row GEMM {e_rows:.3e}, fp32 im2col {e_conv:.3e}; between the two {e_between:.3e}")
    assert e_rows < 5e-4 and e_conv < 5e-4 and e_between < 5e-4


def test_patch_row_path_gradients_match_the_fp32_path():
    oc = H.mid_config()
    batch = O.synthetic_batch(oc, 2, seed=10)
    grads = {}
    for mode in (True, False):
        model, _ = H.build_pair(oc, seed=8)
        model.use_patch_rows = mode
        model.train()
        model(H.Obs(batch, "cuda"), batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda()).mean().backward()
        torch.cuda.synchronize()
        emb = model.paligemma_with_expert.paligemma.model.vision_tower.vision_model.embeddings
        grads[mode] = [emb.patch_embedding.weight.grad.clone(), emb.patch_embedding.bias.grad.clone(),
                       emb.position_embedding.weight.grad.clone()]
    for a, b, name in zip(grads[True], grads[False], ("weight", "bias", "position")):
        assert H.rel_err(a, b) < 2e-2, (name, H.rel_err(a, b))
