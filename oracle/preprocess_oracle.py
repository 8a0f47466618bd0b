"""CPU restatement of the reference's observation preprocessing for the PyTorch path — TEST INFRASTRUCTURE ONLY
(imported by tests/, tools/make_golden_preprocess.py; never by the product).

Follows src/openpi/models_pytorch/preprocessing_pytorch.py:20-173 (`preprocess_observation_pytorch`, "P:" below) and
src/openpi/shared/image_tools.py:55-126 (`resize_with_pad_torch`, "T:" below), op for op in plain torch on the CPU.
The reference draws its augmentation parameters from the global torch RNG while it goes; here the draw is a separate
function (`draw_params`, same calls in the same order) so tests can inject the numbers into both sides.

Pinned: tests/golden/preprocess_*.pt hold outputs of the reference's own function (imported from /root/reference with
jax stubbed out, tools/make_golden_preprocess.py) and tests/test_preprocess_cpu.py checks this file against them.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# columns of one parameter row (one row per image key)
START_H, START_W, ANGLE, BRIGHT, CONTRAST, SATUR = range(6)
NPARAM = 6


def is_wrist(key: str) -> bool:
    return "wrist" in key  # P:59


def draw_params(keys, height: int, width: int, device="cpu") -> torch.Tensor:
    """The random numbers of P:60-133 for `keys`, drawn with the same torch calls in the same order (one draw per
    BATCH, not per sample).  Returns fp32 [len(keys), 6]; unused entries (geometry of wrist cameras) are 0."""
    out = torch.zeros(len(keys), NPARAM, dtype=torch.float32)
    crop_h, crop_w = int(height * 0.95), int(width * 0.95)
    max_h, max_w = height - crop_h, width - crop_w
    for i, key in enumerate(keys):
        if not is_wrist(key):
            if max_h > 0 and max_w > 0:
                out[i, START_H] = float(torch.randint(0, max_h + 1, (1,), device=device))  # P:71
                out[i, START_W] = float(torch.randint(0, max_w + 1, (1,), device=device))  # P:72
            out[i, ANGLE] = float(torch.rand(1, device=device) * 10 - 5)  # P:85
        out[i, BRIGHT] = float(0.7 + torch.rand(1, device=device) * 0.6)  # P:124
        out[i, CONTRAST] = float(0.6 + torch.rand(1, device=device) * 0.8)  # P:129
        out[i, SATUR] = float(0.5 + torch.rand(1, device=device) * 1.0)  # P:137
    return out


def resize_with_pad(images_nhwc: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """T:55-126 for fp32 channels-last input: bilinear resize keeping the aspect ratio, clamp to [-1, 1], pad with -1."""
    x = images_nhwc.permute(0, 3, 1, 2)
    cur_h, cur_w = x.shape[2], x.shape[3]
    ratio = max(cur_w / width, cur_h / height)  # T:88
    rh, rw = int(cur_h / ratio), int(cur_w / ratio)  # T:89-90
    x = F.interpolate(x, size=(rh, rw), mode="bilinear", align_corners=False)  # T:93-95
    x = x.clamp(-1.0, 1.0)  # T:100-101
    ph0, rem_h = divmod(height - rh, 2)  # T:106-109
    pw0, rem_w = divmod(width - rw, 2)
    x = F.pad(x, (pw0, pw0 + rem_w, ph0, ph0 + rem_h), mode="constant", value=-1.0)  # T:112-118
    return x.permute(0, 2, 3, 1)


def _crop_resize(img01: torch.Tensor, start_h: int, start_w: int) -> torch.Tensor:
    """P:62-82 on [B,H,W,C] in [0,1]: 95 % crop at (start_h, start_w), bilinear resize back to HxW."""
    h, w = img01.shape[1:3]
    ch, cw = int(h * 0.95), int(w * 0.95)
    if h - ch > 0 and w - cw > 0:  # P:70
        img01 = img01[:, start_h : start_h + ch, start_w : start_w + cw, :]
    return F.interpolate(img01.permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)


def _rotate(img01: torch.Tensor, angle_deg: float) -> torch.Tensor:
    """P:84-120: rotate by `angle_deg` with grid_sample (bilinear, zero padding) when |angle| > 0.1."""
    angle = torch.tensor([angle_deg], dtype=torch.float32)
    if not bool(torch.abs(angle) > 0.1):  # P:86
        return img01
    h, w = img01.shape[1:3]
    rad = angle * torch.pi / 180.0
    cos_a, sin_a = torch.cos(rad), torch.sin(rad)
    gx = torch.linspace(-1, 1, w)
    gy = torch.linspace(-1, 1, h)
    gy, gx = torch.meshgrid(gy, gx, indexing="ij")
    gx = gx.unsqueeze(0).expand(img01.shape[0], -1, -1)
    gy = gy.unsqueeze(0).expand(img01.shape[0], -1, -1)
    grid = torch.stack([gx * cos_a - gy * sin_a, gx * sin_a + gy * cos_a], dim=-1)  # P:106-110
    out = F.grid_sample(img01.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return out.permute(0, 2, 3, 1)


def _colour(img01: torch.Tensor, bright: float, contrast: float, satur: float) -> torch.Tensor:
    """P:122-143: brightness, contrast about the per-sample mean, saturation about the per-pixel grey, clamp."""
    b = torch.tensor([bright], dtype=torch.float32)
    c = torch.tensor([contrast], dtype=torch.float32)
    s = torch.tensor([satur], dtype=torch.float32)
    x = img01 * b
    mean = x.mean(dim=[1, 2, 3], keepdim=True)
    x = (x - mean) * c + mean
    gray = x.mean(dim=-1, keepdim=True)
    x = gray + (x - gray) * s
    return torch.clamp(x, 0, 1)


def preprocess_images(images: dict, keys, *, train: bool, params: torch.Tensor | None, resolution=(224, 224)) -> dict:
    """P:35-148 for the image dict: returns {key: image} in the layout each input came in (NCHW stays NCHW)."""
    out = {}
    for i, key in enumerate(keys):
        img = images[key]
        channels_first = img.shape[1] == 3  # P:42
        if channels_first:
            img = img.permute(0, 2, 3, 1)
        if tuple(img.shape[1:3]) != tuple(resolution):  # P:48-50
            img = resize_with_pad(img, *resolution)
        if train:
            p = params[i]
            img = img / 2.0 + 0.5  # P:54
            if not is_wrist(key):
                img = _crop_resize(img, int(p[START_H]), int(p[START_W]))
                img = _rotate(img, float(p[ANGLE]))
            img = _colour(img, float(p[BRIGHT]), float(p[CONTRAST]), float(p[SATUR]))
            img = img * 2.0 - 1.0  # P:146
        if channels_first:
            img = img.permute(0, 3, 1, 2)
        out[key] = img
    return out


def rotation_is_identity(angle_deg: float) -> bool:
    return not (math.fabs(angle_deg) > 0.1)
