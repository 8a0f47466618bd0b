"""Host-side input container mirroring `openpi.models.model.Observation` (src/openpi/models/model.py:84-164) for
torch tensors, so callers that cannot import the reference's jax-typed dataclass can still build the same object.
Any object with these attributes is accepted by `PI0Pytorch` (the reference's own `Observation` included)."""
from __future__ import annotations

import dataclasses

import torch


@dataclasses.dataclass
class Observation:
    images: dict
    image_masks: dict
    state: torch.Tensor
    tokenized_prompt: torch.Tensor | None = None
    tokenized_prompt_mask: torch.Tensor | None = None
    token_ar_mask: torch.Tensor | None = None
    token_loss_mask: torch.Tensor | None = None
    # AdvantageEstimator fields (model.py:109-119)
    episode_index: torch.Tensor | None = None
    frame_index: torch.Tensor | None = None
    progress: torch.Tensor | None = None
    episode_length: torch.Tensor | None = None
    image_original: dict | None = None

    @classmethod
    def from_dict(cls, data: dict, *, keep_uint8: bool = False) -> "Observation":
        """model.py:122-157: uint8 [B,H,W,3] images become fp32 [B,3,H,W] in [-1,1]; other fields pass through.
        keep_uint8=True (opt-in, not in the reference): leave uint8 images as they are -- the B200 engine takes them
        directly and applies the identical `x / 255 * 2 - 1` inside its preprocessing kernel (pi05_preprocess_patches),
        which removes three element-wise kernels and a 4x larger fp32 copy of every image from the step."""
        if ("tokenized_prompt" in data) != ("tokenized_prompt_mask" in data):
            raise ValueError("tokenized_prompt and tokenized_prompt_mask must be provided together.")
        images = {}
        for key, img in data["image"].items():
            if img.dtype == torch.uint8 and not keep_uint8:
                img = img.to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0
            images[key] = img
        return cls(
            images=images,
            image_masks=data["image_mask"],
            state=data["state"],
            tokenized_prompt=data.get("tokenized_prompt"),
            tokenized_prompt_mask=data.get("tokenized_prompt_mask"),
            token_ar_mask=data.get("token_ar_mask"),
            token_loss_mask=data.get("token_loss_mask"),
            frame_index=data.get("frame_index"),
            episode_length=data.get("episode_length"),
            progress=data.get("progress"),
            image_original=data.get("image_original"),
            episode_index=data.get("episode_index"),
        )
