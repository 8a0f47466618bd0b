"""CPU restatement of the reference's real-time-chunking (RTC) guided decoding — TEST INFRASTRUCTURE ONLY (SURVEY.md §8
row f4; imported by tests/ only, never by the product).

The reference implements RTC only in its JAX model (src/openpi/models/pi0_rtc.py, "R:" below): `get_prefix_weights`
R:47-61 and the guided Euler loop of `Pi0RTC.sample_actions` R:234-360.  This file restates that algorithm in plain torch
on top of the PyTorch-path network of oracle/pi05_oracle.py (prefill + KV-cache denoise step), which is what row f4
asks for ("RTC guided decoding for the PyTorch path").  The vector-Jacobian product of R:331 (`jax.vjp(denoiser, x)`)
is torch.autograd through the oracle's denoise step.

PARITY PINNED to the reference's own sampler code: jax is not installed here, but tools/reference_rtc_loader.py executes
pi0_rtc.py UNMODIFIED with `jax.numpy` mapped onto torch and the four flax sub-networks replaced by the PyTorch-path
oracle network (itself pinned to the reference's PyTorch model), so `Pi0RTC.embed_prefix`, `make_attn_mask`, the KV-cache
protocol, the scan and the whole guidance computation run as the reference wrote them.  On twelve seeded settings
(schedules, delays, delay masking, 7 / 14 / 36-dim and NaN-carrying previous chunks, clipped arguments, 5 steps, RTC off)
this file reproduces that run BIT FOR BIT in float32 (`tests/test_rtc_oracle_cpu.py`; outputs committed by
tools/make_golden_rtc.py as tests/golden/rtc_reference.pt).  With the JAX-side `embed_suffix` left in place (its sincos
embedding is float32, the PyTorch path's float64) the two differ by <= 7.4e-6 on O(5) values.  What stays unpinned is
JAX's own network arithmetic (XLA dots), which the PyTorch path never matched bit for bit either.
"""
from __future__ import annotations

import math

import torch

from oracle import pi05_oracle as O


def get_prefix_weights(start: int, end: int, total: int, schedule: str) -> torch.Tensor:
    """R:47-61: per-timestep guidance weights over an action chunk of `total` steps (1 up to `start`, decaying to 0 at
    `end`, 0 afterwards)."""
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


def guidance_weight(time: float, max_guidance_weight: float) -> float:
    """R:341-347: tau = 1 - time (clipped to [1e-3, 1]); min(c * inv_r2, max) with c = (1 - tau) / tau."""
    tau = torch.clamp(torch.tensor(1.0 - time, dtype=torch.float32), 1e-3, 1.0)
    sq = (1 - tau) ** 2
    inv_r2 = (sq + tau**2) / sq  # inf at tau = 1 (time = 0), as in the reference: 0 * inf = nan is zeroed by nan_to_num later
    c = torch.nan_to_num((1 - tau) / tau, posinf=max_guidance_weight)
    gw = c * inv_r2
    return float(torch.where(torch.isnan(gw), gw, torch.minimum(gw, torch.tensor(max_guidance_weight))))


def sample_actions_rtc(p, cfg, images, img_masks, lang_tokens, lang_masks, noise, num_steps: int = 10, *,
                       prev_action_chunk: torch.Tensor | None = None, inference_delay: int | None = None,
                       execute_horizon: int | None = None, mask_prefix_delay: bool = False,
                       prefix_attention_schedule: str = "exp", max_guidance_weight: float = 0.5,
                       enable_rtc: bool = True) -> torch.Tensor:
    """R:234-360 on the PyTorch network.  Time runs 1 -> 0 in `num_steps` Euler steps as an fp32 running sum
    (pi0_pytorch.py:401-418); every step with a previous chunk adds the RTC correction to the velocity."""
    bsize = noise.shape[0]
    H, A = cfg.action_horizon, cfg.action_dim
    with torch.no_grad():
        prefix_pad, cache = O.prefill(p, cfg, images, img_masks, lang_tokens, lang_masks)
    dt = torch.tensor(-1.0 / num_steps, dtype=torch.float32)
    use_rtc = enable_rtc and prev_action_chunk is not None
    if use_rtc:
        exec_h = int(min(max(execute_horizon if execute_horizon is not None else H, 1), H))  # R:305-306
        d = int(min(max(0 if inference_delay is None else inference_delay, 0), H))  # R:307-308
        prev = prev_action_chunk.to(torch.float32)
        if prev.dim() == 2:
            prev = prev[None]
        exec_h = min(exec_h, prev.shape[1])  # R:313
        provided_before_pad = prev.shape[-1]
        prev = torch.nan_to_num(prev, nan=0.0, posinf=0.0, neginf=0.0)  # R:317
        if prev.shape[-1] > A:  # R:319-324
            prev = prev[..., :A]
        elif prev.shape[-1] < A:
            prev = torch.cat([prev, torch.zeros(*prev.shape[:-1], A - prev.shape[-1])], dim=-1)
        provided = min(14, provided_before_pad, A)  # R:326
        dim_mask = (torch.arange(A) < provided).to(torch.float32)[None, None, :]
        weights = get_prefix_weights(d, exec_h, H, prefix_attention_schedule)  # R:337
    x_t = noise.to(torch.float32)
    time = torch.tensor(1.0, dtype=torch.float32)
    for _ in range(num_steps):
        t_b = time.expand(bsize)
        if not use_rtc:
            with torch.no_grad():
                v_t = O.denoise_step(p, cfg, prefix_pad, cache, x_t, t_b)
            v_t = torch.nan_to_num(v_t, nan=0.0, posinf=0.0, neginf=0.0)  # R:295
        else:
            x_in = x_t
            if mask_prefix_delay and provided > 0:  # R:328-333
                mask_time = (torch.arange(H) < d)[None, :, None]
                x_in = x_t.clone()
                x_in[..., :provided] = torch.where(mask_time, prev[..., :provided], x_t[..., :provided])
            x_local = x_in.detach().requires_grad_(True)
            v_local = O.denoise_step(p, cfg, prefix_pad, cache, x_local, t_b)
            x_1 = x_local - time * v_local  # R:336-338: the action endpoint of this trajectory
            error = ((prev - x_1.detach()) * weights[None, :, None] * dim_mask)  # R:338
            (corr,) = torch.autograd.grad(x_1, x_local, grad_outputs=error)  # R:331,339: J^T error
            gw = guidance_weight(float(time), max_guidance_weight)
            v_t = torch.nan_to_num(v_local.detach() - gw * corr, nan=0.0, posinf=0.0, neginf=0.0)  # R:348-349
        x_t = x_t + dt * v_t
        time = time + dt
    return torch.nan_to_num(x_t, nan=0.0, posinf=0.0, neginf=0.0)  # R:359
