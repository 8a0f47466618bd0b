"""CPU checks of oracle/rtc_oracle.py (the engine side is tests/test_rtc_gpu.py).
PARITY UNPINNED AGAINST JAX: the reference's RTC exists only in its JAX model (models/pi0_rtc.py) and cannot run here —
these tests pin the restatement to known answers of the schedules, to properties of the guided sampler, and pin the
vector-Jacobian product it relies on with central finite differences of the PyTorch-path denoise step in float32."""
import math

import pytest
import torch

import helpers as H  # noqa: F401
from oracle import pi05_oracle as O
from oracle import rtc_oracle as R


def test_prefix_weight_schedules_known_answers():
    # start 2, end 6, total 8 (pi0_rtc.py:47-61)
    assert torch.equal(R.get_prefix_weights(2, 6, 8, "ones"), torch.tensor([1., 1, 1, 1, 1, 1, 0, 0]))
    assert torch.equal(R.get_prefix_weights(2, 6, 8, "zeros"), torch.tensor([1., 1, 0, 0, 0, 0, 0, 0]))
    lin = R.get_prefix_weights(2, 6, 8, "linear")
    assert torch.allclose(lin, torch.tensor([1., 1, 0.8, 0.6, 0.4, 0.2, 0, 0]))
    exp = R.get_prefix_weights(2, 6, 8, "exp")
    want = lin * torch.expm1(lin) / (math.e - 1)
    assert torch.allclose(exp, want) and float(exp[0]) == pytest.approx(1.0) and float(exp[6]) == 0.0
    # start beyond end is clamped to end
    assert torch.equal(R.get_prefix_weights(9, 3, 5, "linear"), torch.tensor([1., 1, 1, 0, 0]))
    with pytest.raises(ValueError):
        R.get_prefix_weights(1, 2, 3, "cosine")


def test_guidance_weight_formula():
    # time = 1 -> tau clipped to 1e-3: c * inv_r2 is huge -> capped; time = 0.5 -> c = 1, inv_r2 = 2 -> capped at 0.5
    assert R.guidance_weight(1.0, 0.5) == pytest.approx(0.5)
    assert R.guidance_weight(0.5, 5.0) == pytest.approx(2.0)
    assert R.guidance_weight(0.1, 50.0) == pytest.approx((0.1 / 0.9) * ((0.01 + 0.81) / 0.01), rel=1e-4)
    assert R.guidance_weight(0.1, 5.0) == pytest.approx(5.0)
    assert math.isnan(R.guidance_weight(0.0, 0.5))  # 0 * inf, zeroed by nan_to_num in the sampler as in the reference


def test_without_a_previous_chunk_rtc_is_the_plain_sampler():
    oc = O.tiny_config()
    p = O.init_params(oc, seed=2)
    b = O.synthetic_batch(oc, 2)
    args = (b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"])
    plain = O.sample_actions(p, oc, *args)
    assert torch.equal(R.sample_actions_rtc(p, oc, *args, prev_action_chunk=None), plain)
    assert torch.equal(R.sample_actions_rtc(p, oc, *args, prev_action_chunk=plain, enable_rtc=False), plain)


def test_guidance_pulls_the_executed_prefix_towards_the_previous_chunk():
    oc = O.tiny_config()
    p = O.init_params(oc, seed=2)
    b = O.synthetic_batch(oc, 2)
    args = (b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"])
    plain = O.sample_actions(p, oc, *args)
    target = plain + 0.5 * torch.randn(plain.shape, generator=torch.Generator().manual_seed(0))
    target14 = target[..., :14]  # the client provides 14 real dims (R:326)
    guided = R.sample_actions_rtc(p, oc, *args, prev_action_chunk=target14, inference_delay=2, execute_horizon=6,
                                  prefix_attention_schedule="linear", max_guidance_weight=0.5)
    assert guided.shape == plain.shape and bool(torch.isfinite(guided).all())
    w = R.get_prefix_weights(2, 6, oc.action_horizon, "linear")[None, :, None]
    err_plain = ((plain[..., :14] - target14) * w).norm()
    err_guided = ((guided[..., :14] - target14) * w).norm()
    assert float(err_guided) < float(err_plain)
    # dims the client did not provide are never steered directly; steps past the execute horizon get zero weight
    masked = R.sample_actions_rtc(p, oc, *args, prev_action_chunk=target14, inference_delay=2, execute_horizon=6,
                                  mask_prefix_delay=True)
    assert bool(torch.isfinite(masked).all())


def test_vjp_of_the_denoiser_matches_central_finite_differences():
    """pi0_rtc.py:331: `jax.vjp(denoiser, x)` with denoiser(x) = x - t * v(x).  For float32 weights, <J^T e, d> must equal
    <e, (f(x + h d) - f(x - h d)) / 2h> for random directions d (central differences, h = 1e-2 on O(1) inputs)."""
    oc = O.tiny_config()
    p = {k: v.to(torch.float32) for k, v in O.init_params(oc, seed=5).items()}
    b = O.synthetic_batch(oc, 2)
    with torch.no_grad():
        prefix_pad, cache = O.prefill(p, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, oc.action_horizon, oc.action_dim, generator=g)
    e = torch.randn(x.shape, generator=g)
    for t in (1.0, 0.6, 0.1):
        tb = torch.full((2,), t)

        def f(z):
            return z - t * O.denoise_step(p, oc, prefix_pad, cache, z, tb)

        xl = x.clone().requires_grad_(True)
        (vjp,) = torch.autograd.grad(f(xl), xl, grad_outputs=e)
        for _ in range(3):
            d = torch.randn(x.shape, generator=g)
            h = 1e-2
            with torch.no_grad():
                fd = (f(x + h * d) - f(x - h * d)) / (2 * h)
            lhs, rhs = float((vjp * d).sum()), float((e * fd).sum())
            assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), abs(rhs), 1.0), (t, lhs, rhs)
