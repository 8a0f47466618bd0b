"""CPU checks of oracle/rtc_oracle.py (the engine side is tests/test_rtc_gpu.py).

PINNED to the reference's own sampler: `Pi0RTC.sample_actions` of models/pi0_rtc.py is executed in place by
tools/reference_rtc_loader.py (jax.numpy mapped onto torch, the flax sub-networks replaced by the PyTorch-path oracle
network) -- live when /root/reference is present, and through the outputs committed in tests/golden/rtc_reference.pt.
Further: known answers of the schedules, properties of the guided sampler, and the vector-Jacobian product against
central finite differences of the PyTorch-path denoise step in float32."""
import math

import pytest
import torch

import os
import sys

import helpers as H  # noqa: F401
from oracle import pi05_oracle as O
from oracle import rtc_oracle as R

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import make_golden_rtc as MGR  # noqa: E402
import reference_rtc_loader as RRL  # noqa: E402

GOLD_RTC = os.path.join(os.path.dirname(__file__), "golden", "rtc_reference.pt")


def _oracle_run(oc, p, b, kw):
    return R.sample_actions_rtc(p, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"], **kw)


def test_oracle_reproduces_the_references_own_sampler_outputs():
    """tests/golden/rtc_reference.pt: what pi0_rtc.py's `Pi0RTC.sample_actions` returned (tools/make_golden_rtc.py).
    Same arithmetic on another CPU may group float32 sums differently, hence 2e-5 instead of equality (the live test
    below asserts equality on the machine that runs both)."""
    gold = torch.load(GOLD_RTC)
    oc, p, b, plain, prev = MGR.setup("float32")
    assert H.rel_err(plain, gold["plain"]) < 2e-5
    worst = 0.0
    for name, kw in MGR.cases(prev).items():
        got = _oracle_run(oc, p, b, kw)
        e = H.rel_err(got, gold[name])
        worst = max(worst, e)
        assert e < 2e-5, (name, e)
        # the JAX-side suffix embedding (float32 sincos) instead of the PyTorch path's (float64): same answer to 1e-5
        assert H.rel_err(got, gold[name + "/jax_suffix_embedding"]) < 1e-5, name
        guided = name not in ("rtc_disabled", "no_previous_chunk")
        assert (H.rel_err(got, plain) > 0.05) == guided, name  # the guidance really moves the chunk (7 .. 24 %)
    print(f"rtc oracle vs the reference's sampler outputs: worst rel {worst:.1e}")


@pytest.mark.skipif(not RRL.available(), reason="needs /root/reference (build container)")
def test_oracle_equals_the_references_sampler_run_live():
    oc, p, b, plain, prev = MGR.setup("float32")
    for name, kw in MGR.cases(prev).items():
        ref = RRL.sample_actions(p, oc, b, b["noise"], oracle_suffix_embedding=True, **kw)
        got = _oracle_run(oc, p, b, kw)
        assert torch.equal(got, ref), (name, float((got - ref).abs().max()))  # same operations in the same order
    # a 2-D previous chunk (no batch axis) at batch 1 (pi0_rtc.py:310-311), and the bfloat16 dtype map
    oc, p16, b, plain, prev = MGR.setup("bfloat16")
    b1 = {k: ([t[:1] for t in v] if isinstance(v, list) else v[:1]) for k, v in b.items()}
    kw = dict(prev_action_chunk=prev[0, :, :14], inference_delay=2, execute_horizon=6)
    ref = RRL.sample_actions(p16, oc, b1, b1["noise"], oracle_suffix_embedding=True, **kw)
    got = _oracle_run(oc, p16, b1, kw)
    assert torch.equal(got, ref)


def test_prefix_weights_equal_the_references_function():
    """pi0_rtc.py:47-61 executed in place (fixture; live too when the checkout is present) against the oracle's and the
    PRODUCT'S host-side schedule (kai0_b200.pi0_pytorch.PI0Pytorch.rtc_prefix_weights)."""
    from kai0_b200.pi0_pytorch import PI0Pytorch

    gold = torch.load(GOLD_RTC)["prefix_weights"]
    assert len(gold) == 20
    for key, want in gold.items():
        sched, a, e, t = key.split("/")
        a, e, t = int(a), int(e), int(t)
        assert torch.allclose(R.get_prefix_weights(a, e, t, sched), want, rtol=0, atol=1e-7), key
        assert torch.allclose(PI0Pytorch.rtc_prefix_weights(a, e, t, sched), want, rtol=0, atol=1e-7), key
        if RRL.available():
            live = RRL.load().get_prefix_weights(a, e, t, sched).as_subclass(torch.Tensor)
            assert torch.equal(live, want), key


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
