"""CPU checks of oracle/rtc_oracle.py (prepared; the engine does not implement RTC yet).
PARITY UNPINNED: the reference's RTC exists only in JAX (models/pi0_rtc.py) and cannot run here — these tests pin the
restatement to known answers of the schedules and to properties of the guided sampler."""
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
