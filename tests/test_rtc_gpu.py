"""Real-time-chunking guided decoding on the engine (pi05_denoise_rtc behind
`PI0Pytorch.sample_actions(..., prev_action_chunk=...)`) against oracle/rtc_oracle.py, which restates
src/openpi/models/pi0_rtc.py:234-360 on the PyTorch-path network and is pinned bit for bit to that file's own sampler executed
in place (tools/reference_rtc_loader.py; tests/test_rtc_oracle_cpu.py, which also pins the VJP by finite differences).

Tolerances: the velocity of a step is the ordinary decode step (<= 1e-3 on the action chunk, as tests/test_engine_gpu.py);
the vector-Jacobian product is a bf16 backward through 2 x depth layers (1-3 % per tensor, as the training gradients); it
enters the update scaled by guidance * dt <= 0.05 per step, so the guided action chunk is held to 3e-3."""
import pytest
import torch

import helpers as H
from oracle import pi05_oracle as O
from oracle import rtc_oracle as R

pytestmark = pytest.mark.gpu


def _setup(name):
    oc = O.tiny_config() if name == "tiny" else H.mid_config()
    model, params = H.build_pair(oc, seed=6)
    model.eval()
    b = O.synthetic_batch(oc, 2, seed=3, ragged=True)
    b["img_masks"][1][0] = False
    args = (b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"])
    return oc, model, params, b, args


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_rtc_guided_sampler_matches_oracle(name):
    oc, model, params, b, args = _setup(name)
    obs = H.Obs(b, "cuda")
    noise = b["noise"].cuda()
    plain_ref = O.sample_actions(params, oc, *args)
    g = torch.Generator().manual_seed(0)
    target = (plain_ref + 0.5 * torch.randn(plain_ref.shape, generator=g))[..., :14]  # the client provides 14 real dims
    cases = [dict(inference_delay=2, execute_horizon=6, prefix_attention_schedule="linear", max_guidance_weight=0.5),
             dict(inference_delay=3, execute_horizon=oc.action_horizon, prefix_attention_schedule="exp",
                  max_guidance_weight=5.0),
             dict(inference_delay=2, execute_horizon=7, mask_prefix_delay=True, prefix_attention_schedule="ones"),
             dict(inference_delay=0, execute_horizon=4, prefix_attention_schedule="zeros")]
    for kw in cases:
        ref = R.sample_actions_rtc(params, oc, *args, prev_action_chunk=target, **kw)
        got = model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=target.cuda(), **kw)
        err = H.rel_err(got, ref)
        moved = H.rel_err(ref, plain_ref)
        print(f"\\n[rtc {name}] {kw}: engine vs oracle {err:.3e}; guidance moved the chunk by {moved:.3e}")
        assert got.shape == ref.shape and bool(torch.isfinite(got).all())
        assert err < 3e-3, (kw, err)
    # step-0 internals: velocity and the pulled-back error (dv/dx)^T err
    model.set_taps(True)
    kw = cases[0]
    model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=target.cuda(), **kw)
    v0 = model.get_tap("rtc_v_step0").cpu().view(plain_ref.shape)
    vjp0 = model.get_tap("rtc_vjp_step0").cpu().view(plain_ref.shape)
    model.set_taps(False)
    with torch.no_grad():
        prefix_pad, cache = O.prefill(params, oc, *args[:4])
    x = b["noise"].clone().requires_grad_(True)
    tb = torch.ones(2)
    v_ref = O.denoise_step(params, oc, prefix_pad, cache, x, tb)
    prev = torch.cat([target, torch.zeros(*target.shape[:-1], oc.action_dim - 14)], dim=-1)
    w = R.get_prefix_weights(2, 6, oc.action_horizon, "linear")[None, :, None]
    dm = (torch.arange(oc.action_dim) < 14).float()[None, None, :]
    err0 = (prev - (x - 1.0 * v_ref).detach()) * w * dm
    (vjp_ref,) = torch.autograd.grad(v_ref, x, grad_outputs=err0)
    e_v, e_j = H.rel_err(v0, v_ref), H.rel_err(vjp0, vjp_ref)
    print(f"[rtc {name}] step 0: velocity {e_v:.3e}, vector-Jacobian product {e_j:.3e}")
    assert e_v < 4e-3 and e_j < 5e-2


def test_rtc_reduces_to_the_plain_sampler_and_validates_inputs():
    oc, model, params, b, args = _setup("tiny")
    obs = H.Obs(b, "cuda")
    noise = b["noise"].cuda()
    plain = model.sample_actions("cuda", obs, noise=noise)
    assert torch.equal(model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=None), plain)
    assert torch.equal(model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=plain, enable_rtc=False), plain)
    # zero guidance weight: the guided loop is the plain Euler loop (same steps; the JAX sampler takes exactly num_steps)
    zero = model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=plain[..., :14], max_guidance_weight=0.0)
    assert H.rel_err(zero, plain) < 1e-3
    # NaNs from the client stream are zeroed (pi0_rtc.py:317); a wrong horizon is refused
    bad = plain[..., :14].clone()
    bad[0, 0, 0] = float("nan")
    assert bool(torch.isfinite(model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=bad)).all())
    with pytest.raises(ValueError):
        model.sample_actions("cuda", obs, noise=noise, prev_action_chunk=plain[:, :3])
