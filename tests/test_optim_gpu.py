"""GPU parity of the fused clip + AdamW step against the caller-side torch calls it replaces
(scripts/train_pytorch.py:557-560): clip_grad_norm_(max_norm=1.0) + torch.optim.AdamW(betas=(0.9, 0.95), eps=1e-8,
weight_decay=1e-10) with state in the parameter dtype."""
import pytest
import torch

import helpers as H
from oracle import pi05_oracle as O

pytestmark = pytest.mark.gpu


def test_fused_clip_adamw_matches_torch():
    from kai0_b200.optim import FusedClipAdamW

    oc = O.tiny_config()
    model, _ = H.build_pair(oc, seed=5)
    ref, _ = H.build_pair(oc, seed=5)
    for m in (model, ref):
        m.direct_grads = True
        m.train()
    batch = O.synthetic_batch(oc, 2)
    obs = H.Obs(batch, "cuda")
    args = (batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda())
    opt = FusedClipAdamW(model, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-2, max_norm=1.0)
    rparams = ref.flat_parameters()
    ropt = torch.optim.AdamW(rparams, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-2)
    for it in range(3):
        model(obs, *args).mean().backward()
        ref(obs, *args).mean().backward()
        # both replicas hold the same weights at this point (re-synchronised below), so the gradients agree up to the
        # fp32 atomics order of the bias / norm-weight reductions
        assert H.rel_err(model.flat_parameters()[0].grad, rparams[0].grad) < 1e-5
        norm = opt.step()
        rnorm = torch.nn.utils.clip_grad_norm_(rparams, max_norm=1.0)
        ropt.step()
        opt.zero_grad()
        ropt.zero_grad(set_to_none=True)
        assert abs(float(norm) - float(rnorm)) / float(rnorm) < 5e-3  # torch rounds the bf16 arena's norm to bf16
        for a, b in zip(model.flat_parameters(), rparams):
            # same update rule; torch clips with a bf16-rounded norm, so allow a few bf16 ulps on the moved weights
            assert H.rel_err(a, b) < 2e-3, it
        # compare every step in isolation: copy weights and moments of the fused optimiser into the torch replica
        with torch.no_grad():
            for i, (a, b) in enumerate(zip(model.flat_parameters(), rparams)):
                b.copy_(a)
                ropt.state[b]["exp_avg"].copy_(opt.m[i])
                ropt.state[b]["exp_avg_sq"].copy_(opt.v[i])
    # the update actually moved the weights and stayed finite
    assert torch.isfinite(model.flat_parameters()[0].float()).all()
    w0 = H.build_pair(oc, seed=5, device="cuda")[0].flat_parameters()[1]
    assert float((model.flat_parameters()[1] - w0).abs().max()) > 0


def test_fused_adamw_without_clipping_is_elementwise_exact_enough():
    """No clipping (max_norm=0): the only differences to torch's fused kernel are fp32 rounding inside the update."""
    from kai0_b200 import _lib
    import ctypes as C

    n = 1 << 16
    g = torch.Generator(device="cuda").manual_seed(0)
    p = (torch.randn(n, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    grad = (torch.randn(n, device="cuda", generator=g) * 1e-3).to(torch.bfloat16)
    pf = torch.randn(1024, device="cuda", generator=g)
    gf = torch.randn(1024, device="cuda", generator=g) * 1e-3
    rp, rpf = torch.nn.Parameter(p.clone()), torch.nn.Parameter(pf.clone())
    ropt = torch.optim.AdamW([rp, rpf], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10, fused=True)
    m, v, mf, vf = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(pf), torch.zeros_like(pf)
    scratch = torch.zeros(4096, device="cuda")
    for step in range(1, 4):
        rp.grad, rpf.grad = grad.clone(), gf.clone()
        ropt.step()
        rc = _lib.lib().pi05_fused_clip_adamw(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n,
                                              pf.data_ptr(), gf.data_ptr(), mf.data_ptr(), vf.data_ptr(), 1024, 1e-2, 0.9,
                                              0.95, 1e-8, 1e-10, step, 0.0, scratch.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        assert H.rel_err(pf, rpf) < 1e-6
        # bf16 storage: identical except for rare 1-ulp rounding flips
        diff = (p.float() - rp.detach().float()).abs()
        assert float((diff > 0).float().mean()) < 0.02 and H.rel_err(p, rp) < 1e-3
