"""GPU parity of the AdvantageEstimator variant (pi0_pytorch.py:464-644) through the
reference-facing class -> C-ABI (pi05_forward_advantage / pi05_backward / pi05_forward_value) against the CPU oracle.
Tolerances are those of tests/test_engine_gpu.py (bf16 noise floor of the shared backbone)."""
import pytest
import torch

import helpers as H
from oracle import pi05_oracle as O

pytestmark = pytest.mark.gpu

TOL_LOSS = 4e-3
TOL_VALUE = 4e-3
TOL_GRAD = 3e-2

# insertion order deliberately scrambled: the class must sort by (timestep, part) like preprocessing_pytorch.py:196-204
KEYS6 = ["right_wrist_0_rgb", "base_-100_rgb", "left_wrist_0_rgb", "base_0_rgb", "right_wrist_-100_rgb",
         "left_wrist_-100_rgb"]
SORTED6 = ["base_-100_rgb", "left_wrist_-100_rgb", "right_wrist_-100_rgb", "base_0_rgb", "left_wrist_0_rgb",
           "right_wrist_0_rgb"]


class AdvObs:
    def __init__(self, batch, keys_sorted, insertion, progress, device="cuda"):
        by_key = {k: i for i, k in enumerate(keys_sorted)}
        self.images = {k: batch["images"][by_key[k]].to(device) for k in insertion}
        self.image_masks = {k: batch["img_masks"][by_key[k]].to(device) for k in insertion}
        B = batch["tokens"].shape[0]
        self.state = torch.zeros(B, 32, device=device)
        self.tokenized_prompt = batch["tokens"].to(device)
        self.tokenized_prompt_mask = batch["token_mask"].to(device)
        self.token_ar_mask = self.token_loss_mask = None
        self.progress = progress.to(device) if progress is not None else None
        self.frame_index = self.episode_length = self.image_original = self.episode_index = None


def _pair(name, ni):
    from kai0_b200.pi0_pytorch import AdvantageEstimator

    kw = dict(num_images=ni, value_head=True)
    oc = O.tiny_config(**kw) if name == "tiny" else H.mid_config(**kw)
    model, params = H.build_pair(oc, seed=7, cls=AdvantageEstimator)
    return oc, model, params


@pytest.mark.parametrize("name,ni,B,wa,wv", [("tiny", 4, 3, 0.7, 1.3), ("mid", 6, 2, 0.0, 1.0)])
def test_advantage_forward_backward_match_oracle(name, ni, B, wa, wv):
    oc, model, params = _pair(name, ni)
    model.loss_action_weight, model.loss_value_weight = wa, wv
    keys_sorted = SORTED6 if ni == 6 else ["base_-100_rgb", "right_wrist_-100_rgb", "base_0_rgb", "left_wrist_0_rgb"]
    insertion = KEYS6 if ni == 6 else list(reversed(keys_sorted))
    batch = O.synthetic_batch(oc, B, ragged=True)
    batch["img_masks"][1][0] = False
    progress = torch.tensor([0.35, -1.7, 0.9][:B])  # one target outside [-1, 1]: exercises the clamp
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss_ref = O.advantage_forward_loss(pr, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                                        batch["actions"], batch["noise"], batch["time"], progress,
                                        loss_action_weight=wa, loss_value_weight=wv)
    loss_ref.mean().backward()
    obs = AdvObs(batch, keys_sorted, insertion, progress)
    model.train()
    loss, aux = model(obs, batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda(), return_loss_dict=True)
    assert loss.shape == (B, oc.action_horizon) and loss.dtype == torch.float32
    loss.mean().backward()
    torch.cuda.synchronize()
    assert H.rel_err(loss, loss_ref) < TOL_LOSS
    # loss_aux_dict (pi0_pytorch.py:582-583)
    with torch.no_grad():
        v_t, so = O.model_v_t(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                              batch["time"][:, None, None] * batch["noise"] + (1 - batch["time"][:, None, None]) * batch["actions"],
                              batch["time"])
        la = torch.nn.functional.mse_loss(batch["noise"] - batch["actions"], v_t, reduction="none").mean(-1).mean()
        value_ref = O.value_head(params, so)
        lv = ((value_ref - progress.clamp(-1, 1)[:, None]) ** 2 * wv).mean()
    assert abs(float(aux["loss_action"]) - float(la)) <= TOL_LOSS * max(float(la), 1e-6)
    assert abs(float(aux["loss_value"]) - float(lv)) <= 2e-2 * max(float(lv), 1e-6)
    worst = {}
    for n, p in model.named_parameters():
        if n not in pr:
            assert p.grad is None
            continue
        if pr[n].grad is None:  # unreachable from the loss: no gradient on either side (pi0_pytorch.py:350-358)
            assert p.grad is None and n in model._dead_grad_names, n
            continue
        gr = pr[n].grad
        assert p.grad is not None and p.grad.shape == p.shape, n
        if n.endswith("self_attn.k_proj.bias") and "vision_tower" in n:
            # mathematically zero (softmax is invariant to a per-query constant): both sides hold bf16 rounding noise
            qn = float(pr[n.replace("k_proj", "q_proj")].grad.float().norm())
            assert float(p.grad.float().norm()) < 0.05 * qn and float(gr.float().norm()) < 0.05 * qn, n
            continue
        if float(gr.float().norm()) < 1e-5:
            assert float(p.grad.float().abs().max()) < 1e-4, n
            continue
        worst[n] = H.rel_err(p.grad, gr)
    bad = {k: v for k, v in worst.items() if not v < TOL_GRAD}
    assert not bad, bad
    assert any(k.startswith("value_head.") for k in worst)


def test_sample_values_matches_oracle_and_public_call():
    oc, model, params = _pair("mid", 6)
    B = 2
    batch = O.synthetic_batch(oc, B)
    obs = AdvObs(batch, SORTED6, KEYS6, None)
    model.eval()
    images, img_masks, toks, tmask, _ = model._preprocess_observation(obs, train=False)
    val = model._sample_values(images, img_masks, toks, tmask, batch["noise"].cuda(), batch["time"].cuda())
    with torch.no_grad():
        _, so = O.model_v_t(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                            batch["noise"], batch["time"])
        ref = O.value_head(params, so)
    assert val.shape == (B, 1) and H.max_err(val, ref) < TOL_VALUE
    torch.manual_seed(0)
    v2 = model.sample_values("cuda", obs)  # noise / time drawn inside as the reference does (:604-605)
    assert v2.shape == (B, 1) and bool(((v2 > -1) & (v2 < 1)).all())


def test_plain_pi0_loss_on_the_three_camera_keys_still_works_after_six_image_call():
    """The engine is re-planned when the number of images changes; PI0 decode on the same class keeps working."""
    oc, model, params = _pair("tiny", 6)
    batch = O.synthetic_batch(oc, 2)
    obs = AdvObs(batch, SORTED6, KEYS6, torch.zeros(2))
    acts = model.sample_actions("cuda", obs, noise=batch["noise"].cuda())
    ref = O.sample_actions(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                           batch["noise"])
    assert H.rel_err(acts, ref) < 1e-3
    oc3 = O.tiny_config(num_images=3, value_head=True)
    b3 = O.synthetic_batch(oc3, 2)
    obs3 = AdvObs(b3, SORTED6[3:], SORTED6[3:], torch.zeros(2))
    acts3 = model.sample_actions("cuda", obs3, noise=b3["noise"].cuda())
    ref3 = O.sample_actions(params, oc3, b3["images"], b3["img_masks"], b3["tokens"], b3["token_mask"], b3["noise"])
    assert H.rel_err(acts3, ref3) < 1e-3
