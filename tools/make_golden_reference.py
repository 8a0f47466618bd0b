"""Golden outputs of the REFERENCE'S OWN PI0Pytorch (src/openpi/models_pytorch/pi0_pytorch.py, executed in place from
/root/reference through tools/reference_loader.py) on the pin configuration of tools/reference_pin.py, in both of the
reference's precisions ("bfloat16" dtype map and "float32").  Also prints how far oracle/pi05_oracle.py is from it.
Build container only; writes tests/golden/reference_pin.pt (outputs only: weights and inputs are regenerated from
seeds).

    python tools/make_golden_reference.py
"""
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reference_loader as RL  # noqa: E402
import reference_pin as PIN  # noqa: E402
from oracle import pi05_oracle as O  # noqa: E402


class Obs:
    def __init__(self, b):
        self.images = {k: b["images"][i] for i, k in enumerate(PIN.KEYS)}
        self.image_masks = {k: b["img_masks"][i] for i, k in enumerate(PIN.KEYS)}
        self.state = torch.zeros(b["tokens"].shape[0], 32)
        self.tokenized_prompt = b["tokens"]
        self.tokenized_prompt_mask = b["token_mask"]
        self.token_ar_mask = None
        self.token_loss_mask = None


def build_reference(precision: str):
    from transformers.initialization import no_init_weights

    RL.register_variant("pin_pg", RL.SizeRecord(*PIN.PG))
    RL.register_variant("pin_ex", RL.SizeRecord(*PIN.EX))
    p0 = RL.load(vision_layers=PIN.VIT_LAYERS)
    cfg = types.SimpleNamespace(pi05=True, paligemma_variant="pin_pg", action_expert_variant="pin_ex", dtype=precision,
                                action_horizon=50, action_dim=32, max_token_len=PIN.MAX_TOKEN_LEN)
    with no_init_weights():
        m = p0.PI0Pytorch(cfg)
    torch.set_float32_matmul_precision("highest")  # the constructor sets "high" (TF32 on GPUs); keep the CPU run exact
    m.eval()
    return p0, m


def run_reference(p0, m, params, b):
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected and all("lm_head" in k for k in missing), (missing, unexpected)
    # forward() always preprocesses with train=True (augmentation); the network comparison needs it off
    pp = sys.modules["openpi.models_pytorch.preprocessing_pytorch"]
    orig = pp.preprocess_observation_pytorch
    pp.preprocess_observation_pytorch = lambda o, train=False, **k: orig(o, train=False, **k)
    try:
        with torch.no_grad():
            loss = m.forward(Obs(b), b["actions"], b["noise"], b["time"])
            acts = m.sample_actions("cpu", Obs(b), noise=b["noise"], num_steps=10)
    finally:
        pp.preprocess_observation_pytorch = orig
    return loss, acts


def grad_summary(named_grads):
    """Per parameter: L2 norm, abs-max and 256 evenly strided elements of the gradient (the full gradients are 2.6 GB)."""
    out = {}
    for name, g in named_grads:
        if g is None:
            continue
        f = g.detach().to(torch.float32).reshape(-1)
        k = min(256, f.numel())
        idx = (torch.arange(k, dtype=torch.int64) * (f.numel() - 1)) // max(k - 1, 1)
        out[name] = {"norm": float(f.norm()), "absmax": float(f.abs().max()), "sample": f[idx].clone()}
    return out


def run_reference_backward(p0, m, params, b):
    """loss.mean().backward() through the reference (train_pytorch.py:547-549); returns the gradient summaries."""
    m.load_state_dict(params, strict=False)
    pp = sys.modules["openpi.models_pytorch.preprocessing_pytorch"]
    orig = pp.preprocess_observation_pytorch
    pp.preprocess_observation_pytorch = lambda o, train=False, **k: orig(o, train=False, **k)
    try:
        for p in m.parameters():
            p.grad = None
        loss = m.forward(Obs(b), b["actions"], b["noise"], b["time"])
        loss.mean().backward()
    finally:
        pp.preprocess_observation_pytorch = orig
    return grad_summary((n, p.grad) for n, p in m.named_parameters())


def oracle_backward(params, oc, b):
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss = O.forward_loss(pr, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"], b["noise"],
                          b["time"])
    loss.mean().backward()
    return grad_summary((n, p.grad) for n, p in pr.items())


def compare_grads(mine, ref):
    """Worst relative difference over parameters (sampled elements and norm).  Skipped: gradients that are
    mathematically zero and hold only rounding noise on both sides — the SigLIP key biases (softmax is invariant to a
    per-query constant) and tensors whose reference norm is < 1e-9 of the largest gradient norm."""
    top = max(r["norm"] for r in ref.values())
    errs = []
    for name, r in ref.items():
        if name not in mine or r["norm"] < 1e-9 * top:
            continue
        if "vision_tower" in name and name.endswith("self_attn.k_proj.bias"):
            continue
        e = float((mine[name]["sample"] - r["sample"]).norm() / max(float(r["sample"].norm()), 1e-30))
        e = max(e, abs(mine[name]["norm"] - r["norm"]) / r["norm"])
        errs.append((e, name))
    errs.sort(reverse=True)
    return errs[0] if errs else (0.0, None)


ADV_KEYS = ("base_-100_rgb", "left_wrist_-100_rgb", "right_wrist_-100_rgb", "base_0_rgb", "left_wrist_0_rgb",
            "right_wrist_0_rgb")  # the order preprocess_observation_pytorch_custom sorts into (:196-204)
ADV_WA, ADV_WV = 0.7, 1.3
ADV_PROGRESS = (0.35, -1.7)  # one target outside [-1, 1]: exercises the clamp (pi0_pytorch.py:574)


class AdvObs:
    def __init__(self, b, progress):
        order = (3, 0, 5, 1, 4, 2)  # scrambled insertion order: the reference must sort by (timestep, part)
        self.images = {ADV_KEYS[i]: b["images"][i] for i in order}
        self.image_masks = {ADV_KEYS[i]: b["img_masks"][i] for i in order}
        n = b["tokens"].shape[0]
        self.state = torch.zeros(n, 32)
        self.tokenized_prompt = b["tokens"]
        self.tokenized_prompt_mask = b["token_mask"]
        self.progress = progress
        self.token_ar_mask = self.token_loss_mask = self.frame_index = self.episode_length = None
        self.image_original = self.episode_index = None


def adv_config_and_inputs():
    import dataclasses

    oc = dataclasses.replace(PIN.oracle_config(), num_images=6, value_head=True)
    b = O.synthetic_batch(oc, PIN.BATCH, seed=78, ragged=True)
    b["img_masks"][4][1] = False
    return oc, b, torch.tensor(ADV_PROGRESS)


def run_reference_advantage(p0, precision, params, b, progress):
    from transformers.initialization import no_init_weights

    cfg = types.SimpleNamespace(pi05=True, paligemma_variant="pin_pg", action_expert_variant="pin_ex", dtype=precision,
                                action_horizon=50, action_dim=32, max_token_len=PIN.MAX_TOKEN_LEN,
                                loss_value_weight=ADV_WV, loss_action_weight=ADV_WA)
    with no_init_weights():
        m = p0.AdvantageEstimator(cfg)
    torch.set_float32_matmul_precision("highest")
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected and all("lm_head" in k for k in missing), (missing, unexpected)
    m.train()  # AdvantageEstimator.forward preprocesses with train=self.training but apply_aug=False (:488-489)
    loss, aux = m.forward(AdvObs(b, progress), b["actions"], b["noise"], b["time"], return_loss_dict=True)
    for p in m.parameters():
        p.grad = None
    loss.mean().backward()
    grads = grad_summary((n, p.grad) for n, p in m.named_parameters() if "lm_head" not in n)
    m.eval()
    m.sample_noise = lambda shape, device: b["noise"]  # sample_values draws these itself (:604-605): inject ours
    m.sample_time = lambda bsize, device: b["time"]
    with torch.no_grad():
        value = m.sample_values("cpu", AdvObs(b, None))
    return loss.detach(), {k: float(v) for k, v in aux.items()}, value, grads


def oracle_advantage(params, oc, b, progress, with_grads=True):
    pr = {k: v.clone().requires_grad_(with_grads) for k, v in params.items()}
    with torch.set_grad_enabled(with_grads):
        loss = O.advantage_forward_loss(pr, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"],
                                        b["noise"], b["time"], progress, loss_action_weight=ADV_WA,
                                        loss_value_weight=ADV_WV)
    grads = None
    if with_grads:
        loss.mean().backward()
        grads = grad_summary((n, p.grad) for n, p in pr.items())
    with torch.no_grad():
        _, so = O.model_v_t(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"], b["time"])
        value = O.value_head(params, so)
    return loss.detach(), value, grads


def main():
    torch.set_num_threads(int(os.environ.get("PIN_THREADS", "2")))
    oc = PIN.oracle_config()
    specs = {k: v for k, v in O.param_specs(oc).items()}
    b = PIN.pin_inputs()
    out = {"weight_seed": PIN.WEIGHT_SEED, "pg": PIN.PG, "ex": PIN.EX, "vit_layers": PIN.VIT_LAYERS,
           "max_token_len": PIN.MAX_TOKEN_LEN, "batch": PIN.BATCH}
    for precision in ("bfloat16", "float32"):
        t = time.time()
        params = PIN.pin_weights(specs, dtype_map=precision == "bfloat16")
        p0, m = build_reference(precision)
        loss, acts = run_reference(p0, m, params, b)
        grads = run_reference_backward(p0, m, params, b)
        del m
        o_grads = oracle_backward(params, oc, b)
        assert set(k for k in grads if "lm_head" not in k) == set(o_grads), "gradient key sets differ"
        w = compare_grads(o_grads, grads)
        print(f"{precision:9s}: gradients of {len(o_grads)} parameters, oracle autograd vs reference autograd: worst "
              f"relative difference {w[0]:.3e} ({w[1]})")
        out[f"grads_{precision}"] = {k: v for k, v in grads.items() if "lm_head" not in k}
        with torch.no_grad():
            o_loss = O.forward_loss(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"],
                                    b["noise"], b["time"])
            o_acts = O.sample_actions(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"])
        rel = lambda a, r: float((a - r).norm() / r.norm())  # noqa: E731
        print(f"{precision:9s}: reference loss mean {float(loss.mean()):.6f}; oracle vs reference: loss rel "
              f"{rel(o_loss, loss):.3e} (max abs {float((o_loss - loss).abs().max()):.3e}), actions rel "
              f"{rel(o_acts, acts):.3e}; bit-equal loss {torch.equal(o_loss, loss)} actions {torch.equal(o_acts, acts)} "
              f"[{time.time() - t:.0f} s]")
        out[f"loss_{precision}"] = loss.to(torch.float32).contiguous()
        out[f"actions_{precision}"] = acts.to(torch.float32).contiguous()
    # ---- the DEFAULT training forward: preprocessing with train=True (augmentation) and noise / time drawn inside, all
    # from the global torch RNG (pi0_pytorch.py:318-324) -> pins the order in which random numbers are consumed
    params = PIN.pin_weights(specs, dtype_map=False)
    p0, m = build_reference("float32")
    m.load_state_dict(params, strict=False)
    torch.manual_seed(PIN.WEIGHT_SEED + 1)
    with torch.no_grad():
        out["loss_default_float32"] = m.forward(Obs(b), b["actions"]).to(torch.float32).contiguous()
    del m
    print(f"default forward (augmentation + internal noise/time, seed {PIN.WEIGHT_SEED + 1}): loss mean "
          f"{float(out['loss_default_float32'].mean()):.6f}")
    # ---- AdvantageEstimator (pi0_pytorch.py:464-644): 6 images, value head, weighted loss, sample_values
    oc_a, b_a, progress = adv_config_and_inputs()
    for precision in ("bfloat16", "float32"):
        t = time.time()
        params = PIN.pin_weights(O.param_specs(oc_a), dtype_map=precision == "bfloat16")
        p0 = RL.load(vision_layers=PIN.VIT_LAYERS)
        loss, aux, value, grads = run_reference_advantage(p0, precision, params, b_a, progress)
        o_loss, o_value, o_grads = oracle_advantage(params, oc_a, b_a, progress)
        rel = lambda a, r: float((a - r).norm() / r.norm())  # noqa: E731
        assert set(grads) == set(o_grads), "advantage gradient key sets differ"
        w = compare_grads(o_grads, grads)
        print(f"{precision:9s}: AdvantageEstimator loss {tuple(loss.shape)} rel {rel(o_loss, loss):.3e}, value max abs "
              f"{float((o_value - value).abs().max()):.3e}, aux {aux}, gradients worst {w[0]:.3e} ({w[1]}) "
              f"[{time.time() - t:.0f} s]")
        out[f"adv_loss_{precision}"] = loss.to(torch.float32).contiguous()
        out[f"adv_value_{precision}"] = value.to(torch.float32).contiguous()
        out[f"adv_aux_{precision}"] = aux
        out[f"adv_grads_{precision}"] = grads
    path = os.path.join(PIN.ROOT, "tests", "golden", "reference_pin.pt")
    torch.save(out, path)
    print("wrote", path)


if __name__ == "__main__":
    main()
