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
        del m
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
    path = os.path.join(PIN.ROOT, "tests", "golden", "reference_pin.pt")
    torch.save(out, path)
    print("wrote", path)


if __name__ == "__main__":
    main()
