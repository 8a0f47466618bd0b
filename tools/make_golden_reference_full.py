"""Golden outputs of the REFERENCE'S OWN PI0Pytorch at BASELINE.json's FULL architecture (SigLIP-So400m x 27 layers,
PaliGemma gemma_2b x 18, action expert gemma_300m x 18, vocabulary 257152, prompt length 200, 3 cameras 224 x 224,
horizon 50, action dim 32), executed in place from /root/reference through tools/reference_loader.py on the CPU of the
build container.  Outputs only are committed (tests/golden/reference_full.pt); weights and inputs are regenerated from
seeds (oracle.init_params(OracleConfig(), FULL_SEED), oracle.synthetic_batch) wherever the fixture is used.

What is recorded, per reference precision ("bfloat16" = the reference's training / serving dtype map, "float32"):
  * `forward` loss tensor [2, 50, 32] with injected noise / time, preprocessing in eval mode (B = 2: sample 0 has three
    cameras, sample 1 has the two wrist cameras masked out = the "one-camera" variant of SURVEY §8d, ragged prompts);
  * `sample_actions` (10 Euler steps) action chunk at B = 1 for each of those two samples (configs[0] / configs[3]) and
    at B = 2 (bfloat16 only: the reference against itself at another batch shape = its own run-to-run floor);
  * `loss.mean().backward()`: per-parameter gradient norm, abs-max and 256 strided elements (bfloat16; float32 when
    --grads-f32 is given: needs ~50 GB of host memory).

    python tools/make_golden_reference_full.py --precision bfloat16
    python tools/make_golden_reference_full.py --precision float32
"""
import argparse
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reference_loader as RL  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import pi05_oracle as O  # noqa: E402

FULL_SEED = 20250924
INPUT_SEED = 79
KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
PATH = os.path.join(ROOT, "tests", "golden", "reference_full.pt")


def full_inputs():
    oc = O.OracleConfig()
    b = O.synthetic_batch(oc, 2, seed=INPUT_SEED, ragged=True)
    b["img_masks"][1][1] = False  # sample 1: only base_0_rgb is a real camera (Libero convention, libero_policy.py:59-68)
    b["img_masks"][2][1] = False
    return b


def full_weights(dtype_map: bool):
    oc = O.OracleConfig()
    p = O.init_params(oc, FULL_SEED)
    if not dtype_map:
        p = {k: v.to(torch.float32) for k, v in p.items()}  # bf16-representable values held in fp32 (same weights)
    return p


class Obs:
    def __init__(self, b, rows=None):
        sl = (lambda t: t) if rows is None else (lambda t: t[rows])
        self.images = {k: sl(b["images"][i]) for i, k in enumerate(KEYS)}
        self.image_masks = {k: sl(b["img_masks"][i]) for i, k in enumerate(KEYS)}
        self.state = torch.zeros(sl(b["tokens"]).shape[0], 32)
        self.tokenized_prompt = sl(b["tokens"])
        self.tokenized_prompt_mask = sl(b["token_mask"])
        self.token_ar_mask = None
        self.token_loss_mask = None


def grad_summary(named_grads):
    out = {}
    for name, g in named_grads:
        if g is None:
            continue
        f = g.detach().to(torch.float32).reshape(-1)
        k = min(256, f.numel())
        idx = (torch.arange(k, dtype=torch.int64) * (f.numel() - 1)) // max(k - 1, 1)
        out[name] = {"norm": float(f.norm()), "absmax": float(f.abs().max()), "sample": f[idx].clone()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", choices=("bfloat16", "float32"), required=True)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--grads-f32", action="store_true")
    ap.add_argument("--skip-grads", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    from transformers.initialization import no_init_weights

    prec = args.precision
    t0 = time.time()
    p0 = RL.load(vision_layers=None)
    cfg = types.SimpleNamespace(pi05=True, paligemma_variant="gemma_2b", action_expert_variant="gemma_300m", dtype=prec,
                                action_horizon=50, action_dim=32, max_token_len=200)
    with no_init_weights():
        m = p0.PI0Pytorch(cfg)
    torch.set_float32_matmul_precision("highest")
    params = full_weights(dtype_map=prec == "bfloat16")
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected and all("lm_head" in k for k in missing), (missing, unexpected)
    del params
    m.eval()
    print(f"[{prec}] reference built + weights loaded in {time.time() - t0:.0f} s", flush=True)
    b = full_inputs()
    pp = sys.modules["openpi.models_pytorch.preprocessing_pytorch"]
    orig = pp.preprocess_observation_pytorch
    pp.preprocess_observation_pytorch = lambda o, train=False, **k: orig(o, train=False, **k)
    out = torch.load(PATH) if os.path.exists(PATH) else {}
    out.update({"weight_seed": FULL_SEED, "input_seed": INPUT_SEED})
    with torch.no_grad():
        t = time.time()
        loss = m.forward(Obs(b), b["actions"], b["noise"], b["time"])
        out[f"loss_{prec}"] = loss.to(torch.float32).contiguous()
        print(f"[{prec}] forward B=2: loss mean {float(loss.mean()):.6f}  [{time.time() - t:.0f} s]", flush=True)
        for r in (0, 1):
            t = time.time()
            a = m.sample_actions("cpu", Obs(b, slice(r, r + 1)), noise=b["noise"][r:r + 1], num_steps=10)
            out[f"actions_b1_row{r}_{prec}"] = a.to(torch.float32).contiguous()
            print(f"[{prec}] sample_actions B=1 row {r}: |a| {float(a.norm()):.4f}  [{time.time() - t:.0f} s]", flush=True)
        if prec == "bfloat16":
            t = time.time()
            a2 = m.sample_actions("cpu", Obs(b), noise=b["noise"], num_steps=10)
            out[f"actions_b2_{prec}"] = a2.to(torch.float32).contiguous()
            for r in (0, 1):
                d = a2[r:r + 1].float() - out[f"actions_b1_row{r}_{prec}"]
                print(f"[{prec}] reference vs itself, B=2 vs B=1, row {r}: rel "
                      f"{float(d.norm() / out[f'actions_b1_row{r}_{prec}'].norm()):.3e}  [{time.time() - t:.0f} s]",
                      flush=True)
    torch.save(out, PATH)
    if not args.skip_grads and (prec == "bfloat16" or args.grads_f32):
        t = time.time()
        for p in m.parameters():
            p.grad = None
        loss = m.forward(Obs(b), b["actions"], b["noise"], b["time"])
        loss.mean().backward()
        out[f"grads_{prec}"] = {k: v for k, v in grad_summary((n, p.grad) for n, p in m.named_parameters()).items()
                                if "lm_head" not in k}
        print(f"[{prec}] backward: {len(out[f'grads_{prec}'])} parameter gradients  [{time.time() - t:.0f} s]", flush=True)
        torch.save(out, PATH)
    pp.preprocess_observation_pytorch = orig
    print("wrote", PATH, sorted(out))


if __name__ == "__main__":
    main()
