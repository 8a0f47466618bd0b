"""Gradient parity probe (GPU): engine backward vs torch.autograd through the CPU oracle.

    python tools/grad_probe.py [tiny|mid] [batch]
"""
import os
import re
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as Hh  # noqa: E402
from oracle import pi05_oracle as O  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    oc = O.tiny_config() if which == "tiny" else Hh.mid_config()
    model, params = Hh.build_pair(oc, seed=0)
    batch = O.synthetic_batch(oc, B, ragged=True)
    batch["tokens"][0, 1] = batch["tokens"][0, 0]  # a repeated token id (scatter-add path)
    if B > 1:
        batch["img_masks"][1][B - 1] = False
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    t0 = time.time()
    loss_ref = O.forward_loss(pr, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                              batch["actions"], batch["noise"], batch["time"])
    loss_ref.mean().backward()
    print(f"oracle fwd+bwd {time.time() - t0:.2f}s  loss {loss_ref.mean().item():.6f}", flush=True)
    obs = Hh.Obs(batch, "cuda")
    model.train()
    loss = model(obs, batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda())
    loss.mean().backward()
    torch.cuda.synchronize()
    print(f"engine loss {loss.mean().item():.6f}  rel {Hh.rel_err(loss, loss_ref):.3e}", flush=True)
    groups = {}
    for name, p in model.named_parameters():
        if name not in pr:
            continue
        g, gr = p.grad, pr[name].grad
        key = re.sub(r"\.\d+\.", ".N.", name)
        if g is None:
            groups.setdefault(key, []).append((float("nan"), name, 0.0))
            continue
        if gr is None:
            gr = torch.zeros_like(pr[name])
        e = Hh.rel_err(g, gr) if float(gr.float().norm()) > 0 else float(g.float().abs().max())
        groups.setdefault(key, []).append((e, name, float(gr.float().norm())))
    for key, lst in groups.items():
        worst = max(lst, key=lambda t: (t[0] if t[0] == t[0] else 1e9))
        print(f"  {key:95s} worst rel={worst[0]:.3e} (|ref|={worst[2]:.3e}) n={len(lst)}", flush=True)


if __name__ == "__main__":
    main()
