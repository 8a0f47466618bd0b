"""Backward determinism probe: same model twice, and two identically-seeded models; report which gradients differ."""
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as Hh  # noqa: E402
from oracle import pi05_oracle as O  # noqa: E402


def grads_of(model, obs, args):
    model.train()
    model.zero_grad(set_to_none=True)
    model(obs, *args).mean().backward()
    torch.cuda.synchronize()
    return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


def report(tag, g1, g2):
    groups = {}
    for n in g1:
        e = Hh.rel_err(g1[n], g2[n]) if float(g2[n].float().norm()) > 0 else float(g1[n].float().abs().max())
        key = re.sub(r"\.\d+\.", ".N.", n)
        groups[key] = max(groups.get(key, 0.0), e)
    bad = {k: v for k, v in groups.items() if v > 0}
    print(f"{tag}: {len(bad)} of {len(groups)} groups differ")
    for k, v in sorted(bad.items(), key=lambda kv: -kv[1])[:14]:
        print(f"   {v:.3e}  {k}")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    oc = O.tiny_config() if which == "tiny" else Hh.mid_config()
    m1, _ = Hh.build_pair(oc, seed=5)
    m2, _ = Hh.build_pair(oc, seed=5)
    batch = O.synthetic_batch(oc, 2)
    obs = Hh.Obs(batch, "cuda")
    args = (batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda())
    a = grads_of(m1, obs, args)
    b = grads_of(m1, obs, args)
    c = grads_of(m2, obs, args)
    report("same model, run 1 vs run 2", a, b)
    report("model 1 vs identically seeded model 2", a, c)
    m1.direct_grads = True
    d = grads_of(m1, obs, args)
    report("autograd vs direct_grads", a, d)


if __name__ == "__main__":
    main()
