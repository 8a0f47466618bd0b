"""Run sample_actions / forward twice on identical inputs and report the first intermediate that differs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as Hh  # noqa: E402
from oracle import pi05_oracle as O  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "mid"
    if which == "full":
        from kai0_b200.pi0_pytorch import PI0Pytorch, Pi05EngineConfig
        oc = O.OracleConfig()
        model = PI0Pytorch(Pi05EngineConfig(), init_weights=False).to("cuda")
        model.augment = False
        model.reset_parameters(seed=7)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "layernorm" in n or n.endswith("model.norm.weight") or "norm.dense" in n:
                    p.normal_(0.0, 0.02)
    else:
        oc = O.tiny_config() if which == "tiny" else Hh.mid_config()
        model, _ = Hh.build_pair(oc, seed=0)
    B = 1
    batch = O.synthetic_batch(oc, B)
    obs = Hh.Obs(batch, "cuda")
    model.set_taps(True)
    model.eval()
    noise = batch["noise"].cuda()
    names = ["vit_embed"] + [f"vit_layer{l}" for l in range(oc.vit_depth)] + ["prefix_embs", "suffix_embs", "adarms_cond"]
    names += [x for l in range(oc.paligemma.depth) for x in (f"layer{l}_prefix", f"layer{l}_suffix")]
    names += ["prefix_out", "suffix_out", "v_t"]
    runs = []
    for r in range(3):
        with torch.no_grad():
            loss = model(obs, batch["actions"].cuda(), noise, batch["time"].cuda())
        taps = {n: model.get_tap(n).clone() for n in names}
        taps["loss"] = loss.clone()
        runs.append(taps)
    for n in names + ["loss"]:
        same = all(torch.equal(runs[0][n], runs[r][n]) for r in (1, 2))
        if not same:
            d = (runs[0][n].float() - runs[1][n].float()).abs()
            print(f"forward: FIRST DIFFERENCE at {n}: max {float(d.max()):.3e} count {int((d > 0).sum())}/{d.numel()}")
            break
    else:
        print("forward: 3 runs bit-identical")
    acts = [model.sample_actions("cuda", obs, noise=noise).clone() for _ in range(3)]
    v0 = []
    for r in range(2):
        model.sample_actions("cuda", obs, noise=noise)
        v0.append(model.get_tap("v_t_step0").clone())
    print("decode: actions identical:", all(torch.equal(acts[0], a) for a in acts[1:]),
          " v_t_step0 identical:", torch.equal(v0[0], v0[1]))


if __name__ == "__main__":
    main()
