"""Parity probe (GPU): engine vs CPU oracle on small configs, printing the error of every tap.

    python tools/parity_probe.py [tiny|mid] [batch]
"""
import os
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
    torch.manual_seed(0)
    model, params = Hh.build_pair(oc, seed=0)
    batch = O.synthetic_batch(oc, B, ragged=True)
    if B > 1:
        batch["img_masks"][1][B - 1] = False  # one masked camera (Libero convention)
    # ---- oracle
    t0 = time.time()
    taps = {}
    with torch.no_grad():
        loss_ref = O.forward_loss(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                                  batch["actions"], batch["noise"], batch["time"], taps)
    print(f"oracle forward {time.time() - t0:.2f}s", flush=True)
    # ---- engine
    obs = Hh.Obs(batch, "cuda")
    model.set_taps(True)
    model.eval()
    with torch.no_grad():
        loss = model(obs, batch["actions"].cuda(), batch["noise"].cuda(), batch["time"].cuda())
    torch.cuda.synchronize()
    T, P, A = oc.num_patches, oc.num_images * oc.num_patches + oc.max_token_len, oc.action_horizon
    pad = torch.cat([m[:, None].expand(B, T) for m in batch["img_masks"]] + [batch["token_mask"]], dim=1)  # [B,P]

    def cmp(name, got, ref, rowmask=None):
        got = got.float().cpu().reshape(ref.shape)
        ref = ref.float()
        if rowmask is not None:
            got = got[rowmask]
            ref = ref[rowmask]
        print(f"  {name:28s} rel={Hh.rel_err(got, ref):.3e} max={Hh.max_err(got, ref):.3e} "
              f"|ref|max={float(ref.abs().max()):.3e}", flush=True)

    ve = model.get_tap("vit_embed").view(oc.num_images, B, T, oc.vit_width)
    for n in range(oc.num_images):
        cmp(f"img{n}_vit_embed", ve[n], taps[f"img{n}_vit_embed"])
    for l in range(oc.vit_depth):
        x = model.get_tap(f"vit_layer{l}").view(oc.num_images, B, T, oc.vit_width)
        for n in range(oc.num_images):
            cmp(f"img{n}_vit_layer{l}", x[n], taps[f"img{n}_vit_layer{l}"])
    cmp("prefix_embs", model.get_tap("prefix_embs"), taps["prefix_embs"])
    cmp("suffix_embs", model.get_tap("suffix_embs"), taps["suffix_embs"])
    cmp("adarms_cond", model.get_tap("adarms_cond"), taps["adarms_cond"])
    for l in range(oc.paligemma.depth):
        cmp(f"layer{l}_prefix(valid rows)", model.get_tap(f"layer{l}_prefix"), taps[f"layer{l}_prefix"], pad)
        cmp(f"layer{l}_suffix", model.get_tap(f"layer{l}_suffix"), taps[f"layer{l}_suffix"])
    cmp("prefix_out(valid rows)", model.get_tap("prefix_out"), taps["prefix_out"], pad)
    cmp("suffix_out", model.get_tap("suffix_out"), taps["suffix_out"])
    cmp("v_t", model.get_tap("v_t"), taps["v_t"])
    cmp("loss", loss, loss_ref)
    # ---- decode
    t0 = time.time()
    acts_ref = O.sample_actions(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                                batch["noise"])
    print(f"oracle sample_actions {time.time() - t0:.2f}s", flush=True)
    acts = model.sample_actions("cuda", obs, noise=batch["noise"].cuda(), num_steps=10)
    torch.cuda.synchronize()
    cmp("sample_actions", acts, acts_ref)


if __name__ == "__main__":
    main()
