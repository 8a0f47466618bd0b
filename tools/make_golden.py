"""Generate tests/golden/*.pt from the CPU oracle (the reference itself cannot run here: SURVEY.md §8c).

    python tools/make_golden.py

Each fixture holds the seeds/config name, the inputs that are not re-derivable, and the oracle outputs: loss tensor,
v_t, the action chunk of sample_actions, and a few per-layer taps.  Weights are re-created from the seed
(oracle.init_params), not stored.  The GPU parity tests compare the engine against these files, so they run on the
GPU box without /root/reference and pin the oracle against accidental edits.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle import pi05_oracle as O  # noqa: E402


def make(name, oc, batch_size, seed):
    torch.manual_seed(0)
    params = O.init_params(oc, seed)
    batch = O.synthetic_batch(oc, batch_size, ragged=True)
    if batch_size > 1:
        batch["img_masks"][1][batch_size - 1] = False
    taps = {}
    with torch.no_grad():
        loss = O.forward_loss(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                              batch["actions"], batch["noise"], batch["time"], taps)
        acts = O.sample_actions(params, oc, batch["images"], batch["img_masks"], batch["tokens"], batch["token_mask"],
                                batch["noise"])
    keep = ["prefix_embs", "suffix_embs", "adarms_cond", "suffix_out", "v_t", "layer0_suffix",
            f"layer{oc.paligemma.depth - 1}_suffix"]
    out = {
        "config": name, "batch": batch_size, "weight_seed": seed,
        "img_masks": torch.stack(batch["img_masks"]), "tokens": batch["tokens"], "token_mask": batch["token_mask"],
        "actions": batch["actions"], "noise": batch["noise"], "time": batch["time"],
        "image_checksum": torch.stack([i.double().sum() for i in batch["images"]]),
        "loss": loss, "sample_actions": acts,
        "taps": {k: taps[k].clone() for k in keep},
    }
    path = os.path.join(ROOT, "tests", "golden", f"{name}_b{batch_size}.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path) // 1024, "KiB", float(loss.mean()), float(acts.abs().mean()))


if __name__ == "__main__":
    make("tiny", O.tiny_config(), 2, 0)
    make("mid", H.mid_config(), 2, 0)
