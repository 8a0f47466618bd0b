"""Golden outputs of the REFERENCE'S OWN real-time-chunking sampler (`Pi0RTC.sample_actions`,
src/openpi/models/pi0_rtc.py:234-360, executed in place through tools/reference_rtc_loader.py over the PyTorch-path
network of oracle/pi05_oracle.py) on seeded inputs.  Build container only; writes tests/golden/rtc_reference.pt (outputs
only: weights, observations, noise and previous chunks are regenerated from seeds by `cases()`).

    python tools/make_golden_rtc.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "rtc_reference.pt")
WEIGHT_SEED = 2


def setup(precision: str = "float32"):
    """(config, parameters, batch of 2 ragged observations with one masked camera, unguided chunk, previous chunk)."""
    from oracle import pi05_oracle as O

    oc = O.tiny_config()
    p = O.init_params(oc, seed=WEIGHT_SEED)
    if precision == "float32":
        p = {k: v.to(torch.float32) for k, v in p.items()}
    b = O.synthetic_batch(oc, 2, ragged=True)
    b["img_masks"][1][0] = False
    with torch.no_grad():
        plain = O.sample_actions(p, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"])
    prev = plain.to(torch.float32) + 0.5 * torch.randn(plain.shape, generator=torch.Generator().manual_seed(0))
    return oc, p, b, plain, prev


def cases(prev: torch.Tensor):
    """name -> keyword arguments of sample_actions (pi0_rtc.py:236-249)."""
    wide = torch.cat([prev, torch.ones(*prev.shape[:-1], 4)], dim=-1)  # more dims than the model has: cut (:319-321)
    dirty = prev[..., :14].clone()
    dirty[0, 1, 2], dirty[1, 0, 0], dirty[1, 3, 5] = float("nan"), float("inf"), float("-inf")  # zeroed (:317)
    return {
        "linear_d2_h6": dict(prev_action_chunk=prev[..., :14], inference_delay=2, execute_horizon=6,
                             prefix_attention_schedule="linear"),
        "exp_d3_h8_strong": dict(prev_action_chunk=prev[..., :14], inference_delay=3, execute_horizon=8,
                                 prefix_attention_schedule="exp", max_guidance_weight=5.0),
        "masked_delay": dict(prev_action_chunk=prev[..., :14], inference_delay=2, execute_horizon=6, mask_prefix_delay=True),
        "ones_defaults": dict(prev_action_chunk=prev[..., :14], inference_delay=None, execute_horizon=None,
                              prefix_attention_schedule="ones"),
        "zeros_d4_h5": dict(prev_action_chunk=prev[..., :14], inference_delay=4, execute_horizon=5,
                            prefix_attention_schedule="zeros"),
        "clipped_arguments": dict(prev_action_chunk=prev[..., :14], inference_delay=99, execute_horizon=99),
        "seven_dims": dict(prev_action_chunk=prev[..., :7], inference_delay=1, execute_horizon=9, mask_prefix_delay=True),
        "wider_than_model": dict(prev_action_chunk=wide, inference_delay=1, execute_horizon=7),
        "nan_inf_in_chunk": dict(prev_action_chunk=dirty, inference_delay=1, execute_horizon=7),
        "five_steps": dict(prev_action_chunk=prev[..., :14], inference_delay=2, execute_horizon=6, num_steps=5),
        "rtc_disabled": dict(prev_action_chunk=prev[..., :14], inference_delay=2, execute_horizon=6, enable_rtc=False),
        "no_previous_chunk": dict(),
    }


def main():
    import reference_rtc_loader as RL

    oc, p, b, plain, prev = setup("float32")
    out = {"weight_seed": WEIGHT_SEED, "precision": "float32", "plain": plain}
    for name, kw in cases(prev).items():
        out[name] = RL.sample_actions(p, oc, b, b["noise"], oracle_suffix_embedding=True, **kw)
        out[name + "/jax_suffix_embedding"] = RL.sample_actions(p, oc, b, b["noise"], **kw)
        moved = float((out[name] - plain).norm() / plain.norm())
        print(f"{name:22s} moved the chunk by {moved:.3f} (relative), "
              f"JAX-side suffix embedding differs by {float((out[name] - out[name + '/jax_suffix_embedding']).abs().max()):.1e}")
    # the schedules on their own (pi0_rtc.py:47-61)
    mod = RL.load()
    out["prefix_weights"] = {f"{s}/{a}/{e}/{t}": mod.get_prefix_weights(a, e, t, s).as_subclass(torch.Tensor).clone()
                             for s in ("ones", "zeros", "linear", "exp") for (a, e, t) in ((2, 6, 8), (0, 10, 10), (9, 3, 5),
                                                                                          (3, 50, 50), (0, 1, 4))}
    torch.save(out, OUT)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e3:.1f} kB)")


if __name__ == "__main__":
    main()
