"""Pins oracle/pi05_oracle.py to the REFERENCE'S OWN PyTorch model.

tests/golden/reference_pin.pt holds outputs (loss tensor of `PI0Pytorch.forward`, action chunk of
`PI0Pytorch.sample_actions`) of /root/reference/src/openpi/models_pytorch/pi0_pytorch.py itself, executed in place on the
CPU of the build container (tools/reference_loader.py: patched transformers files loaded over the installed
transformers, jax stubbed; tools/make_golden_reference.py) in both of the reference's precisions.  Weights and inputs
are regenerated from seeds (tools/reference_pin.py).

* float32: the oracle must agree to fp32 accumulation-order noise (asserted 1e-5 relative; measured 1.5e-7 / 9e-8);
* bfloat16 dtype map: to the bf16 noise floor (asserted 3e-3 on the loss, 1e-3 on the action chunk = north_star's
  tolerance; measured 1.1e-3 / 3.7e-4).
"""
import os
import sys

import pytest
import torch

import helpers as H  # noqa: F401
from oracle import pi05_oracle as O

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import reference_pin as PIN  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_pin.pt")


def _oracle_outputs(precision):
    oc = PIN.oracle_config()
    params = PIN.pin_weights(O.param_specs(oc), dtype_map=precision == "bfloat16")
    b = PIN.pin_inputs()
    with torch.no_grad():
        loss = O.forward_loss(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"],
                              b["noise"], b["time"])
        acts = O.sample_actions(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["noise"])
    return loss, acts


@pytest.mark.parametrize("precision,tol_loss,tol_actions", [("float32", 1e-5, 1e-5), ("bfloat16", 3e-3, 1e-3)])
def test_oracle_matches_the_reference_model(precision, tol_loss, tol_actions):
    g = torch.load(GOLD)
    assert g["weight_seed"] == PIN.WEIGHT_SEED and tuple(g["pg"]) == PIN.PG and tuple(g["ex"]) == PIN.EX
    loss, acts = _oracle_outputs(precision)
    assert loss.shape == g[f"loss_{precision}"].shape == (PIN.BATCH, 50, 32)
    assert H.rel_err(loss, g[f"loss_{precision}"]) < tol_loss
    assert H.rel_err(acts, g[f"actions_{precision}"]) < tol_actions


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/openpi"), reason="reference checkout not present (GPU box)")
def test_committed_reference_outputs_are_reproducible_here():
    """Build container only: run the reference again (float32 precision, the cheaper one) and compare bit-exactly."""
    import make_golden_reference as MG

    torch.set_num_threads(2)
    oc = PIN.oracle_config()
    params = PIN.pin_weights(O.param_specs(oc), dtype_map=False)
    p0, m = MG.build_reference("float32")
    loss, acts = MG.run_reference(p0, m, params, PIN.pin_inputs())
    g = torch.load(GOLD)
    assert torch.equal(loss, g["loss_float32"]) and torch.equal(acts, g["actions_float32"])
