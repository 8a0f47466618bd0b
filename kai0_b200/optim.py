"""Fused global-norm clip + AdamW over the model's flat arenas (opt-in).

Drop-in for the two caller lines `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)` +
`torch.optim.AdamW(...).step()` of scripts/train_pytorch.py:557-560: same update rule (torch's fused AdamW math,
moments kept in the parameter dtype), one deterministic norm reduction + one streaming pass per dtype arena instead of
~1400 kernel launches.  Requires `model.direct_grads = True` (gradients stay in the flat gradient arenas).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedClipAdamW:
    def __init__(self, model, lr=2.5e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10, max_norm=1.0):
        self.model = model
        frozen = [n for n, p in model.named_parameters() if not p.requires_grad and "lm_head" not in n]
        if frozen:
            raise ValueError("FusedClipAdamW updates every element of the flat arenas; frozen parameters are not supported "
                             f"(requires_grad=False on {frozen[:3]}...): use torch.optim.AdamW over model.parameters()")
        model.direct_grads = True
        self.param_groups = [dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, max_norm=max_norm)]
        self.flat = model.flat_parameters()  # [bf16 (trainable prefix), fp32]
        self.m = [torch.zeros_like(p.data) for p in self.flat]
        self.v = [torch.zeros_like(p.data) for p in self.flat]
        self.step_count = 0
        self._scratch = torch.zeros(4096, dtype=torch.float32, device=self.flat[0].device)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.flat:
            p.grad = None

    @torch.no_grad()
    def step(self):
        """Clips (global L2 norm over every trainable gradient) and applies AdamW. Returns the pre-clip gradient
        norm as a 0-dim device tensor (what clip_grad_norm_ returns)."""
        live = self.model._flat_params
        if live is None or live[0] is not self.flat[0] or live[1] is not self.flat[1]:
            raise RuntimeError("FusedClipAdamW: the model's arenas were re-created (.to() / .cuda() after the optimiser was "
                               "built); create the optimiser after moving the model, as train_pytorch.py:417,469 does")
        gb, gf = self.flat[0].grad, self.flat[1].grad
        if gb is None or gf is None:
            raise RuntimeError("FusedClipAdamW.step(): no gradients (run backward with model.direct_grads = True)")
        g = self.param_groups[0]
        self.step_count += 1
        dev = self.flat[0].device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        # model.grad_scale = 1 / world when the data-parallel exchange leaves SUMS in the arenas (average="optimizer")
        rc = _lib.lib().pi05_fused_clip_adamw_scaled(
            self.flat[0].data_ptr(), gb.data_ptr(), self.m[0].data_ptr(), self.v[0].data_ptr(), self.flat[0].numel(),
            self.flat[1].data_ptr(), gf.data_ptr(), self.m[1].data_ptr(), self.v[1].data_ptr(), self.flat[1].numel(),
            float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
            int(self.step_count), float(g["max_norm"] or 0.0), float(getattr(self.model, "grad_scale", 1.0)),
            self._scratch.data_ptr(), stream)
        _lib.check(rc, "pi05_fused_clip_adamw_scaled")
        return self._scratch[0]

    def state_dict(self):
        return {"step": self.step_count, "param_groups": self.param_groups,
                "exp_avg": [t.clone() for t in self.m], "exp_avg_sq": [t.clone() for t in self.v]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.param_groups = sd["param_groups"]
        for dst, src in zip(self.m, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.v, sd["exp_avg_sq"]):
            dst.copy_(src)
