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

    # ---- checkpoint interchange -----------------------------------------------------------------------------
    def _named_slots(self):
        """(index in list(model.parameters()), name, arena index, offset, numel, shape) of every parameter the update
        touches AND the reference's autograd gives a gradient: the entries a stock torch.optim.AdamW over
        model.parameters() holds state for (train_pytorch.py:469-475)."""
        from .pi0_pytorch import _UNUSED

        model = self.model
        used_bf16 = self.flat[0].numel()
        slots = []
        for i, (name, p) in enumerate(model.named_parameters()):
            dt, off, n, shape = model._offsets[name]
            if name in _UNUSED or name in model._dead_grad_names or not p.requires_grad:
                continue
            arena = 0 if dt == torch.bfloat16 else 1
            if arena == 0 and off + n > used_bf16:
                continue
            slots.append((i, name, arena, off, n, shape))
        return slots

    def state_dict(self, format: str = "flat"):
        """format="flat": this class's own compact record (two moment arenas per statistic).
        format="torch": the layout `torch.optim.AdamW(model.parameters(), ...).state_dict()` has -- per-parameter `step`,
        `exp_avg`, `exp_avg_sq` keyed by the parameter's position in `model.parameters()` -- so an `optimizer.pt` written
        here (train_pytorch.py:170) resumes under the reference's stock optimiser and vice versa."""
        if format == "flat":
            return {"step": self.step_count, "param_groups": self.param_groups,
                    "exp_avg": [t.clone() for t in self.m], "exp_avg_sq": [t.clone() for t in self.v]}
        if format != "torch":
            raise ValueError("format must be 'flat' or 'torch'")
        g = self.param_groups[0]
        n_params = sum(1 for _ in self.model.parameters())
        proto = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=g["lr"], betas=g["betas"], eps=g["eps"],
                                  weight_decay=g["weight_decay"]).state_dict()["param_groups"][0]
        group = {**proto, "params": list(range(n_params))}
        state = {}
        if self.step_count > 0:
            for i, _, arena, off, n, shape in self._named_slots():
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[arena][off:off + n].view(shape).clone(),
                            "exp_avg_sq": self.v[arena][off:off + n].view(shape).clone()}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts both layouts of `state_dict` (a stock AdamW `optimizer.pt` of the reference included)."""
        if "state" in sd and "exp_avg" not in sd:
            groups = sd["param_groups"]
            if len(groups) != 1:
                raise ValueError("FusedClipAdamW holds one parameter group (train_pytorch.py:469-475 builds one)")
            g = groups[0]
            mine = self.param_groups[0]
            self.param_groups = [dict(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], weight_decay=g["weight_decay"],
                                      max_norm=mine["max_norm"])]
            for t in (*self.m, *self.v):
                t.zero_()
            steps = set()
            slots = {i: s for s in self._named_slots() for i in (s[0],)}
            for key, rec in sd["state"].items():
                slot = slots.get(int(key))
                if slot is None:
                    raise ValueError(f"optimizer state for parameter #{key}, which this engine does not update")
                _, name, arena, off, n, shape = slot
                if tuple(rec["exp_avg"].shape) != tuple(shape):
                    raise ValueError(f"optimizer state of {name}: shape {tuple(rec['exp_avg'].shape)} != {tuple(shape)}")
                self.m[arena][off:off + n].view(shape).copy_(rec["exp_avg"])
                self.v[arena][off:off + n].view(shape).copy_(rec["exp_avg_sq"])
                steps.add(int(float(rec["step"])))
            if len(steps) > 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the fused update keeps ONE step count")
            self.step_count = steps.pop() if steps else 0
            return
        self.step_count = int(sd["step"])
        self.param_groups = sd["param_groups"]
        for dst, src in zip(self.m, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.v, sd["exp_avg_sq"]):
            dst.copy_(src)
