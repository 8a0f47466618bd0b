"""Runs the REFERENCE'S OWN training step (test / measurement infrastructure; never imported by kai0_b200/).

`PI0Pytorch(config).to(device)`, `gradient_checkpointing_enable()`, `torch.optim.AdamW(model.parameters(), ...)` and,
per step, `model(observation, actions).mean().backward()`, `clip_grad_norm_(1.0)`, `optim.step()`,
`optim.zero_grad(set_to_none=True)` -- scripts/train_pytorch.py:417-421,469-475,540-561 -- on the reference's own module
executed in place through tools/reference_loader.py (from /root/reference when present, else from the offline install
under baseline/_ref).  Used by bench.py for `--impl reference`, `cpu_baseline` (host cores) and the `reference_gpu`
anchor (the same eager module on the B200: what a kai0 user runs today).
"""
from __future__ import annotations

import os
import sys
import time
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import reference_loader as RL  # noqa: E402

KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
# forward TFLOP per sample by depth-scaled piece: SigLIP tower (3 cameras), joint Gemma stack, the rest
TF_VIT, TF_JOINT, TF_REST, TF_FWD = 0.6605, 3.8368 + 0.0311 + 0.1457, 0.0003 + 3 * 0.0012, 4.674


def available() -> bool:
    return RL.available()


def physical_threads() -> int:
    """Host threads for the CPU arm: one per physical core when that can be read (SMT siblings only fight over the AMX /
    AVX-512 units in a GEMM-bound pass), else os.cpu_count(); never more than the process's affinity mask."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        cores = set()
        phys = core = None
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("physical id"):
                phys = ln.split(":")[1].strip()
            elif ln.startswith("core id"):
                core = ln.split(":")[1].strip()
            elif not ln.strip():
                if phys is not None and core is not None:
                    cores.add((phys, core))
                phys = core = None
        if cores:
            n = min(n, len(cores))
    except OSError:
        pass
    return max(1, n)


class Obs:
    """What `Observation.from_dict` hands the model (models/model.py:122-157), duck-typed."""

    def __init__(self, d):
        self.images = d["image"]
        self.image_masks = d["image_mask"]
        self.state = d["state"]
        self.tokenized_prompt = d["tokenized_prompt"]
        self.tokenized_prompt_mask = d["tokenized_prompt_mask"]
        self.token_ar_mask = None
        self.token_loss_mask = None


def from_dict(d):
    """uint8 NHWC -> fp32 NCHW in [-1, 1] exactly as Observation.from_dict (models/model.py:129-133)."""
    out = dict(d)
    out["image"] = {k: (v.to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0) if v.dtype == torch.uint8 else v
                    for k, v in d["image"].items()}
    return Obs(out)


def build(device, *, vit_layers=None, depth=None, seed=0):
    """The reference PI0Pytorch at BASELINE.json's architecture (bf16 dtype map), random weights (values do not matter
    for timing; N(0, 0.02)-like pattern so activations stay finite), in train mode with its own gradient checkpointing
    on.  `vit_layers` / `depth` truncate the stacks for the BOUNDED-sample fallback only."""
    from transformers.initialization import no_init_weights

    if depth is not None:
        RL.register_variant("gemma_2b_trunc", RL.SizeRecord(2048, depth, 16384, 8, 1, 256))
        RL.register_variant("gemma_300m_trunc", RL.SizeRecord(1024, depth, 4096, 8, 1, 256))
    if RL._loaded is not None and vit_layers is not None:
        raise RuntimeError("reference already loaded with another vision depth in this process")
    p0 = RL.load(vision_layers=vit_layers)
    cfg = types.SimpleNamespace(pi05=True, paligemma_variant="gemma_2b" if depth is None else "gemma_2b_trunc",
                                action_expert_variant="gemma_300m" if depth is None else "gemma_300m_trunc",
                                dtype="bfloat16", action_horizon=50, action_dim=32, max_token_len=200)
    with no_init_weights():
        m = p0.PI0Pytorch(cfg)
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1 << 22, generator=g) * 0.02
    with torch.no_grad():
        for name, p in m.named_parameters():
            n = p.numel()
            reps = (n + base.numel() - 1) // base.numel()
            src = base.repeat(reps)[:n].view(p.shape)
            if name.endswith("layer_norm1.weight") or name.endswith("layer_norm2.weight") or name.endswith(
                    "post_layernorm.weight"):
                src = src + 1.0
            p.copy_(src.to(p.dtype))
    m = m.to(device)
    m.train()
    m.gradient_checkpointing_enable()  # train_pytorch.py:419-421
    return m


class Stepper:
    """One optimiser step of scripts/train_pytorch.py:540-561 on the reference model."""

    def __init__(self, model, lr=2.5e-5):
        self.model = model
        self.optim = torch.optim.AdamW(model.parameters(), lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)

    def step(self, host_d, actions, device):
        d = {"image": {k: v.to(device) for k, v in host_d["image"].items()},
             "image_mask": {k: v.to(device) for k, v in host_d["image_mask"].items()},
             "state": host_d["state"].to(device), "tokenized_prompt": host_d["tokenized_prompt"].to(device),
             "tokenized_prompt_mask": host_d["tokenized_prompt_mask"].to(device)}
        a = actions.to(torch.float32).to(device)
        losses = self.model(from_dict(d), a)
        loss = losses.mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=1.0)
        self.optim.step()
        self.optim.zero_grad(set_to_none=True)
        return float(loss.item())


def slice_batch(d, actions, n):
    out = {"image": {k: v[:n] for k, v in d["image"].items()}, "image_mask": {k: v[:n] for k, v in d["image_mask"].items()},
           "state": d["state"][:n], "tokenized_prompt": d["tokenized_prompt"][:n],
           "tokenized_prompt_mask": d["tokenized_prompt_mask"][:n]}
    return out, actions[:n]


def time_cpu(host_d, actions, steps: int, warmup: int, budget_s: float, threads: int | None = None):
    """Times `steps` reference training steps on the host cores after `warmup` untimed ones, each step ONE whole sample
    (B = 1 of the B = 32 workload: the CPU pass is sequential in the batch, so samples/s does not depend on B).  If the
    first (warm-up) step shows that steps + warmup whole samples cannot end within `budget_s`, the remaining steps run a
    BOUNDED sample instead -- the same model with its stacks truncated to `depth` of 18 joint layers and a proportional
    part of the 27 SigLIP layers -- and the result is scaled by the stated FLOP ratio; both figures are returned.
    Returns a dict (see bench.py)."""
    threads = threads or physical_threads()
    torch.set_num_threads(threads)
    t_build = time.time()
    m = build("cpu")
    st = Stepper(m)
    d1, a1 = slice_batch(host_d, actions, 1)
    t_build = time.time() - t_build
    t0 = time.time()
    st.step(d1, a1, "cpu")
    t_first = time.time() - t0
    remaining = steps + max(warmup - 1, 0)
    info = {"threads": threads, "build_s": t_build, "first_whole_sample_step_s": t_first}
    if t_first * remaining <= max(budget_s - t_build - t_first, 0.0) or remaining == 0:
        for _ in range(max(warmup - 1, 0)):
            st.step(d1, a1, "cpu")
        t1 = time.time()
        for _ in range(steps):
            st.step(d1, a1, "cpu")
        dt = (time.time() - t1) / max(steps, 1) if steps else t_first
        info.update({"mode": "whole", "s_per_step": dt, "samples_per_s": 1.0 / dt, "scale": 1.0,
                     "sample": f"B=1 WHOLE sample per step (all 27 SigLIP + 18+18 Gemma layers, own gradient "
                               f"checkpointing, clip + AdamW), {steps} timed steps after {max(warmup, 1)} warm-up, "
                               f"{threads} threads"})
        return info
    # bounded fallback: truncated stacks on a second instance of the reference model
    del st, m
    left = max(budget_s - t_build - t_first, 30.0)
    frac_target = min(1.0, left / (t_first * remaining * 1.3))
    depth = max(1, min(17, int(18 * frac_target)))
    vit = max(1, min(26, int(27 * depth / 18)))
    b = time_cpu_bounded(host_d, actions, steps, max(warmup - 1, 0), depth, vit, threads)
    b.update({"first_whole_sample_step_s": t_first, "measured_whole_sample_samples_per_s": 1.0 / t_first,
              "build_s": t_build + b["build_s"],
              "sample": b["sample"] + f"; ONE whole untruncated sample (all layers, same step) was run first as a warm-up: "
                                      f"{t_first:.1f} s = {1.0 / t_first:.4f} samples/s incl. first-call overheads"})
    return b


def time_cpu_bounded(host_d, actions, steps: int, warmup: int, depth: int, vit: int, threads: int | None = None):
    """`steps` timed reference training steps (B = 1) after `warmup` untimed ones on the reference model truncated to
    `depth` of 18 joint Gemma layers and `vit` of 27 SigLIP layers (same widths, same per-layer code); samples/s is the
    measured rate times the stated forward-FLOP fraction of a whole sample.  The embedding table and its
    AdamW update are not depth-scaled, so the scaled figure slightly under-states the CPU."""
    threads = threads or physical_threads()
    torch.set_num_threads(threads)
    t_build = time.time()
    m = build("cpu", depth=depth)
    enc = m.paligemma_with_expert.paligemma.model.vision_tower.vision_model.encoder
    enc.layers = torch.nn.ModuleList(list(enc.layers)[:vit])  # the tower's depth is fixed when the modules are first loaded
    st = Stepper(m)
    t_build = time.time() - t_build
    d1, a1 = slice_batch(host_d, actions, 1)
    frac = (TF_VIT * vit / 27 + TF_JOINT * depth / 18 + TF_REST) / TF_FWD
    for _ in range(warmup):
        st.step(d1, a1, "cpu")
    t1 = time.time()
    for _ in range(steps):
        st.step(d1, a1, "cpu")
    dt = (time.time() - t1) / max(steps, 1)
    return {"threads": threads, "build_s": t_build, "mode": "bounded", "s_per_step": dt, "samples_per_s": frac / dt,
            "scale": frac,
            "sample": f"B=1 per step through the reference's own PI0Pytorch truncated to {vit}/27 SigLIP and {depth}/18 joint "
                      f"Gemma layers = {100 * frac:.1f} % of a sample's forward FLOPs ({dt:.2f} s/step measured over {steps} "
                      f"steps after {warmup} warm-up), scaled to a whole sample by that ratio; {threads} threads"}


def time_gpu(host_d, actions, device, steps=3, warmup=2, batch=32):
    """The same reference training step, eager, on `device` (a B200): samples/s at the largest batch <= `batch` that
    fits.  CUDA-event timing, synchronised on both sides."""
    m = build(device)
    st = Stepper(m)
    b = batch
    while True:
        try:
            d, a = slice_batch(host_d, actions, b)
            for _ in range(warmup):
                st.step(d, a, device)
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                last = st.step(d, a, device)
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1) / steps
            return {"value": b / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "batch": b, "steps": steps,
                    "warmup": warmup, "last_loss": last, "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 1e9,
                    "what": "the reference's own PI0Pytorch (unmodified, executed in place), eager, bf16 dtype map, its own "
                            "per-layer torch.utils.checkpoint, clip_grad_norm_ + torch.optim.AdamW: "
                            "scripts/train_pytorch.py:417-421,469-475,540-561 on this B200"}
        except torch.OutOfMemoryError:
            st.optim.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            if b == 1:
                raise
            b //= 2
