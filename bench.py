"""bench.py — the pi0.5 training step on N B200s (one process per GPU), in the driver's contract.

    python bench.py --gpus 1 --steps K --warmup W                 # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                    # N > 1 (NCCL, one flat all-reduce per dtype arena)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the CPU arm (oracle port of the reference)

A step = what scripts/train_pytorch.py:531-561 does per iteration for BASELINE.json configs[1]
("pi0.5 full fine-tune bf16, 3-cam 224x224, batch 32, 1xB200"): uint8 batch -> Observation.from_dict ->
model(observation, actions) -> loss.mean().backward() -> (gradient all-reduce) -> clip_grad_norm_(1.0) -> AdamW step ->
zero_grad(set_to_none).  Weights are seeded random-init of the full pi0.5 architecture (3.35 B trainable parameters),
data is synthetic.  `value` times the step with the uint8 batch already in HBM; `e2e` times the same
step fed from pinned host memory, with the H2D copies and the loss read-back inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# forward FLOPs per sample (SURVEY.md §8d, 2 flops/MAC, block-sparse attention) and the training multiplier
FWD_TFLOP_PER_SAMPLE = 4.674
TRAIN_TFLOP_PER_SAMPLE = 3 * FWD_TFLOP_PER_SAMPLE
METRIC = "train_samples_per_sec"
UNIT = "samples/s"


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the four dominant GEMM classes at the bench's
# shapes, from the committed `ncu --set full` captures (profiles/r01_ncu_gemm_{geglu,dgrad,wgrad,down}.md);
# key = (M, N, K, epilogue, majors) as printed by pi05_gemm_profile_report
NCU_TRAFFIC_BYTES = {
    (30976, 32768, 2048, 5, 0): 5332.3e6,   # GeGLU forward (fused gate|up weight)
    (30976, 2048, 32768, 0, 1): 5550.0e6,   # dgrad of gate|up
    (32768, 2048, 30976, 0, 3): 5455.0e6,   # wgrad of gate|up
    (30976, 2048, 16384, 4, 0): 2944.2e6,   # down projection + residual
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# synthetic data in pinned host memory
# ------------------------------------------------------------------------------------------------------------
KEYS = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")


def make_host_batch(B, rank, image_size=224, L=200, vocab=257152, horizon=50, adim=32, pin=True):
    g = torch.Generator().manual_seed(1234 + rank)
    d = {"image": {}, "image_mask": {}}
    for k in KEYS:
        d["image"][k] = torch.randint(0, 256, (B, image_size, image_size, 3), generator=g, dtype=torch.uint8)
        d["image_mask"][k] = torch.ones(B, dtype=torch.bool)
    state = torch.zeros(B, 32)
    state[:, :14] = torch.rand(B, 14, generator=g) * 2 - 1
    d["state"] = state
    toks = torch.zeros(B, L, dtype=torch.int64)
    nv = min(96, L)
    toks[:, :nv] = torch.randint(0, vocab, (B, nv), generator=g)
    mask = torch.zeros(B, L, dtype=torch.bool)
    mask[:, :nv] = True
    d["tokenized_prompt"] = toks
    d["tokenized_prompt_mask"] = mask
    actions = torch.randn(B, horizon, adim, generator=g)
    actions[..., 14:] = 0
    if pin and torch.cuda.is_available():
        for k in KEYS:
            d["image"][k] = d["image"][k].pin_memory()
            d["image_mask"][k] = d["image_mask"][k].pin_memory()
        d["state"] = d["state"].pin_memory()
        d["tokenized_prompt"] = toks.pin_memory()
        d["tokenized_prompt_mask"] = mask.pin_memory()
        actions = actions.pin_memory()
    return d, actions


def h2d_bytes(d, actions):
    n = actions.numel() * actions.element_size()
    for k in KEYS:
        n += d["image"][k].numel() + d["image_mask"][k].numel()
    for k in ("state", "tokenized_prompt", "tokenized_prompt_mask"):
        n += d[k].numel() * d[k].element_size()
    return n


def to_device(d, actions, dev):
    out = {"image": {}, "image_mask": {}}
    for k in KEYS:
        out["image"][k] = d["image"][k].to(dev, non_blocking=True)
        out["image_mask"][k] = d["image_mask"][k].to(dev, non_blocking=True)
    for k in ("state", "tokenized_prompt", "tokenized_prompt_mask"):
        out[k] = d[k].to(dev, non_blocking=True)
    return out, actions.to(dev, non_blocking=True)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, timed on the host cores (cpu_baseline / --impl reference)
# ------------------------------------------------------------------------------------------------------------
# forward TFLOP per sample of the pieces of the path, used to scale the bounded CPU sample
_TF_VIT_LAYER = 0.6605 / 27          # all three cameras, one SigLIP layer
_TF_JOINT_LAYER = (3.8368 + 0.0311 + 0.1457) / 18   # one joint PaliGemma + expert layer incl. attention
_TF_REST = 0.0003 + 3 * 0.0012       # adaRMS / heads / projector (not depth-scaled)


_LAST_CPU_DECODE = None  # filled by cpu_reference: CPU timing of the decode metric (reported beside `decode`)


def _pick_cpu_threads() -> int:
    """All the host threads the port can USE: torch's bf16 CPU GEMM gets slower when oversubscribed on many-thread hosts
    (128 threads: 3.5x slower than 8 on the same pass), so a ~1 s probe picks the fastest of a few thread counts."""
    n = os.cpu_count() or 1
    cands = sorted({c for c in (n, n // 2, n // 4, 32, 16, 8) if 1 <= c <= n}, reverse=True)
    a = (torch.randn(968, 2048) * 0.1).to(torch.bfloat16)
    w = (torch.randn(4096, 2048) * 0.02).to(torch.bfloat16)
    best, best_t = n, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.linear(a, w)
        t0 = time.time()
        for _ in range(3):
            torch.nn.functional.linear(a, w)
        dt = time.time() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def cpu_reference(steps: int, warmup: int, budget_s: float, full: bool = True):
    """Times the oracle port of the reference on all host threads on a BOUNDED sample of the workload: forward +
    backward (torch.autograd) of ONE sample through the full-width architecture truncated to 1 of 27 SigLIP layers and
    1 of 18 joint Gemma layers (same tensor shapes, same kernels per layer), scaled to a whole sample by the FLOP
    ratio of SURVEY.md §8d.  A whole sample is ~14 TFLOP and takes minutes per pass on host cores.
    Returns (samples_per_sec, cores, sample_description, steps_done)."""
    import dataclasses

    from oracle import pi05_oracle as O

    cores = _pick_cpu_threads()
    torch.set_num_threads(cores)
    if full:
        base_cfg = O.OracleConfig()
        dv, dg = 1, 1
        oc = dataclasses.replace(base_cfg, vit_depth=dv,
                                 paligemma=dataclasses.replace(base_cfg.paligemma, depth=dg),
                                 expert=dataclasses.replace(base_cfg.expert, depth=dg))
        frac = (dv * _TF_VIT_LAYER + dg * _TF_JOINT_LAYER + _TF_REST) / FWD_TFLOP_PER_SAMPLE
    else:
        oc = O.tiny_config()
        frac = 1.0
    params = {}
    # values do not matter for timing: tile a 4M-element N(0, 0.02) pattern (bf16 normal_ on CPU is very slow)
    base = {torch.bfloat16: (torch.randn(1 << 22) * 0.02).to(torch.bfloat16), torch.float32: torch.randn(1 << 22) * 0.02}
    for name, (shape, dt) in O.param_specs(oc).items():
        n = 1
        for s in shape:
            n *= s
        reps = (n + (1 << 22) - 1) >> 22
        params[name] = base[dt].repeat(reps)[:n].view(shape).clone().requires_grad_(True)
    b = O.synthetic_batch(oc, 1)

    def one():
        for p in params.values():
            p.grad = None
        loss = O.forward_loss(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"],
                              b["noise"], b["time"])
        loss.mean().backward()

    t0 = time.time()
    one()  # first pass doubles as warm-up and cost probe
    t_first = time.time() - t0
    done_warm = 1
    while done_warm < warmup and (time.time() - t0) + t_first * (1 + steps) < budget_s:
        one()
        done_warm += 1
    k = max(1, min(steps, int((budget_s - (time.time() - t0)) / max(t_first, 1e-3))))
    t1 = time.time()
    for _ in range(k):
        one()
    dt = (time.time() - t1) / k
    # secondary metric beside it: the 10-step action decode on the same truncated architecture, scaled by decode FLOPs
    # (prefix pass 4.635 TFLOP = SigLIP 0.6605 + PaliGemma 3.84 + rest, 10 steps x 0.0389 expert)
    global _LAST_CPU_DECODE
    _LAST_CPU_DECODE = None
    if full and (time.time() - t0) < budget_s:
        dfrac = (oc.vit_depth * 0.6605 / 27 + oc.paligemma.depth * (3.84 + 10 * 0.0389) / 18 + 0.13) / 5.024
        with torch.no_grad():
            td = time.time()
            O.sample_actions({k: v.detach() for k, v in params.items()}, oc, b["images"], b["img_masks"], b["tokens"],
                             b["token_mask"], b["noise"])
            td = time.time() - td
        _LAST_CPU_DECODE = {"decode_ms_scaled": 1e3 * td / dfrac, "measured_s": td, "flop_fraction": dfrac,
                            "sample": "sample_actions (10 steps, B=1) of the same truncated oracle, one pass, scaled by "
                                      "the decode FLOP ratio"}
    desc = (f"B=1 forward+backward of the oracle port, full widths, truncated to {oc.vit_depth}/27 SigLIP and "
            f"{oc.paligemma.depth}/18 joint Gemma layers ({100 * frac:.1f} % of a sample's FLOPs: {dt:.1f} s per pass), "
            f"scaled to a whole sample by that ratio; {k} timed pass(es) after {done_warm} warm-up, {cores} threads, "
            "bf16 weights as the reference") if full else f"tiny debug architecture, {k} passes"
    return frac / dt, cores, desc, k


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (configs[1]/[2]: 32)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds the CPU legs may take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small", action="store_true", help="debug: tiny architecture (not a valid bench number)")
    ap.add_argument("--per-param-optimizer", action="store_true",
                    help="AdamW/clip over model.parameters() exactly as the unchanged script (slower: ~1400 launches)")
    ap.add_argument("--torch-optimizer", action="store_true",
                    help="torch.optim.AdamW(fused) + clip_grad_norm_ over the 2 flat arenas instead of the engine's "
                         "fused clip+AdamW kernel (kai0_b200.optim.FusedClipAdamW)")
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, torchrun children) are routed to stderr; the JSON goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = "pi0.5 full fine-tune bf16, 3-cam 224x224, batch 32 per GPU (BASELINE.json configs[1]; configs[2] at N=8)"

    # ---------------------------------------------------------------- CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sps, cores, desc, k = cpu_reference(args.steps, min(args.warmup, 1), max(60.0, args.cpu_budget * 1.5),
                                            full=not args.small)
        line = {
            "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
            "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 / sps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "note": "reference's PyTorch path cannot be imported (needs patched "
                       "transformers 4.53.2 + jax); this is its CPU oracle port (oracle/pi05_oracle.py)"},
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        if _LAST_CPU_DECODE is not None:
            line["cpu_baseline"]["decode"] = _LAST_CPU_DECODE
        emit(line)
        return 0

    # ---------------------------------------------------------------- B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the B200 arm needs a CUDA device (there is no CPU fallback)")
    import torch.distributed as dist

    from kai0_b200 import _lib
    from kai0_b200.model import Observation
    from kai0_b200.pi0_pytorch import GemmaVariant, Pi05EngineConfig, PI0Pytorch

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    B = args.batch
    if args.small:
        cfg = Pi05EngineConfig(paligemma_variant=GemmaVariant(512, 2, 1024, 8, 1, 256),
                               action_expert_variant=GemmaVariant(256, 2, 512, 8, 1, 256), vit_width=288, vit_depth=2,
                               vit_mlp_dim=560, vit_heads=4, vocab_size=4096)
    else:
        cfg = Pi05EngineConfig()
    torch.manual_seed(1234)  # same weights on every rank (DDP would broadcast rank 0's, train_pytorch.py:441)
    model = PI0Pytorch(cfg, max_batch=B, init_weights=False).to(dev)
    model.reset_parameters()
    model.check_inputs = False
    model.direct_grads = True  # public knob: .grad = views of the flat gradient arena (no per-parameter autograd copies)
    model.train()
    if world > 1:
        model.enable_flat_allreduce()
    # optimiser over the two flat arenas (public opt-in): element-wise identical to the per-parameter
    # AdamW / global-norm clip of train_pytorch.py:469-475,557, in 2 tensors instead of ~700
    params = model.flat_parameters() if not args.per_param_optimizer else [p for p in model.parameters() if p.requires_grad]
    use_fused = not (args.per_param_optimizer or args.torch_optimizer)
    if use_fused:
        from kai0_b200.optim import FusedClipAdamW

        # global-norm clip + AdamW over the flat arenas in one engine pass (same update rule)
        optim = FusedClipAdamW(model, lr=2.5e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10, max_norm=1.0)
    else:
        optim = torch.optim.AdamW(params, lr=2.5e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10, fused=True)

    host_d, host_a = make_host_batch(B, rank, image_size=cfg.image_size, L=cfg.max_token_len, vocab=cfg.vocab_size,
                                     horizon=cfg.action_horizon, adim=cfg.action_dim)
    dev_d, dev_a = to_device(host_d, host_a, dev)
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.pi05_launch_count.restype = C.c_ulonglong

    def step(from_host: bool):
        if from_host:
            d, a = to_device(host_d, host_a, dev)
        else:
            d = {"image": dict(dev_d["image"]), "image_mask": dev_d["image_mask"], "state": dev_d["state"],
                 "tokenized_prompt": dev_d["tokenized_prompt"],
                 "tokenized_prompt_mask": dev_d["tokenized_prompt_mask"]}
            a = dev_a
        obs = Observation.from_dict(d)  # uint8 NHWC -> fp32 NCHW in [-1,1] (models/model.py:129-133)
        losses = model(obs, a)
        loss = losses.mean()
        loss.backward()
        if use_fused:
            optim.step()  # clip_grad_norm_(1.0) + AdamW in one pass
        else:
            torch.nn.utils.clip_grad_norm_(params, max_norm=1.0)
            optim.step()
        optim.zero_grad(set_to_none=True)
        if from_host:
            return float(loss.item())  # device -> host read of the step's result
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        torch.cuda.nvtx.range_push("pi05_timed_host" if from_host else "pi05_timed_dev")
        prof = (not from_host) and os.environ.get("PI05_CUDA_PROFILER")  # ncu --profile-from-start off
        if prof:
            torch.cuda.profiler.start()
        for _ in range(k):
            last = step(from_host)
        if prof:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        torch.cuda.nvtx.range_pop()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    for _ in range(max(args.warmup, 3)):
        step(False)
    step(True)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.pi05_gemm_profile_enable(1)
    n0 = lib.pi05_launch_count()
    ms_dev, _ = timed(args.steps, False)
    n1 = lib.pi05_launch_count()
    lib.pi05_gemm_profile_enable(0)
    buf = C.create_string_buffer(1 << 16)
    lib.pi05_gemm_profile_report(buf, len(buf))
    ms_e2e, last_loss = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else None

    if world > 1 and os.environ.get("PI05_BENCH_VERBOSE"):
        # cost of the exchange step alone (the two flat all-reduces + the 1/world scaling), all ranks in lock-step
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for i in range(3):
            ev[i].record()
            model._allreduce_flat_grads()
        ev[3].record()
        barrier()
        if rank == 0:
            nbytes = sum(g.numel() * g.element_size() for g in model._flat_grad.values() if g is not None)
            nbytes -= 2 * (model._flat_grad[torch.bfloat16].numel() - model._offsets[
                "paligemma_with_expert.gemma_expert.lm_head.weight"][1])  # the unused lm_head is not exchanged
            ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
            print(f"  [allreduce] {nbytes / 1e9:.2f} GB of gradients: {ts} ms -> "
                  f"{nbytes / 1e9 / (min(ts) / 1e3):.0f} GB/s algorithmic", file=sys.stderr)

    if rank == 0 and os.environ.get("PI05_TORCH_PROFILE"):
        # low-overhead per-kernel breakdown of ONE step (CUPTI via torch.profiler): where the non-GEMM time goes
        from torch.profiler import ProfilerActivity, profile

        lib.pi05_debug_set_pdl(0)  # a PDL-staged kernel's duration would include the wait for its predecessor
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(False)
            torch.cuda.synchronize()
        lib.pi05_debug_set_pdl(1)
        agg = {}
        for ev in prof.events():
            if ev.device_type is not None and str(ev.device_type).endswith("CUDA"):
                name = ev.name[:90]
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot = sum(v[1] for v in agg.values())
        print(f"  [torch.profiler] one step: {tot / 1e3:.2f} ms of kernel time", file=sys.stderr)
        for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            print(f"  {us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  x{cnt:5d}  {name}", file=sys.stderr)

    # secondary metric of BASELINE.json (configs[3]): single-frame 10-step action decode, p50 latency through the
    # public API with host inputs (H2D of one uint8 observation, D2H of the [1,50,32] action chunk), rank 0 only
    decode = None
    if rank == 0:
        model.eval()
        one_d, one_a = make_host_batch(1, 0, image_size=cfg.image_size, L=cfg.max_token_len, vocab=cfg.vocab_size,
                                       horizon=cfg.action_horizon, adim=cfg.action_dim)
        lat = []
        for i in range(5 + 20):
            t0 = time.perf_counter()
            d1, _ = to_device(one_d, one_a, dev)
            acts = model.sample_actions(dev, Observation.from_dict(d1), num_steps=10)
            acts_host = acts.cpu()  # synchronises
            if i >= 5:
                lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        decode = {"metric": "action_chunk_decode_p50_ms", "p50_ms": lat[len(lat) // 2], "min_ms": lat[0],
                  "max_ms": lat[-1], "iters": len(lat), "batch": 1, "num_steps": 10,
                  "finite": bool(torch.isfinite(acts_host).all())}
        model.train()

    if rank == 0:
        peaks, peaks_kind = load_peaks()
        samples = world * B * args.steps
        value = samples / (ms_dev / 1e3)
        e2e = samples / (ms_e2e / 1e3)
        # dominant kernel: the tcgen05 GEMM class with the largest share of event time in the timed region
        classes = []
        for ln in buf.value.decode().strip().splitlines():
            M, N, K, bt, epi, maj, cnt, ms = ln.split()
            fl = 2.0 * int(M) * int(N) * int(K) * int(bt) * int(cnt)
            classes.append({"M": int(M), "N": int(N), "K": int(K), "batch": int(bt), "epi": int(epi),
                            "majors": int(maj), "launches": int(cnt), "ms": float(ms), "flop": fl})
        if os.environ.get("PI05_BENCH_VERBOSE"):
            for c in sorted(classes, key=lambda c: -c["ms"])[:40]:
                print(f"  gemm M={c['M']:6d} N={c['N']:6d} K={c['K']:6d} b={c['batch']:5d} epi={c['epi']} maj={c['majors']} "
                      f"x{c['launches']:4d}  {c['ms'] / args.steps:8.3f} ms/step  "
                      f"{c['flop'] / (c['ms'] / 1e3) / 1e12:7.1f} TFLOP/s", file=sys.stderr)
            print(f"  step {ms_dev / args.steps:.2f} ms, tcgen05 GEMMs {sum(c['ms'] for c in classes) / args.steps:.2f} ms",
                  file=sys.stderr)
        gemm_ms = sum(c["ms"] for c in classes)
        gemm_flop = sum(c["flop"] for c in classes)
        top = max(classes, key=lambda c: c["ms"]) if classes else None
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1470.0)))
        roofline = None
        if top:
            ach = top["flop"] / (top["ms"] / 1e3) / 1e12
            roofline = {
                "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": NCU_TRAFFIC_BYTES.get((top["M"], top["N"], top["K"], top["epi"], top["majors"])),
                "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_ncu_gemm_*.md)",
                "algorithmic_bytes": 2.0 * (top["M"] * top["K"] + top["N"] * top["K"] + top["M"] * top["N"]) * top["batch"],
                "kernel": f"gemm_kernel<256,{top['epi']}> M={top['M']} N={top['N']} K={top['K']} "
                          f"majors={top['majors']} ({top['launches']} launches in the timed region)",
                "peak_source": f"{peaks_kind} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                "all_tcgen05_gemms": {"tflops": gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms else None,
                                      "share_of_step": gemm_ms / ms_dev},
                "step_model_flops_utilisation": (value / world) * TRAIN_TFLOP_PER_SAMPLE / peak_tf,
            }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload if not args.small else "DEBUG small architecture (invalid as a bench number)",
                       "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "optimizer": ("kai0_b200.optim.FusedClipAdamW: clip_grad_norm_(1.0) + AdamW(0.9,0.95,1e-8,wd 1e-10) "
                                     "over the 2 flat arenas in one engine pass" if use_fused else
                                     "torch.optim.AdamW(fused) + clip_grad_norm_(1.0) as scripts/train_pytorch.py, over "
                                     + ("model.parameters()" if args.per_param_optimizer
                                        else "model.flat_parameters() (2 flat arenas)")),
                       "l2": "per-step activations (>100 GB) and weights (7 GB) far exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(host_d, host_a),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss},
            "gpu_launches": int(n1 - n0),
            "decode": decode,
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            del optim
            torch.cuda.empty_cache()
            sps, cores, desc, _ = cpu_reference(1, 0, args.cpu_budget, full=not args.small)
            line["cpu_baseline"] = {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
            if _LAST_CPU_DECODE is not None:
                line["cpu_baseline"]["decode"] = _LAST_CPU_DECODE
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
