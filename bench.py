"""bench.py — the pi0.5 training step on N B200s (one process per GPU), in the driver's contract.

    python bench.py --gpus 1 --steps K --warmup W                 # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                    # N > 1 (NCCL, one flat all-reduce per dtype arena)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the CPU arm: the reference's OWN PI0Pytorch on host cores

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
# FLOPs the engine does NOT execute: the last joint layer's prefix-stream o_proj, MLP and prefix-query attention feed
# nothing on the loss path (pi0_pytorch.py:350-358 keeps suffix_out only), so forward and backward skip them
# (engine.cu joint_layer_forward `prefix_live`): 2*968*2048*2048 + 3*2*968*2048*16384 + 2*2*(968*8)*968*256 flop forward
SKIPPED_FWD_TFLOP_PER_SAMPLE = (2 * 968 * 2048 * 2048 + 3 * 2 * 968 * 2048 * 16384 + 2 * 2 * 968 * 8 * 968 * 256) / 1e12
EXECUTED_TRAIN_TFLOP_PER_SAMPLE = TRAIN_TFLOP_PER_SAMPLE - 3 * SKIPPED_FWD_TFLOP_PER_SAMPLE


def executed_train_tflop_per_sample(prefix_rows: int) -> float:
    """FLOPs the engine executes per trained sample when the prefix has `prefix_rows` rows (968 dense; 768 + the batch's
    longest prompt rounded up to 8 with prompt padding removal): SigLIP and the expert stream do not depend on it, the
    PaliGemma projections / MLP scale with it, attention with prefix^2 + suffix * (prefix + suffix); the last layer's dead
    prefix work is skipped either way (SURVEY.md §8d constants)."""
    P, A = prefix_rows, 50
    lin = 3.8368 * P / 968.0
    attn = 4 * 8 * 256 * (P * P + A * (P + A)) * 18 / 1e12
    fwd = 0.6605 + lin + 0.0311 + attn + 0.0003
    dead = (2 * P * 2048 * 2048 + 3 * 2 * P * 2048 * 16384 + 2 * 2 * P * 8 * P * 256) / 1e12
    return 3 * (fwd - dead)
METRIC = "train_samples_per_sec"
UNIT = "samples/s"


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the four dominant GEMM classes at the bench's
# shapes, from the committed `ncu --set full` captures (profiles/r02_ncu_gemm_{geglu,dgrad,wgrad,down}.md, round-2 code);
# key = (M, N, K, epilogue, majors) as printed by pi05_gemm_profile_report
NCU_TRAFFIC_BYTES = {
    (30976, 32768, 2048, 5, 0): 5301.8e6,   # GeGLU forward (fused gate|up weight)          profiles/r02_ncu_gemm_geglu.md
    (30976, 2048, 32768, 0, 1): 6036.1e6,   # dgrad of gate|up                              profiles/r02_ncu_gemm_dgrad.md
    (32768, 2048, 30976, 0, 3): 5536.2e6,   # wgrad of gate|up                              profiles/r02_ncu_gemm_wgrad.md
    (30976, 2048, 16384, 4, 0): 3068.0e6,   # down projection + residual                    profiles/r02_ncu_gemm_down.md
}


def ncu_traffic(M, N, K, epi, majors):
    """(bytes per launch, note) of the ncu capture for this GEMM class.  The round-2 captures were taken at the dense prefix
    (B * 968 = 30976 prompt + image rows); with prompt padding removal the same class runs on B * 864 = 27648 rows: the
    captured figure is reported with that stated, not rescaled."""
    hit = NCU_TRAFFIC_BYTES.get((M, N, K, epi, majors))
    if hit is not None:
        return hit, "captured at this shape"
    for (m, n, k, e, mj), v in NCU_TRAFFIC_BYTES.items():
        same_class = e == epi and mj == majors and ((n == N and k == K) or (m == M and n == N) or (m == M and k == K))
        if same_class:
            return v, (f"captured on the same GEMM class at the dense prefix (M={m}, N={n}, K={k}); this run's launch has "
                       f"M={M}, N={N}, K={K} (prompt padding removed): expect ~{min(M, m) * min(N, n) * min(K, k) / (m * n * k):.2f}x")
    return None, "no capture for this class"


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
# CPU arm: the reference's OWN PI0Pytorch training step timed on the host cores (cpu_baseline / --impl reference)
# ------------------------------------------------------------------------------------------------------------
def _reference_runner():
    """tools/reference_runner.py when the reference package is reachable (/root/reference in the build container, the
    offline install under baseline/_ref on the GPU box), else None."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import reference_runner as RR
    except Exception:  # noqa: BLE001
        return None
    return RR if RR.available() else None


def cpu_reference(host_d, host_a, steps: int, warmup: int, budget_s: float, whole_first: bool):
    """The reference's own training step (train_pytorch.py:540-561: forward with its own augmentation and per-layer
    gradient checkpointing, backward, clip_grad_norm_, torch.optim.AdamW) on the host cores, B = 1 per step.
    whole_first: the first warm-up step is ONE WHOLE untruncated sample (measured and reported); if K + W whole samples
    do not fit the budget, the timed steps run the same reference model truncated in depth and are scaled by the stated
    FLOP ratio (tools/reference_runner.time_cpu).  Returns the runner's dict, or the oracle port's if the reference is
    not reachable (kind "port")."""
    RR = _reference_runner()
    if RR is not None:
        if whole_first:
            info = RR.time_cpu(host_d, host_a, steps, warmup, budget_s)
        else:
            info = RR.time_cpu_bounded(host_d, host_a, steps, warmup, depth=1, vit=2)
        info["kind"] = "reference"
        return info
    return _cpu_port(steps, warmup, budget_s)


def _cpu_port(steps: int, warmup: int, budget_s: float):
    """Fallback when no reference package is reachable: forward + backward of the oracle port (oracle/pi05_oracle.py) at
    full widths, truncated to 1 of 27 SigLIP and 1 of 18 joint layers, scaled by the FLOP ratio."""
    import dataclasses

    from oracle import pi05_oracle as O

    threads = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    base_cfg = O.OracleConfig()
    oc = dataclasses.replace(base_cfg, vit_depth=1, paligemma=dataclasses.replace(base_cfg.paligemma, depth=1),
                             expert=dataclasses.replace(base_cfg.expert, depth=1))
    frac = (0.6605 / 27 + (3.8368 + 0.0311 + 0.1457) / 18 + 0.0039) / FWD_TFLOP_PER_SAMPLE
    base = {torch.bfloat16: (torch.randn(1 << 22) * 0.02).to(torch.bfloat16), torch.float32: torch.randn(1 << 22) * 0.02}
    params = {}
    for name, (shape, dt) in O.param_specs(oc).items():
        n = 1
        for d in shape:
            n *= d
        params[name] = base[dt].repeat((n + (1 << 22) - 1) >> 22)[:n].view(shape).clone().requires_grad_(True)
    b = O.synthetic_batch(oc, 1)

    def one():
        for p in params.values():
            p.grad = None
        O.forward_loss(params, oc, b["images"], b["img_masks"], b["tokens"], b["token_mask"], b["actions"], b["noise"],
                       b["time"]).mean().backward()

    for _ in range(max(warmup, 1)):
        one()
    t1 = time.time()
    for _ in range(steps):
        one()
    dt = (time.time() - t1) / max(steps, 1)
    return {"kind": "port", "threads": threads, "mode": "bounded", "s_per_step": dt, "samples_per_s": frac / dt,
            "scale": frac, "sample": f"oracle PORT (no reference package reachable): B=1 forward+backward truncated to 1/27 "
                                     f"SigLIP + 1/18 joint layers = {100 * frac:.1f} % of a sample's FLOPs, scaled; "
                                     f"{threads} threads"}


def serving_leg(timeout_s: float = 150.0):
    """SURVEY §8 row f4 (serving side), reported beside the decode latency: `tools/serving_probe.py` in a CHILD process (its
    own CUDA context: nothing it does can touch the measured line) -- the full architecture behind
    `kai0_b200.serving.Policy.infer_batch`, request dicts in host memory in, reply dicts out, for batches of 1 and 8."""
    import subprocess
    import tempfile

    out = os.path.join(tempfile.mkdtemp(prefix="kai0_serving_"), "probe.jsonl")
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "serving_probe.py"), "--out", out, "--batches", "1,8",
                        "--iters", "8"], timeout=timeout_s, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        recs = [json.loads(ln) for ln in open(out) if ln.strip()]
        by_batch = {r["batch"]: r for r in recs if "batch" in r}
        if not by_batch or any("error" in r for r in by_batch.values()):
            return {"unavailable": str([r.get("error") for r in by_batch.values()])[:300]}
        return {"what": "Policy.infer_batch at the full architecture: host request dicts -> transforms -> pinned staging -> "
                        "H2D -> sample_actions (10 steps, CUDA graph) -> D2H -> reply transforms (tools/serving_probe.py)",
                "batches": {str(b): {k: r[k] for k in ("p50_ms", "min_ms", "max_ms", "iters", "model_call_p50_ms",
                                                      "requests_per_s", "rel_err_vs_batch_of_one", "finite")}
                            for b, r in sorted(by_batch.items())}}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}


def bench_config(world: int, B: int, small: bool):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    return {"workload": ("pi0.5 full fine-tune bf16, 3-cam 224x224, batch 32 per GPU (BASELINE.json configs[1]; configs[2] "
                         "at N=8)") if not small else "DEBUG small architecture (invalid as a bench number)",
            "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
            "step": "uint8 batch -> Observation.from_dict -> forward (train-time augmentation on) -> backward -> gradient "
                    "all-reduce (N>1) -> clip_grad_norm_(1.0) -> AdamW(0.9, 0.95, eps 1e-8, wd 1e-10, state in the parameter "
                    "dtype) -> zero_grad  (scripts/train_pytorch.py:531-561)",
            "l2": "per-step activations (>100 GB) and weights (7 GB) far exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (configs[1]/[2]: 32)")
    ap.add_argument("--ref-budget", type=float, default=420.0,
                    help="--impl reference: seconds the whole run (build + warm-up + K steps) should stay within")
    ap.add_argument("--no-skip-padding", action="store_true",
                    help="compute the padded prompt slots too (dense, as the reference does); default: the engine drops the "
                         "slots that are padding in every sample of the batch, outputs and gradients unchanged")
    ap.add_argument("--overlap", action="store_true",
                    help="N>1: the engine's chunked exchange overlapped with backward instead of one exchange at the end of "
                         "backward (measured slower on power-capped B200s: profiles/r02_exchange_overlap.md)")
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the reference_gpu anchor (the reference's own eager module timed on this GPU, N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-serving", action="store_true",
                    help="skip the serving leg (tools/serving_probe.py in a child process: request batches of 1 and 8 "
                         "through kai0_b200.serving.Policy.infer_batch, host dicts in / host replies out)")
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

    # ---------------------------------------------------------------- CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        host_d, host_a = make_host_batch(args.batch, 0, pin=False)
        info = cpu_reference(host_d, host_a, args.steps, args.warmup, args.ref_budget, whole_first=not args.small)
        sps = info["samples_per_s"]
        cb = {"value": sps, "unit": UNIT, "cores": info["threads"], "kind": info["kind"], "sample": info["sample"],
              "mode": info["mode"], "flop_fraction_per_step": info["scale"], "measured_s_per_step": info["s_per_step"]}
        for k in ("first_whole_sample_step_s", "measured_whole_sample_samples_per_s", "build_s"):
            if k in info:
                cb[k] = info[k]
        emit({
            "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * info["s_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(args.gpus, args.batch, args.small),
            "implementation": "the reference's own PI0Pytorch (executed in place, unmodified) on the host CPU; each step = one "
                              "sample (B = 1) of the B = 32 workload, see cpu_baseline.sample; ms_per_step is the measured "
                              "time of that step, value = flop_fraction_per_step / measured_s_per_step",
            "cpu_baseline": cb,
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
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
    model.skip_prompt_padding = not args.no_skip_padding
    model.direct_grads = True  # public knob: .grad = views of the flat gradient arena (no per-parameter autograd copies)
    model.train()
    use_fused = not (args.per_param_optimizer or args.torch_optimizer)
    if world > 1:
        # engine-owned exchange (pi05_set_grad_exchange): ncclAllReduce of the two gradient arenas issued by pi05_backward
        # itself; with the fused optimiser the 1/world average is folded into the update instead of a separate pass over
        # the 7 GB of gradients
        model.enable_flat_allreduce(overlap=args.overlap, average="optimizer" if use_fused else "in_place")
    # optimiser over the two flat arenas (public opt-in): element-wise identical to the per-parameter
    # AdamW / global-norm clip of train_pytorch.py:469-475,557, in 2 tensors instead of ~700
    params = model.flat_parameters() if not args.per_param_optimizer else [p for p in model.parameters() if p.requires_grad]
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
        # uint8 NHWC images go to the engine as they are: from_dict's `x / 255 * 2 - 1` (models/model.py:129-133) is taken
        # inside the preprocessing kernel, which writes the patch-embedding GEMM operand directly
        obs = Observation.from_dict(d, keep_uint8=True)
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
    # each timed region gets its own clock samples: on a power-capped part the SM clock drifts between back-to-back
    # regions, so `value` and `e2e` are only comparable together with the clocks they ran at
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.pi05_gemm_profile_enable(1)
    n0 = lib.pi05_launch_count()
    ms_dev, _ = timed(args.steps, False)
    n1 = lib.pi05_launch_count()
    lib.pi05_gemm_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    buf = C.create_string_buffer(1 << 16)
    lib.pi05_gemm_profile_report(buf, len(buf))
    sampler2 = ClockSampler(local_rank)
    if rank == 0:
        sampler2.start()
    ms_e2e, last_loss = timed(args.steps, True)
    clocks_e2e = sampler2.stop() if rank == 0 else None

    exchange = None
    if world > 1:
        # the exchange step ALONE on an otherwise idle GPU (one-shot C-ABI entry, same communicator): its cost if it were
        # not overlapped, and the NCCL bus bandwidth it reaches
        barrier()
        h = model._train_engine_handle()
        comm = model._dp_comm
        calls, nbytes = C.c_int64(), C.c_int64()
        if comm is not None and h is not None:
            lib.pi05_grad_exchange_stats(h, C.byref(calls), C.byref(nbytes))
            overl_calls, payload = calls.value, nbytes.value
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.pi05_allreduce_grads(h, comm, world, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                           "pi05_allreduce_grads")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = torch.tensor([min(ts)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            exchange = {"payload_gb": payload / 1e9, "collectives_per_backward": overl_calls, "max_ctas": model._dp_max_ctas,
                        "standalone_ms": ms, "standalone_algbw_gbs": payload / 1e9 / (ms / 1e3),
                        "standalone_busbw_gbs": 2.0 * (world - 1) / world * payload / 1e9 / (ms / 1e3),
                        "note": "standalone = the same payload as ONE exchange on an idle GPU through this communicator "
                                "(max_ctas > 0: maxCTAs of the overlapped mode; < 0: minCTAs of the end-of-backward mode)"}
        barrier()

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
            acts = model.sample_actions(dev, Observation.from_dict(d1, keep_uint8=True), num_steps=10)
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
        ent_keep = model._engines[model._train_key]["keep"][0]
        prefix_rows = 3 * (cfg.image_size // cfg.vit_patch) ** 2 + int(ent_keep[2].shape[1])  # image tokens + prompt slots kept
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
                "traffic": ncu_traffic(top["M"], top["N"], top["K"], top["epi"], top["majors"])[0],
                "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r02_ncu_gemm_*.md); "
                                + ncu_traffic(top["M"], top["N"], top["K"], top["epi"], top["majors"])[1],
                "algorithmic_bytes": 2.0 * (top["M"] * top["K"] + top["N"] * top["K"] + top["M"] * top["N"]) * top["batch"],
                "kernel": f"gemm_kernel<256,{top['epi']}> M={top['M']} N={top['N']} K={top['K']} "
                          f"majors={top['majors']} ({top['launches']} launches in the timed region)",
                "peak_source": f"{peaks_kind} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                "all_tcgen05_gemms": {"tflops": gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms else None,
                                      "share_of_step": gemm_ms / ms_dev},
                # model FLOPs (14.02 TFLOP per trained sample, dead last-layer prefix work included) and the
                # FLOPs the engine actually executes (that work is skipped: 13.39 TFLOP), both against the same peak
                "step_model_flops_utilisation": (value / world) * TRAIN_TFLOP_PER_SAMPLE / peak_tf,
                "step_hardware_flops_utilisation": (value / world) * executed_train_tflop_per_sample(prefix_rows) / peak_tf,
                "tflop_per_sample": {"model": TRAIN_TFLOP_PER_SAMPLE, "executed_dense_prefix": EXECUTED_TRAIN_TFLOP_PER_SAMPLE,
                                     "executed": executed_train_tflop_per_sample(prefix_rows), "prefix_rows": prefix_rows},
            }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(world, B, args.small),
            "implementation": {
                "engine": "kai0_b200/libpi05.so (hand-written sm_100a kernels) behind kai0_b200.pi0_pytorch.PI0Pytorch",
                "optimizer": ("kai0_b200.optim.FusedClipAdamW: the clip + AdamW of the step in one engine pass over the 2 "
                              "flat arenas" if use_fused else
                              "torch.optim.AdamW(fused) + clip_grad_norm_(1.0) over "
                              + ("model.parameters()" if args.per_param_optimizer else "model.flat_parameters()")),
                "gradient_exchange": None if world == 1 else model.exchange_description(), "exchange": exchange,
                "prompt_padding": ("removed: the synthetic prompts hold 96 valid of 200 slots (SURVEY.md §8d); the slots that are "
                                   "padding in every sample of the batch are not computed (pi05_batch.token_len), outputs and "
                                   "gradients unchanged (tests/test_engine_gpu.py::test_prompt_padding_removal_is_lossless); "
                                   f"prefix rows {prefix_rows} of 968" if not args.no_skip_padding else
                                   "computed (dense, --no-skip-padding)")},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(host_d, host_a),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss,
                    "clocks": clocks_e2e},
            "gpu_launches": int(n1 - n0),
            "decode": decode,
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1:
            # free the engine (108 GB of workspace) before the reference legs
            del optim
            model._destroy_engine()
            del model
            torch.cuda.empty_cache()
        if world == 1 and not args.small and not args.no_serving:
            line["serving"] = serving_leg()
        if world == 1 and not args.no_cpu_baseline:
            try:  # a failure of the reported CPU context must not cost the measured GPU line
                info = cpu_reference(host_d, host_a, 2, 1, 0.0, whole_first=False)
                line["cpu_baseline"] = {"value": info["samples_per_s"], "unit": UNIT, "cores": info["threads"],
                                        "kind": info["kind"], "sample": info["sample"],
                                        "flop_fraction_per_step": info["scale"], "measured_s_per_step": info["s_per_step"],
                                        "note": "`bench.py --impl reference` additionally times one WHOLE sample"}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": f"unavailable: {type(exc).__name__}: {exc}"[:300]}
        if world == 1 and not args.no_reference_gpu and not args.small:
            RR = _reference_runner()
            if RR is None:
                line["reference_gpu"] = {"unavailable": "no reference package reachable (/root/reference or baseline/_ref)"}
            else:
                try:
                    line["reference_gpu"] = RR.time_gpu(host_d, host_a, dev, steps=3, warmup=2, batch=B)
                    line["reference_gpu"]["speedup_of_this_engine"] = value / line["reference_gpu"]["value"]
                except Exception as exc:  # noqa: BLE001
                    line["reference_gpu"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
