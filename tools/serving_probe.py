"""Serving-side batching on a B200: latency and throughput of `kai0_b200.serving.Policy.infer_batch` at the FULL pi0.5
architecture for request batches of 1, 2, 4, 8 (host buffers in, host replies out: request transforms, pinned staging,
H2D, CUDA-graph decode, D2H, reply transforms all inside the timed region), plus a parity check of every batch size
against the batch-of-one reply of the same request and noise.

    python tools/serving_probe.py [--out gpurun_out/serving_probe.jsonl] [--batches 1,2,4,8] [--iters 12]

One JSON line per batch size is appended (and flushed) as soon as it is measured, so a run cut short keeps what it has.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kai0_b200 import serving as S  # noqa: E402


class HashTokenizer:
    """Stand-in for paligemma_tokenizer.model (not reachable offline): pi0.5-length prompts (96 valid of 200 slots, SURVEY
    §8d) with ids derived from the prompt text and the discretised state, so every request differs."""

    def __init__(self, max_len: int, vocab: int, valid: int = 96):
        self.max_len, self.vocab, self.valid = max_len, vocab, min(valid, max_len)

    def tokenize(self, prompt, state=None):
        bins = np.digitize(state, bins=np.linspace(-1, 1, 257)[:-1]) - 1
        n = self.valid
        ids = [2] + [int((ord(prompt[i % len(prompt)]) * 131 + int(bins[i % len(bins)]) * 17 + 7 * i) % (self.vocab - 1)) + 1
                     for i in range(n - 1)]
        pad = self.max_len - n
        return np.asarray(ids + [0] * pad), np.asarray([True] * n + [False] * pad)


def make_requests(n: int, image_size: int, seed: int = 0):
    g = np.random.default_rng(seed)
    cams = ("top_head", "hand_left", "hand_right")
    return [{"images": {c: g.integers(0, 256, (3, image_size, image_size), dtype=np.uint8) for c in cams},
             "state": g.uniform(-1, 1, 14).astype(np.float32), "prompt": f"flatten and fold the cloth number {i}"}
            for i in range(n)]


def make_stats(dim: int = 32):
    g = np.random.default_rng(3)
    out = {}
    for key in ("state", "actions"):
        mean = g.normal(0, 0.3, dim)
        q01, q99 = mean - g.uniform(1.0, 2.0, dim), mean + g.uniform(1.0, 2.0, dim)
        for a in (mean, q01, q99):
            a[14:] = 0.0
        out[key] = S.NormStats(mean=mean, std=np.ones(dim), q01=q01, q99=q99)
    return out


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def probe(model, ecfg, device, batches, iters, emit, sync):
    tok = HashTokenizer(ecfg.max_token_len, ecfg.vocab_size)
    ins, outs = S.agilex_pi05_transforms(action_dim=ecfg.action_dim, max_token_len=ecfg.max_token_len, tokenizer=tok,
                                         norm_stats=make_stats(ecfg.action_dim), default_prompt="fold the cloth",
                                         image_size=ecfg.image_size)
    pol = S.Policy(model, transforms=ins, output_transforms=outs, pytorch_device=device, max_batch=max(batches))
    reqs = make_requests(max(batches), ecfg.image_size)
    g = np.random.default_rng(11)
    noise = [g.normal(size=(ecfg.action_horizon, ecfg.action_dim)).astype(np.float32) for _ in reqs]
    alone = {}
    for B in batches:
        rec = {"batch": B}
        try:
            out = pol.infer_batch(reqs[:B], noise=noise[:B])  # warm-up: engine plan, graph capture
            for i in range(B):  # parity: same request + noise served alone (batch of one)
                if i not in alone:
                    alone[i] = pol.infer(reqs[i], noise=noise[i])["actions"]
            rec["rel_err_vs_batch_of_one"] = max(rel(out[i]["actions"], alone[i]) for i in range(B))
            rec["finite"] = bool(all(np.isfinite(o["actions"]).all() for o in out))
            pol.infer_batch(reqs[:B], noise=noise[:B])
            lat, model_ms = [], []
            for _ in range(iters):
                sync()
                t0 = time.perf_counter()
                out = pol.infer_batch(reqs[:B], noise=noise[:B])
                lat.append((time.perf_counter() - t0) * 1e3)
                model_ms.append(out[0]["policy_timing"]["infer_ms"])
            lat.sort()
            model_ms.sort()
            rec.update({"p50_ms": lat[len(lat) // 2], "min_ms": lat[0], "max_ms": lat[-1], "iters": iters,
                        "model_call_p50_ms": model_ms[len(model_ms) // 2],
                        "requests_per_s": B / (lat[len(lat) // 2] / 1e3),
                        "what": "Policy.infer_batch: host request dicts -> transforms -> pinned staging -> H2D -> "
                                "sample_actions (10 steps, CUDA graph) -> D2H -> reply transforms"})
        except Exception as exc:  # noqa: BLE001
            rec["error"] = f"{type(exc).__name__}: {exc}"[:400]
        emit(rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "serving_probe.jsonl"))
    ap.add_argument("--batches", default="1,2,4,8")
    ap.add_argument("--iters", type=int, default=12)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("serving_probe.py needs a B200 (the engine has no CPU path)")
    from kai0_b200.pi0_pytorch import Pi05EngineConfig, PI0Pytorch

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    f = open(args.out, "a")

    def emit(rec):
        f.write(json.dumps(rec) + "\n")
        f.flush()
        os.fsync(f.fileno())
        print(json.dumps(rec), flush=True)

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = Pi05EngineConfig()
    batches = [int(b) for b in args.batches.split(",")]
    t0 = time.time()
    torch.manual_seed(1234)
    model = PI0Pytorch(cfg, max_batch=max(batches), init_weights=False).to(dev)
    model.reset_parameters()
    model.check_inputs = False
    emit({"stage": "model built", "s": time.time() - t0, "gpu": torch.cuda.get_device_name(0)})
    probe(model, model.ecfg, "cuda", batches, args.iters, emit, torch.cuda.synchronize)
    emit({"stage": "done", "s": time.time() - t0})


if __name__ == "__main__":
    main()
