"""Serving path on the B200 (second half): `kai0_b200.serving.Policy` around the engine.

Request dicts in the Agilex client format go through the host transforms (pinned to the reference bit for bit on the
CPU: tests/test_serving_cpu.py), `Observation.from_dict(keep_uint8=True)`, the engine's `sample_actions` (CUDA graph) and
the reply transforms.  Checked here:
  * the model output of a batch of requests against the CPU oracle fed with the SAME transformed inputs (tolerance: the
    action-chunk tolerance of tests/test_engine_gpu.py, x2 because the replies are compared after an affine map);
  * `infer_batch` against one `infer` per request (same weights, other batch size: bf16 noise only); batches of 2 and 1,
    the decode batch sizes the engine's own parity tests cover on this configuration;
  * `RequestBatcher`: concurrent clients get their own replies from shared model calls issued by the worker thread.
"""
import threading

import numpy as np
import pytest
import torch

import helpers as H
from kai0_b200 import serving as S
from oracle import pi05_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

TOL_RAW = 2e-3      # model output (normalised action chunk) vs the oracle: 2x tests/test_engine_gpu.py::TOL_ACTIONS
TOL_BATCH = 2e-3    # batch of 2 vs two batches of 1 (each is within TOL_ACTIONS of the same oracle value)


class _Tokenizer:
    """Deterministic stand-in with RAGGED lengths inside the tiny configuration's 24 prompt slots (the real tokenizer is
    pinned on the CPU; its pi0.5 prompts would fill all 24 slots and hide the padding path)."""

    def __init__(self, max_len, vocab):
        self.max_len, self.vocab = max_len, vocab

    def tokenize(self, prompt, state=None):
        bins = np.digitize(state, bins=np.linspace(-1, 1, 257)[:-1]) - 1
        n = 7 + (sum(map(ord, prompt)) % (self.max_len - 9))
        ids = [2] + [int((ord(prompt[i % len(prompt)]) * 7 + int(bins[i % len(bins)]) + 3 * i) % (self.vocab - 1)) + 1
                     for i in range(n - 1)]
        pad = self.max_len - n
        return np.asarray(ids + [0] * pad), np.asarray([True] * n + [False] * pad)


def _stats():
    g = np.random.default_rng(11)
    out = {}
    for key in ("state", "actions"):
        mean = g.normal(0, 0.3, 32)
        q01, q99 = mean - g.uniform(1.0, 2.0, 32), mean + g.uniform(1.0, 2.0, 32)
        for a in (mean, q01, q99):
            a[14:] = 0.0
        out[key] = S.NormStats(mean=mean, std=np.ones(32), q01=q01, q99=q99)
    return out


def _requests(n=3):
    g = np.random.default_rng(2024)
    cams = ("top_head", "hand_left", "hand_right")
    prompts = ["fold the cloth", "hang the shirt on the hanger", "pick up the cup", "open the drawer"]
    return [{"images": {c: g.integers(0, 256, (3, 90, 120), dtype=np.uint8) for c in cams},
             "state": g.uniform(-1, 1, 14).astype(np.float32), "prompt": prompts[i % len(prompts)]} for i in range(n)]


def _setup():
    oc = O.tiny_config()
    model, params = H.build_pair(oc, seed=5)
    tok = _Tokenizer(oc.max_token_len, oc.vocab_size)
    ins, outs = S.agilex_pi05_transforms(action_dim=oc.action_dim, max_token_len=oc.max_token_len, tokenizer=tok,
                                         norm_stats=_stats(), default_prompt="fold the cloth", image_size=oc.image_size)
    return oc, model, params, ins, outs


def _oracle_raw(oc, params, ins, reqs, noise):
    xs = [S.compose(ins)(S._copy_structure(r)) for r in reqs]
    keys = ("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")
    images = [torch.from_numpy(np.stack([x["image"][k] for x in xs])).to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0
              for k in keys]  # models/model.py:129-133
    masks = [torch.ones(len(xs), dtype=torch.bool) for _ in keys]
    toks = torch.from_numpy(np.stack([x["tokenized_prompt"] for x in xs])).to(torch.int64)
    tmask = torch.from_numpy(np.stack([x["tokenized_prompt_mask"] for x in xs]))
    with torch.no_grad():
        raw = O.sample_actions(params, oc, images, masks, toks, tmask, noise)
    return xs, raw.numpy()


def test_policy_batch_matches_oracle_and_single_requests():
    oc, model, params, ins, outs = _setup()
    reqs = _requests(2)
    noise = torch.randn(2, oc.action_horizon, oc.action_dim, generator=torch.Generator().manual_seed(8))
    xs, raw_ref = _oracle_raw(oc, params, ins, reqs, noise)
    lens = [int(x["tokenized_prompt_mask"].sum()) for x in xs]
    assert len(set(lens)) > 1 and max(lens) < oc.max_token_len  # ragged prompts, padding present

    raw_pol = S.Policy(model, transforms=ins, output_transforms=(), pytorch_device="cuda", max_batch=2)
    got = raw_pol.infer_batch(reqs, noise=[n.numpy() for n in noise])
    raw = np.stack([g["actions"] for g in got])
    assert raw.shape == (2, oc.action_horizon, oc.action_dim) and raw.dtype == np.float32
    err = H.rel_err(torch.from_numpy(raw), torch.from_numpy(raw_ref))
    print(f"serving: batch of 2 vs oracle (normalised chunk) rel {err:.2e}")
    assert err < TOL_RAW
    for i in range(2):  # the state a reply carries is the transformed (normalised, padded) input state
        assert np.array_equal(got[i]["state"], xs[i]["state"])
        assert got[i]["policy_timing"]["batch"] == 2

    pol = S.Policy(model, transforms=ins, output_transforms=outs, pytorch_device="cuda", max_batch=2)
    full = pol.infer_batch(reqs, noise=[n.numpy() for n in noise])
    for i in range(2):
        want = S.compose(outs)({"state": xs[i]["state"].copy(), "actions": raw_ref[i].copy()})["actions"]
        assert full[i]["actions"].shape == (oc.action_horizon, 14)
        e = H.rel_err(torch.from_numpy(full[i]["actions"]), torch.from_numpy(want))
        assert e < 2 * TOL_RAW, (i, e)
        one = pol.infer(reqs[i], noise=noise[i].numpy())["actions"]
        eb = H.rel_err(torch.from_numpy(full[i]["actions"]), torch.from_numpy(one))
        print(f"serving: request {i}: reply vs oracle {e:.2e}, batched vs alone {eb:.2e}")
        assert eb < TOL_BATCH, (i, eb)
    # a second identical call replays the captured graph on refreshed staging blocks: identical replies
    again = pol.infer_batch(reqs, noise=[n.numpy() for n in noise])
    for a, b in zip(full, again):
        assert np.array_equal(a["actions"], b["actions"])


def test_request_batcher_on_the_engine():
    oc, model, params, ins, outs = _setup()
    pol = S.Policy(model, transforms=ins, output_transforms=outs, pytorch_device="cuda", max_batch=2)
    reqs = _requests(4)
    alone = [pol.infer(r)["actions"] for r in reqs]  # internal noise differs per call: compare shapes / finiteness only
    assert all(a.shape == (oc.action_horizon, 14) and np.isfinite(a).all() for a in alone)
    results, errors = {}, []
    with S.RequestBatcher(pol, max_batch=2, max_wait_ms=500.0) as rb:

        def client(i):
            try:
                results[i] = rb.infer(reqs[i])
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        ts = [threading.Thread(target=client, args=(i,)) for i in range(4)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=300)
        assert not errors, errors
        assert sorted(results) == [0, 1, 2, 3]
        assert rb.requests_served == 4 and rb.batches_served < 4  # at least two requests shared a model call
    for i in range(4):
        a = results[i]["actions"]
        assert a.shape == (oc.action_horizon, 14) and np.isfinite(a).all()
        assert results[i]["policy_timing"]["batch"] >= 1
