"""Golden outputs of the REFERENCE'S OWN serving-side host code (transforms.py, shared/normalize.py,
models/tokenizer.py, policies/agilex_policy.py, openpi_client/image_tools.py -- executed in place through
tools/reference_serving_loader.py) on seeded requests.  Build container only.  Writes

    tests/golden/serving_spm_tiny.model   a 400-piece SentencePiece model trained here on synthetic prompts (the real
                                          paligemma_tokenizer.model lives in a bucket this image cannot reach)
    tests/golden/serving_reference.npz    outputs only (requests and statistics are regenerated from seeds)

    python tools/make_golden_serving.py
"""
import io
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SPM = os.path.join(GOLD, "serving_spm_tiny.model")
OUT = os.path.join(GOLD, "serving_reference.npz")

IMAGE_SIZE = 56
MAX_TOKEN_LEN = 68  # quantile case: prompts of 66 / 70 / 61 pieces (two padded, one truncated); z-score case: all truncated
ACTION_DIM = 32
HORIZON = 50
DEFAULT_PROMPT = "flatten and fold the cloth"
DELTA_MASK_DIMS = (6, -1, 6, -1)


def train_tokenizer() -> bytes:
    import sentencepiece as spm

    rnd = random.Random(0)
    verbs = ["pick", "place", "fold", "flatten", "hang", "open", "close", "push", "pull", "grasp", "lift", "wipe", "stack"]
    objs = ["cloth", "shirt", "towel", "box", "cup", "drawer", "block", "bottle", "hanger", "basket", "lid", "sponge"]
    lines = []
    for _ in range(4000):
        states = " ".join(str(rnd.randrange(256)) for _ in range(32))
        lines.append(f"Task: {rnd.choice(verbs)} the {rnd.choice(objs)} and {rnd.choice(verbs)} the {rnd.choice(objs)}, "
                     f"State: {states}; Action: ")
        lines.append(f"{rnd.choice(verbs)} {rnd.choice(objs)} Advantage: {rnd.random():.4f}")
    buf = io.BytesIO()
    spm.SentencePieceTrainer.train(sentence_iterator=iter(lines), model_writer=buf, vocab_size=400, model_type="bpe",
                                   character_coverage=1.0, user_defined_symbols=["\n"], bos_id=2, eos_id=1, unk_id=3,
                                   pad_id=0, byte_fallback=True, normalization_rule_name="identity", minloglevel=2)
    return buf.getvalue()


def requests():
    """Three seeded raw requests in the Agilex client format (agilex_policy.py:18-24)."""
    g = np.random.default_rng(20250924)
    cams = ("top_head", "hand_left", "hand_right")
    r0 = {"images": {c: g.integers(0, 256, (3, 120, 160), dtype=np.uint8) for c in cams},
          "state": g.uniform(-1, 1, 14).astype(np.float32), "prompt": "fold_the cloth\n now"}
    r0["state"][3] = 4.0  # beyond pi: zeroed by AgilexInputs
    r1 = {"images": {c: g.random((3, 90, 90)).astype(np.float32) for c in cams}, "state": g.uniform(-1.5, 1.5, 14)}
    r2 = {"images": {c: g.integers(0, 256, (IMAGE_SIZE, IMAGE_SIZE, 3), dtype=np.uint8) for c in cams},
          "state": g.uniform(-1, 1, 14).astype(np.float32), "prompt": np.asarray("hang the shirt")}
    r2["state"][0] = -3.5
    return [r0, r1, r2]


def norm_stats_arrays():
    g = np.random.default_rng(7)
    out = {}
    for key in ("state", "actions"):
        mean = g.normal(0, 0.3, ACTION_DIM)
        std = g.uniform(0.2, 1.0, ACTION_DIM)
        q01 = mean - g.uniform(1.0, 2.0, ACTION_DIM)
        q99 = mean + g.uniform(1.0, 2.0, ACTION_DIM)
        for a in (mean, std, q01, q99):
            a[14:] = 0.0  # padded dimensions carry zeros in real statistics
        out[key] = dict(mean=mean, std=std, q01=q01, q99=q99)
    return out


def model_outputs():
    """A seeded stand-in for what the model returns for each request: normalised action chunk [H, 32]."""
    g = np.random.default_rng(99)
    return [g.normal(0, 0.5, (HORIZON, ACTION_DIM)).astype(np.float32) for _ in range(3)]


def running_stats_batches():
    g = np.random.default_rng(5)
    return [g.normal(0, 1, (64, 5)), g.normal(0.5, 2.0, (3, 7, 5)), g.normal(0, 0.1, (16, 5))]


def copy_request(r):
    return {k: ({c: v.copy() for c, v in val.items()} if isinstance(val, dict) else
                (val.copy() if isinstance(val, np.ndarray) else val)) for k, val in r.items()}


def run(lib, tokenizer_cls, ns_cls, *, use_quantiles: bool):
    """`lib`: a namespace with the transform classes (the reference's `openpi.transforms` + agilex classes, or
    kai0_b200.serving).  Returns (per-request model inputs, per-request replies)."""
    stats = {k: ns_cls(**v) for k, v in norm_stats_arrays().items()}
    mask = lib.make_bool_mask(*DELTA_MASK_DIMS)
    tok = tokenizer_cls(MAX_TOKEN_LEN)
    ins = lib.compose([lib.InjectDefaultPrompt(DEFAULT_PROMPT), lib.agilex_inputs(ACTION_DIM), lib.DeltaActions(mask),
                       lib.Normalize(stats, use_quantiles=use_quantiles), lib.InjectDefaultPrompt(DEFAULT_PROMPT),
                       lib.ResizeImages(IMAGE_SIZE, IMAGE_SIZE), lib.TokenizePrompt(tok, discrete_state_input=True),
                       lib.PadStatesAndActions(ACTION_DIM)])
    outs = lib.compose([lib.Unnormalize(stats, use_quantiles=use_quantiles), lib.AbsoluteActions(mask), lib.agilex_outputs()])
    inputs, replies = [], []
    for r, acts in zip(requests(), model_outputs()):
        x = ins(copy_request(r))
        inputs.append(x)
        replies.append(outs({"state": np.asarray(x["state"]).copy(), "actions": acts.copy()}))
    return inputs, replies


def reference_lib():
    import types

    import reference_serving_loader as L

    if not os.path.exists(SPM):
        os.makedirs(GOLD, exist_ok=True)
        with open(SPM, "wb") as f:
            f.write(train_tokenizer())
    L.set_tokenizer_model(SPM)
    R = L.load()
    T = R.transforms
    lib = types.SimpleNamespace(
        compose=T.compose, InjectDefaultPrompt=T.InjectDefaultPrompt, DeltaActions=T.DeltaActions, Normalize=T.Normalize,
        ResizeImages=T.ResizeImages, TokenizePrompt=T.TokenizePrompt, PadStatesAndActions=T.PadStatesAndActions,
        Unnormalize=T.Unnormalize, AbsoluteActions=T.AbsoluteActions, make_bool_mask=T.make_bool_mask,
        agilex_inputs=lambda d: R.agilex_policy.AgilexInputs(action_dim=d, model_type=R.ModelType.PI05),
        agilex_outputs=R.agilex_policy.AgilexOutputs)
    return R, lib, R.tokenizer.PaligemmaTokenizer, R.normalize.NormStats


def flatten_case(prefix, inputs, replies, store):
    for i, (x, y) in enumerate(zip(inputs, replies)):
        for k, v in x["image"].items():
            store[f"{prefix}/in{i}/image/{k}"] = np.asarray(v)
            store[f"{prefix}/in{i}/image_mask/{k}"] = np.asarray(x["image_mask"][k])
        for k in ("state", "tokenized_prompt", "tokenized_prompt_mask"):
            store[f"{prefix}/in{i}/{k}"] = np.asarray(x[k])
        store[f"{prefix}/out{i}/actions"] = np.asarray(y["actions"])


def main():
    R, lib, tok_cls, ns_cls = reference_lib()
    store = {}
    for name, q in (("quantile", True), ("zscore", False)):
        inputs, replies = run(lib, tok_cls, ns_cls, use_quantiles=q)
        flatten_case(name, inputs, replies, store)
    # wire format + streaming statistics
    stats = {k: ns_cls(**v) for k, v in norm_stats_arrays().items()}
    stats["no_quantiles"] = ns_cls(mean=np.arange(3.0), std=np.ones(3))
    store["norm_stats_json"] = np.frombuffer(R.normalize.serialize_json(stats).encode(), dtype=np.uint8)
    rs = R.normalize.RunningStats()
    for b in running_stats_batches():
        rs.update(b)
    st = rs.get_statistics()
    for f in ("mean", "std", "q01", "q99"):
        store[f"running/{f}"] = np.asarray(getattr(st, f))
    np.savez_compressed(OUT, **store)
    print(f"wrote {OUT}: {len(store)} arrays, {os.path.getsize(OUT) / 1e3:.1f} kB; tokenizer {os.path.getsize(SPM)} B")


if __name__ == "__main__":
    main()
