"""Host side of the serving path (second half): kai0_b200/serving.py and kai0_b200/checkpoint.py against
the reference's own code.

Three kinds of checks:
  * fixtures: tests/golden/serving_reference.npz holds what the REFERENCE'S `transforms.py`, `agilex_policy.py`,
    `tokenizer.py`, `normalize.py` and `openpi_client/image_tools.py` produced on seeded requests
    (tools/make_golden_serving.py); this module must reproduce every array bit for bit, dtype included;
  * live: when /root/reference is present the reference code is executed again and compared the same way;
  * the reference's own unit tests (transforms_test.py, shared/normalize_test.py) restated on this module.
The request batcher, the checkpoint directory and the optimiser-state interchange have no counterpart to compare with
and are tested for their contracts.
"""
import json
import os
import sys
import threading
import types

import numpy as np
import pytest
import torch

import helpers as H
from kai0_b200 import checkpoint as CK
from kai0_b200 import serving as S
from oracle import pi05_oracle as O

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import make_golden_serving as MG  # noqa: E402
import reference_serving_loader as RSL  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIX = np.load(os.path.join(GOLD, "serving_reference.npz"))


def _mine():
    lib = types.SimpleNamespace(
        compose=S.compose, InjectDefaultPrompt=S.InjectDefaultPrompt, DeltaActions=S.DeltaActions, Normalize=S.Normalize,
        ResizeImages=S.ResizeImages, TokenizePrompt=S.TokenizePrompt, PadStatesAndActions=S.PadStatesAndActions,
        Unnormalize=S.Unnormalize, AbsoluteActions=S.AbsoluteActions, make_bool_mask=S.make_bool_mask,
        agilex_inputs=lambda d: S.AgilexInputs(action_dim=d, pi05=True), agilex_outputs=S.AgilexOutputs)
    return lib, (lambda n: S.PaligemmaTokenizer(n, model_path=MG.SPM)), S.NormStats


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, f"{what}: {a.dtype}{a.shape} vs {b.dtype}{b.shape}"
    assert np.array_equal(a, b), f"{what}: max |diff| {np.abs(a.astype(np.float64) - b.astype(np.float64)).max()}"


@pytest.mark.parametrize("name,quant", [("quantile", True), ("zscore", False)])
def test_request_and_reply_transforms_reproduce_the_reference_outputs(name, quant):
    lib, tok, ns = _mine()
    inputs, replies = MG.run(lib, tok, ns, use_quantiles=quant)
    got = {}
    MG.flatten_case(name, inputs, replies, got)
    keys = [k for k in FIX.files if k.startswith(name + "/")]
    assert sorted(keys) == sorted(got)
    for k in keys:
        _same(got[k], FIX[k], k)
    # what the chain is expected to have done (agilex_policy.py:95-97, tokenizer.py:24-30, image_tools.py:44-57)
    assert inputs[0]["state"].shape == (32,) and replies[0]["actions"].shape == (MG.HORIZON, 14)
    assert inputs[0]["image"]["base_0_rgb"].shape == (MG.IMAGE_SIZE, MG.IMAGE_SIZE, 3)
    assert inputs[0]["image"]["base_0_rgb"].dtype == np.uint8
    lens = [int(np.asarray(x["tokenized_prompt_mask"]).sum()) for x in inputs]
    if quant:
        assert max(lens) == MG.MAX_TOKEN_LEN and min(lens) < MG.MAX_TOKEN_LEN  # one truncated, the others padded
    assert set(inputs[0]["image"]) == {"base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb"}


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
@pytest.mark.parametrize("quant", [True, False])
def test_request_and_reply_transforms_equal_the_reference_run_live(quant):
    R, rlib, rtok, rns = MG.reference_lib()
    lib, tok, ns = _mine()
    ri, ro = MG.run(rlib, rtok, rns, use_quantiles=quant)
    mi, mo = MG.run(lib, tok, ns, use_quantiles=quant)
    for i, (a, b) in enumerate(zip(mi, ri)):
        fa, fb = S.flatten_dict(a), R.transforms.flatten_dict(b)
        assert list(fa) == list(fb), (list(fa), list(fb))  # same keys in the same order
        for k in fa:
            _same(fa[k], fb[k], f"request {i} {k}")
    for i, (a, b) in enumerate(zip(mo, ro)):
        assert list(a) == list(b) == ["actions"]
        _same(a["actions"], b["actions"], f"reply {i}")


# ------------------------------------------------------------------ the reference's own unit tests, restated
def test_repack_transform():  # transforms_test.py:8-17
    t = S.RepackTransform(structure={"a": {"b": "b/c"}, "d": "e/f"})
    assert t({"b": {"c": 1}, "e": {"f": 2}}) == {"a": {"b": 1}, "d": 2}


def test_delta_and_absolute_actions():  # transforms_test.py:20-67
    item = {"state": np.array([1, 2, 3]), "actions": np.array([[3, 4, 5], [5, 6, 7]])}
    out = S.DeltaActions(mask=[False, True])(item)
    assert np.all(out["state"] == np.array([1, 2, 3]))
    assert np.all(out["actions"] == np.array([[3, 2, 5], [5, 4, 7]]))
    item = {"state": np.array([1, 2, 3]), "actions": np.array([[3, 4, 5], [5, 6, 7]])}
    out = S.AbsoluteActions(mask=[False, True])(item)
    assert np.all(out["actions"] == np.array([[3, 6, 5], [5, 8, 7]]))
    for cls in (S.DeltaActions, S.AbsoluteActions):
        item = {"state": np.array([1, 2, 3]), "actions": np.array([[3, 4, 5], [5, 6, 7]])}
        assert cls(mask=None)(item) is item
        del item["actions"]
        assert cls(mask=[True, False])(item) is item


def test_insert_advantage_into_prompt_is_the_awbc_prompt_format():  # transforms.py:113-121
    out = S.InsertAdvantageIntoPrompt()({"prompt": "fold the cloth", "advantage": 0.123456})
    assert out["prompt"] == "fold the cloth, Advantage: 0.1235"
    with pytest.raises(AssertionError, match="advantage is not in data"):
        S.InsertAdvantageIntoPrompt()({"prompt": "x"})
    if RSL.available():
        R = RSL.load()
        for adv in (0.0, -1.0, 0.99995, np.float32(0.3)):
            a = S.InsertAdvantageIntoPrompt()({"prompt": "p", "advantage": adv})["prompt"]
            assert a == R.transforms.InsertAdvantageIntoPrompt()({"prompt": "p", "advantage": adv})["prompt"]
    # first in the chain when the data config asks for it (training/config.py:431-432)
    ins, _ = S.agilex_pi05_transforms(action_dim=32, max_token_len=8, tokenizer=None, norm_stats=None,
                                      insert_advantage_into_prompt=True)
    assert isinstance(ins[1], S.InsertAdvantageIntoPrompt) and isinstance(ins[2], S.AgilexInputs)


def test_make_bool_mask():  # transforms_test.py:70-72
    assert S.make_bool_mask(2, -2, 2) == (True, True, False, False, True, True)
    assert S.make_bool_mask(2, 0, 2) == (True, True, True, True)


def test_tokenize_prompt():  # transforms_test.py:75-90
    tok = S.PaligemmaTokenizer(max_len=12, model_path=MG.SPM)
    data = S.TokenizePrompt(tok)({"prompt": "Hello, world!"})
    t, m = tok.tokenize("Hello, world!")
    assert np.allclose(t, data["tokenized_prompt"]) and np.allclose(m, data["tokenized_prompt_mask"])
    assert len(t) == 12 and "prompt" not in data
    with pytest.raises(ValueError, match="Prompt is required"):
        S.TokenizePrompt(tok)({})
    with pytest.raises(ValueError, match="State is required"):
        S.TokenizePrompt(tok, discrete_state_input=True)({"prompt": "x"})
    with pytest.raises(ValueError, match="exactly one"):
        S.PaligemmaTokenizer(8)


def test_pi05_prompt_carries_the_discretised_state():  # tokenizer.py:24-30
    tok = S.PaligemmaTokenizer(max_len=200, model_path=MG.SPM)
    import sentencepiece

    sp = sentencepiece.SentencePieceProcessor(model_file=MG.SPM)
    state = np.array([-1.0, -0.999, 0.0, 0.5, 0.9999, 1.0, 3.0])
    ids, mask = tok.tokenize("pick_up the\ncup ", state)
    text = sp.decode([int(i) for i in ids[mask]])
    # 256 bins over [-1, 1): -1 -> 0, 0 -> 128, 0.5 -> 192, values >= 1 - 1/128 -> 255
    # (this SentencePiece model strips the trailing blank of "Action: " when encoding)
    assert text == "Task: pick up the cup, State: 0 0 128 192 255 255 255;\nAction:"
    assert int(ids[0]) == sp.bos_id() and not mask[-1] and int(ids[-1]) == 0


def test_running_stats():  # normalize_test.py:6-44
    arr = np.arange(12).reshape(4, 3)
    rs = S.RunningStats()
    for i in range(len(arr)):
        rs.update(arr[i: i + 1])
    st = rs.get_statistics()
    assert np.allclose(st.mean, arr.mean(0)) and np.allclose(st.std, arr.std(0))
    arr = np.random.default_rng(0).random((2, 3, 4))
    rs = S.RunningStats()
    rs.update(arr)
    st = rs.get_statistics()
    assert np.allclose(st.mean, arr.reshape(-1, 4).mean(0)) and np.allclose(st.std, arr.reshape(-1, 4).std(0))
    with pytest.raises(ValueError, match="less than 2"):
        S.RunningStats().get_statistics()
    with pytest.raises(ValueError, match="does not match"):
        rs.update(np.zeros((2, 5)))


def test_running_stats_equal_the_reference_bit_for_bit():
    rs = S.RunningStats()
    for b in MG.running_stats_batches():  # the second batch widens the range: histograms are re-binned
        rs.update(b)
    st = rs.get_statistics()
    for f in ("mean", "std", "q01", "q99"):
        _same(getattr(st, f), FIX[f"running/{f}"], f"running/{f}")


# ------------------------------------------------------------------ norm_stats.json wire format
def test_norm_stats_wire_format_is_the_references():
    ref_json = FIX["norm_stats_json"].tobytes().decode()
    stats = S.deserialize_json(ref_json)  # written by the reference's pydantic model
    want = MG.norm_stats_arrays()
    for k, rec in want.items():
        for f, v in rec.items():
            _same(getattr(stats[k], f), v, f"{k}.{f}")
    assert stats["no_quantiles"].q01 is None and stats["no_quantiles"].q99 is None
    # and what this module writes is, byte for byte, what the reference wrote
    assert S.serialize_json(stats) == ref_json
    with pytest.raises(ValueError):
        S.deserialize_json(json.dumps({"norm_stats": {"state": {"mean": [0.0]}}}))


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_reference_reads_what_this_module_writes(tmp_path):
    R = RSL.load()
    stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    S.save(tmp_path / "assets" / "agilex", stats)
    back = R.normalize.load(tmp_path / "assets" / "agilex")
    for k in stats:
        for f in ("mean", "std", "q01", "q99"):
            _same(getattr(back[k], f), getattr(stats[k], f), f"{k}.{f}")
    R.normalize.save(tmp_path / "other", back)
    again = S.load(tmp_path / "other")
    assert (tmp_path / "other" / "norm_stats.json").read_text() == (tmp_path / "assets" / "agilex" / "norm_stats.json").read_text()
    _same(again["state"].q99, stats["state"].q99, "round trip")
    with pytest.raises(FileNotFoundError, match="Norm stats file not found"):
        S.load(tmp_path / "missing")


def test_normalize_edge_cases():
    st = {"state": S.NormStats(mean=np.array([1.0, 2.0]), std=np.array([2.0, 4.0]))}
    with pytest.raises(ValueError, match="missing q01 or q99"):
        S.Normalize(st, use_quantiles=True)
    with pytest.raises(ValueError, match="missing q01 or q99"):
        S.Unnormalize(st, use_quantiles=True)
    assert S.Normalize(None)({"state": 1}) == {"state": 1}
    # statistics shorter than the leaf: z-score pads mean 0 / std 1, quantile passes the tail through (transforms.py:177-191)
    x = {"state": np.array([3.0, 6.0, 5.0])}
    out = S.Unnormalize(st)(dict(x))["state"]
    assert np.allclose(out, [3 * (2 + 1e-6) + 1, 6 * (4 + 1e-6) + 2, 5 * (1 + 1e-6)])
    stq = {"state": S.NormStats(mean=np.zeros(2), std=np.ones(2), q01=np.array([-1.0, 0.0]), q99=np.array([1.0, 4.0]))}
    out = S.Unnormalize(stq, use_quantiles=True)(dict(x))["state"]
    assert np.allclose(out, [(3 + 1) / 2 * (2 + 1e-6) - 1, (6 + 1) / 2 * (4 + 1e-6), 5.0])
    with pytest.raises(ValueError, match="Selector key state not found"):
        S.Unnormalize(st)({"actions": np.zeros(2)})
    with pytest.raises(ValueError, match="Selector key state not found"):
        S.Normalize(st, strict=True)({"actions": np.zeros(2)})
    assert "actions" in S.Normalize(st)({"actions": np.zeros(2)})  # non-strict: untouched


def test_agilex_inputs_errors_and_training_fields():
    t = S.AgilexInputs(action_dim=32)
    img = np.zeros((3, 8, 8), np.uint8)
    good = {"images": {"top_head": img, "hand_left": img, "hand_right": img}, "state": np.zeros(14)}
    with pytest.raises(ValueError, match="Expected images to contain"):
        t({**good, "images": {**good["images"], "elbow": img}})
    with pytest.raises(ValueError, match="Camera hand_right not found"):
        t({**good, "images": {"top_head": img, "hand_left": img}})
    out = t({**good, "actions": np.full((50, 14), 4.0), "progress": 0.5,
             "images": {**good["images"], "his_-100_top_head": img}})
    assert out["actions"].shape == (50, 32) and float(np.abs(out["actions"]).max()) == 0.0  # |x| > pi -> 0
    assert "base_-100_rgb" in out["image"] and out["progress"] == 0.5 and "action_mask" not in out
    assert S.AgilexInputs(action_dim=32, pi05=False)({**good, "actions": np.zeros((50, 14))})["action_mask"].all()
    assert float(np.abs(S.AgilexInputs(action_dim=32, mask_state=True)({**good, "state": np.ones(14)})["state"]).max()) == 0


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_arx_inputs_equal_the_references_arx_policy_live():
    """policies/arx_policy.py executed in place: the ARX robot's transform keeps out-of-range state values."""
    import importlib.util

    R = RSL.load()
    spec = importlib.util.spec_from_file_location("openpi.policies.arx_policy", os.path.join(RSL.SRC, "policies", "arx_policy.py"))
    arx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(arx)
    ref_in, my_in = arx.ARXInputs(action_dim=32, model_type=R.ModelType.PI05), S.ARXInputs(32)
    for r in MG.requests():
        r = dict(MG.copy_request(r), actions=np.random.default_rng(0).uniform(-4, 4, (MG.HORIZON, 14)))
        a, b = ref_in(MG.copy_request(r)), my_in(MG.copy_request(r))
        fa, fb = R.transforms.flatten_dict(a), S.flatten_dict(b)
        assert list(fa) == list(fb)
        for k in fa:
            if isinstance(fa[k], str):
                assert fa[k] == fb[k]
            else:
                _same(fb[k], fa[k], k)
    out = my_in(MG.copy_request(MG.requests()[0]))
    assert float(out["state"][3]) == 4.0  # the Agilex transform would have zeroed it
    chunk = {"actions": np.arange(MG.HORIZON * 32, dtype=np.float32).reshape(MG.HORIZON, 32)}
    _same(S.ARXOutputs()(dict(chunk))["actions"], arx.ARXOutputs()(dict(chunk))["actions"], "ARXOutputs")


# ------------------------------------------------------------------ Policy / batching (stub model on the CPU)
class _StubModel:
    """Stands in for PI0Pytorch on the CPU: a deterministic function of every input, so that routing mistakes show."""

    def __init__(self):
        self.calls = []

    def to(self, device):
        return self

    def eval(self):
        return self

    def sample_actions(self, device, obs, noise=None, num_steps=10, **kw):
        B = obs.state.shape[0]
        self.calls.append({"B": B, "kw": {k: (v if not torch.is_tensor(v) else tuple(v.shape)) for k, v in kw.items()},
                           "num_steps": num_steps, "img_dtype": obs.images["base_0_rgb"].dtype})
        base = obs.state.to(torch.float32)[:, None, :].expand(B, MG.HORIZON, 32)
        img = torch.stack([obs.images[k].to(torch.float32).mean(dim=(1, 2, 3)) for k in sorted(obs.images)], 1).sum(1)
        tok = (obs.tokenized_prompt * obs.tokenized_prompt_mask).sum(1).to(torch.float32)
        out = 0.1 * base + (img / 255.0)[:, None, None] + 1e-4 * tok[:, None, None]
        out = out + torch.arange(MG.HORIZON)[None, :, None] * 0.01
        if noise is not None:
            out = out + noise
        if "prev_action_chunk" in kw:
            prev = torch.nn.functional.pad(kw["prev_action_chunk"], (0, 32 - kw["prev_action_chunk"].shape[-1]))
            out = out + 0.5 * prev + kw.get("inference_delay", 0)
        return out.to(torch.float32)


def _policy(model=None, **kw):
    tok = S.PaligemmaTokenizer(MG.MAX_TOKEN_LEN, model_path=MG.SPM)
    stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    ins, outs = S.agilex_pi05_transforms(action_dim=32, max_token_len=MG.MAX_TOKEN_LEN, tokenizer=tok, norm_stats=stats,
                                         default_prompt=MG.DEFAULT_PROMPT, image_size=MG.IMAGE_SIZE)
    model = model or _StubModel()
    return S.Policy(model, transforms=ins, output_transforms=outs, pytorch_device="cpu", metadata={"robot": "agilex"}, **kw), model


def test_policy_infer_is_the_reference_call_sequence():
    """policy.py:68-124: copy -> input transforms -> batch of one -> from_dict -> sample_actions -> [0] -> output
    transforms -> policy_timing; the caller's dict is left alone."""
    pol, model = _policy()
    req = MG.requests()[0]
    keep = MG.copy_request(req)
    out = pol.infer(req)
    assert set(out) == {"actions", "policy_timing"} and out["actions"].shape == (MG.HORIZON, 14)
    assert out["policy_timing"]["infer_ms"] >= 0 and pol.metadata == {"robot": "agilex"}
    assert set(req) == set(keep) and all(np.array_equal(req["images"][c], keep["images"][c]) for c in keep["images"])
    assert model.calls[0]["B"] == 1 and model.calls[0]["img_dtype"] == torch.uint8  # uint8 reaches the engine as is
    # by hand
    lib, tok, ns = _mine()
    stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    ins, outs = S.agilex_pi05_transforms(action_dim=32, max_token_len=MG.MAX_TOKEN_LEN, tokenizer=tok(MG.MAX_TOKEN_LEN),
                                         norm_stats=stats, default_prompt=MG.DEFAULT_PROMPT, image_size=MG.IMAGE_SIZE)
    x = S.compose(ins)(MG.copy_request(keep))
    from kai0_b200.model import Observation

    batch = {"image": {k: torch.from_numpy(v)[None] for k, v in x["image"].items()},
             "image_mask": {k: torch.from_numpy(np.asarray(v))[None] for k, v in x["image_mask"].items()},
             "state": torch.from_numpy(x["state"])[None], "tokenized_prompt": torch.from_numpy(x["tokenized_prompt"])[None],
             "tokenized_prompt_mask": torch.from_numpy(x["tokenized_prompt_mask"])[None]}
    acts = _StubModel().sample_actions("cpu", Observation.from_dict(batch, keep_uint8=True))[0].numpy()
    want = S.compose(outs)({"state": x["state"], "actions": acts})
    _same(out["actions"], want["actions"], "infer vs by hand")
    # noise: [H, A] or [1, H, A] (policy.py:97-102)
    nz = np.random.default_rng(1).normal(size=(MG.HORIZON, 32)).astype(np.float32)
    a = pol.infer(req, noise=nz)["actions"]
    b = pol.infer(req, noise=nz[None])["actions"]
    _same(a, b, "noise rank")
    assert not np.array_equal(a, out["actions"])


class _CanonicalStub(_StubModel):
    """Same function of the inputs whether the images arrive as uint8 NHWC (this repo's Policy) or as the fp32 NCHW tensors
    the reference's `Observation.from_dict` makes of them (models/model.py:129-133)."""

    def sample_actions(self, device, obs, noise=None, num_steps=10, **kw):
        imgs = {k: (v.to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0 if v.dtype == torch.uint8 else v)
                for k, v in obs.images.items()}
        B = obs.state.shape[0]
        kw_seen = {k: (np.asarray(v.cpu()).tolist() if torch.is_tensor(v) else np.asarray(v).tolist()) for k, v in kw.items()}
        self.calls.append({"B": B, "kw": kw_seen, "num_steps": num_steps, "state_dtype": obs.state.dtype})
        img = torch.stack([imgs[k].mean(dim=(1, 2, 3)) for k in sorted(imgs)], 1).sum(1)
        tok = (obs.tokenized_prompt * obs.tokenized_prompt_mask).sum(1).to(torch.float32)
        out = 0.1 * obs.state.to(torch.float32)[:, None, :].expand(B, MG.HORIZON, 32) + img[:, None, None]
        out = out + 1e-4 * tok[:, None, None] + torch.arange(MG.HORIZON)[None, :, None] * 0.01 + 0.001 * num_steps
        if noise is not None:
            out = out + noise
        if "prev_action_chunk" in kw:
            prev = torch.as_tensor(np.asarray(kw["prev_action_chunk"].cpu() if torch.is_tensor(kw["prev_action_chunk"])
                                              else kw["prev_action_chunk"]), dtype=torch.float32)
            prev = prev[None] if prev.dim() == 2 else prev
            out = out + 0.5 * torch.nn.functional.pad(prev, (0, 32 - prev.shape[-1])) + kw.get("inference_delay", 0)
        return out.to(torch.float32)


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_policy_infer_equals_the_references_own_policy_run_live():
    """`openpi/policies/policy.py` and `openpi/models/model.py` executed in place (tools/reference_serving_loader.py):
    the reference's `Policy.infer` with the reference's transforms against this repo's `Policy.infer` with this repo's
    transforms, around the same stand-in model -- replies equal bit for bit, same keyword arguments reach the model."""
    ref_policy, ref_model = RSL.load_policy()
    R, rlib, rtok, rns = MG.reference_lib()
    lib, tok, ns = _mine()
    mask = S.make_bool_mask(*MG.DELTA_MASK_DIMS)

    def chain(L, tokenizer, stats_cls):
        stats = {k: stats_cls(**v) for k, v in MG.norm_stats_arrays().items()}
        ins = [L.InjectDefaultPrompt(MG.DEFAULT_PROMPT), L.agilex_inputs(32), L.DeltaActions(mask),
               L.Normalize(stats, use_quantiles=True), L.InjectDefaultPrompt(MG.DEFAULT_PROMPT),
               L.ResizeImages(MG.IMAGE_SIZE, MG.IMAGE_SIZE), L.TokenizePrompt(tokenizer, discrete_state_input=True),
               L.PadStatesAndActions(32)]
        outs = [L.Unnormalize(stats, use_quantiles=True), L.AbsoluteActions(mask), L.agilex_outputs()]
        return ins, outs

    r_ins, r_outs = chain(rlib, rtok(MG.MAX_TOKEN_LEN), rns)
    m_ins, m_outs = chain(lib, tok(MG.MAX_TOKEN_LEN), ns)
    ref_stub, my_stub = _CanonicalStub(), _CanonicalStub()
    kwargs = {"num_steps": 7}
    ref = ref_policy.Policy(ref_stub, transforms=r_ins, output_transforms=r_outs, sample_kwargs=kwargs,
                            metadata={"robot": "agilex"}, is_pytorch=True, pytorch_device="cpu")
    mine = S.Policy(my_stub, transforms=m_ins, output_transforms=m_outs, sample_kwargs=kwargs,
                    metadata={"robot": "agilex"}, pytorch_device="cpu")
    assert ref.metadata == mine.metadata
    nz = np.random.default_rng(4).normal(size=(MG.HORIZON, 32)).astype(np.float32)
    prev = np.random.default_rng(5).normal(size=(MG.HORIZON, 14)).astype(np.float32)
    for i, req in enumerate(MG.requests()):
        for extra, noise in (({}, None), ({}, nz), ({}, nz[None]),
                             ({"prev_action_chunk": prev, "inference_delay": 3, "execute_horizon": 20}, None)):
            a = ref.infer({**MG.copy_request(req), **extra}, noise=noise)
            b = mine.infer({**MG.copy_request(req), **extra}, noise=noise)
            assert set(a) == set(b) == {"actions", "policy_timing"} and "infer_ms" in b["policy_timing"]
            _same(b["actions"], a["actions"], f"request {i} {sorted(extra)} noise={noise is not None}")
            ca, cb = ref_stub.calls[-1], my_stub.calls[-1]
            assert ca["B"] == cb["B"] == 1 and ca["num_steps"] == cb["num_steps"] == 7
            assert ca["state_dtype"] == cb["state_dtype"]  # float64 after Normalize on both sides
            assert set(ca["kw"]) == set(cb["kw"])
            for k in ("inference_delay", "execute_horizon"):
                assert ca["kw"].get(k) == cb["kw"].get(k)
            if "prev_action_chunk" in extra:  # the reference hands the raw [H, 14] array on, this repo a batch of one
                assert np.array_equal(np.asarray(cb["kw"]["prev_action_chunk"])[0], np.asarray(ca["kw"]["prev_action_chunk"]))
    # Observation.from_dict of the reference (models/model.py:122-157) vs kai0_b200.model.Observation.from_dict
    from kai0_b200.model import Observation

    g = torch.Generator().manual_seed(0)

    def data():
        return {"image": {"base_0_rgb": torch.randint(0, 256, (2, 8, 8, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)),
                          "left_wrist_0_rgb": torch.rand(2, 3, 8, 8, generator=torch.Generator().manual_seed(2))},
                "image_mask": {"base_0_rgb": torch.ones(2, dtype=torch.bool)}, "state": torch.zeros(2, 32),
                "tokenized_prompt": torch.zeros(2, 4, dtype=torch.int64), "tokenized_prompt_mask": torch.ones(2, 4, dtype=torch.bool),
                "progress": torch.rand(2, generator=g), "frame_index": torch.arange(2)}

    ro, mo = ref_model.Observation.from_dict(data()), Observation.from_dict(data())
    for k in ("base_0_rgb", "left_wrist_0_rgb"):
        assert torch.equal(ro.images[k], mo.images[k]) and ro.images[k].dtype == mo.images[k].dtype
    for f in ("state", "tokenized_prompt", "tokenized_prompt_mask", "frame_index", "token_ar_mask", "episode_length"):
        x, y = getattr(ro, f), getattr(mo, f)
        assert (x is None and y is None) or torch.equal(x, y), f
    with pytest.raises(ValueError, match="must be provided together"):
        ref_model.Observation.from_dict({k: v for k, v in data().items() if k != "tokenized_prompt_mask"})
    with pytest.raises(ValueError, match="must be provided together"):
        Observation.from_dict({k: v for k, v in data().items() if k != "tokenized_prompt_mask"})


class _CheckpointableStub(torch.nn.Module):
    """A stand-in with the constructor / checkpoint surface `create_trained_policy` needs (policy_config.py:52-55)."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config
        self.w = torch.nn.Parameter(torch.zeros(3))
        self.paligemma_with_expert = types.SimpleNamespace(to_bfloat16_for_selected_params=lambda p: setattr(self, "cast", p))
        self._impl = _CanonicalStub()
        self._max_batch_hint = None
        self.ecfg = types.SimpleNamespace(action_dim=32, max_token_len=MG.MAX_TOKEN_LEN, image_size=MG.IMAGE_SIZE)

    def sample_actions(self, device, obs, noise=None, num_steps=10, **kw):
        return self._impl.sample_actions(device, obs, noise=noise, num_steps=num_steps, **kw) + self.w.sum()


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_the_references_own_create_trained_policy_builds_the_same_policy(tmp_path):
    """policies/policy_config.py executed in place (tools/reference_serve_harness.py): ITS `create_trained_policy` -- model
    through the reference's `load_pytorch`, dtype cast, norm stats from the checkpoint, the Agilex chain assembled as
    training/config.py does -- against `serving.create_trained_policy` on the same checkpoint directory."""
    import safetensors.torch

    import reference_serve_harness as RSH

    stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    trained = _CheckpointableStub()
    with torch.no_grad():
        trained.w.copy_(torch.tensor([0.25, -0.5, 1.0]))
    step_dir = tmp_path / "7"
    os.makedirs(step_dir)
    safetensors.torch.save_model(trained, str(step_dir / "model.safetensors"))
    S.save(step_dir / "assets" / "agilex", stats)
    R, rlib, rtok, rns = MG.reference_lib()
    mod = types.ModuleType("stand_in")
    mod.PI0Pytorch = _CheckpointableStub
    fields = dict(action_dim=32, action_horizon=MG.HORIZON, max_token_len=MG.MAX_TOKEN_LEN)
    ref = RSH.reference_policy(mod, fields, step_dir, asset_id="agilex", tokenizer=rtok(MG.MAX_TOKEN_LEN),
                               default_prompt=MG.DEFAULT_PROMPT, sample_kwargs={"num_steps": 5}, pytorch_device="cpu",
                               image_size=MG.IMAGE_SIZE, metadata={"robot": "agilex"})
    assert type(ref).__module__ == "openpi.policies.policy" and ref.metadata == {"robot": "agilex"}
    assert ref._model.cast == "bfloat16" and torch.equal(ref._model.w.detach(), trained.w.detach())  # loaded, then cast (:53-55)
    mine = S.create_trained_policy(_CheckpointableStub(), step_dir, asset_id="agilex",
                                   tokenizer=S.PaligemmaTokenizer(MG.MAX_TOKEN_LEN, model_path=MG.SPM),
                                   default_prompt=MG.DEFAULT_PROMPT, sample_kwargs={"num_steps": 5}, pytorch_device="cpu",
                                   metadata={"robot": "agilex"})
    nz = np.random.default_rng(9).normal(size=(MG.HORIZON, 32)).astype(np.float32)
    for i, req in enumerate(MG.requests()):
        a = ref.infer(MG.copy_request(req), noise=nz)
        b = mine.infer(MG.copy_request(req), noise=nz)
        _same(b["actions"], a["actions"], f"request {i}")
    assert ref._model._impl.calls[-1]["num_steps"] == mine._model._impl.calls[-1]["num_steps"] == 5
    with pytest.raises(ValueError, match="Asset id is required"):
        RSH.reference_policy(mod, fields, step_dir, asset_id=None, tokenizer=rtok(MG.MAX_TOKEN_LEN), pytorch_device="cpu")


def test_infer_batch_equals_one_by_one_and_groups_rtc_requests():
    pol, model = _policy()
    reqs = MG.requests()
    single = [pol.infer(r)["actions"] for r in reqs]
    model.calls.clear()
    batched = pol.infer_batch(reqs)
    assert len(model.calls) == 1 and model.calls[0]["B"] == 3  # ONE model call
    for i in range(3):
        _same(batched[i]["actions"], single[i], f"request {i}")
        assert batched[i]["policy_timing"]["batch"] == 3
    # real-time-chunking keys travel as sample_actions kwargs (policy.py:84-90); requests with different scalar settings
    # cannot share a call
    prev = np.random.default_rng(2).normal(size=(MG.HORIZON, 14)).astype(np.float32)
    r = [dict(reqs[0], prev_action_chunk=prev, inference_delay=3, execute_horizon=25),
         dict(reqs[1]),
         dict(reqs[2], prev_action_chunk=2 * prev, inference_delay=3, execute_horizon=25),
         dict(reqs[0], prev_action_chunk=prev, inference_delay=5, execute_horizon=25)]
    model.calls.clear()
    out = pol.infer_batch(r)
    assert sorted(c["B"] for c in model.calls) == [1, 1, 2]
    two = next(c for c in model.calls if c["B"] == 2)
    assert two["kw"]["prev_action_chunk"] == (2, MG.HORIZON, 14) and two["kw"]["inference_delay"] == 3
    _same(out[1]["actions"], single[1], "plain request inside a mixed batch")
    alone = pol.infer(r[2])["actions"]
    _same(out[2]["actions"], alone, "rtc request batched vs alone")
    assert pol.infer_batch([]) == []
    # max_batch: longer request lists are served in slices, and the engine is told to plan for that batch up front
    m2 = _StubModel()
    m2._max_batch_hint = None
    pol2, _ = _policy(m2, max_batch=2)
    model_calls = m2.calls
    out = pol2.infer_batch(reqs)
    assert [c["B"] for c in model_calls] == [2, 1] and m2._max_batch_hint == 2
    for i in range(3):
        _same(out[i]["actions"], single[i], f"sliced batch, request {i}")
    assert S.RequestBatcher(pol2, max_batch=16)._max_batch == 2
    with pytest.raises(ValueError, match="one entry"):
        pol.infer_batch(reqs, noise=[None])


def test_staging_blocks_are_reused_and_grow():
    st = S._Staging("cpu")
    a = st.put("x", [np.arange(4, dtype=np.float32), np.arange(4, dtype=np.float32) + 1])
    blk = st._blocks["x"]
    b = st.put("x", [np.ones(4, dtype=np.float32)])
    assert st._blocks["x"] is blk and tuple(b.shape) == (1, 4) and tuple(a.shape) == (2, 4)
    assert float(a[1, 0]) == 1.0  # earlier result is its own tensor
    st.put("x", [np.zeros(4, dtype=np.float32)] * 5)
    assert st._blocks["x"].shape[0] == 5
    assert st.put("m", [np.True_, np.False_]).dtype == torch.bool
    with pytest.raises(ValueError, match="agree in shape"):
        st.put("x", [np.zeros(4, np.float32), np.zeros(3, np.float32)])


def test_request_batcher_serves_concurrent_clients_in_shared_calls():
    pol, model = _policy()
    reqs = MG.requests()
    single = [pol.infer(r)["actions"] for r in reqs]
    model.calls.clear()
    results = {}
    with S.RequestBatcher(pol, max_batch=4, max_wait_ms=200.0) as rb:
        assert rb.metadata == {"robot": "agilex"}

        def client(i):
            results[i] = rb.infer(reqs[i % 3])["actions"]

        ts = [threading.Thread(target=client, args=(i,)) for i in range(8)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=60)
        assert len(results) == 8
        for i in range(8):
            _same(results[i], single[i % 3], f"client {i}")  # every client got ITS reply
        assert rb.requests_served == 8 and rb.batches_served <= 4  # 8 requests, at most 4 per call, within 200 ms
        assert max(c["B"] for c in model.calls) <= 4
        # a bad request fails alone
        bad = dict(reqs[0], images={"top_head": reqs[0]["images"]["top_head"]})
        f_bad, f_ok = rb.submit(bad), rb.submit(reqs[1])
        with pytest.raises(ValueError, match="not found"):
            f_bad.result(timeout=60)
        _same(f_ok.result(timeout=60)["actions"], single[1], "neighbour of a bad request")
    with pytest.raises(RuntimeError, match="closed"):
        rb.submit(reqs[0])
    with pytest.raises(ValueError):
        S.RequestBatcher(pol, max_batch=0)


# ------------------------------------------------------------------ server wire format
def _reference_msgpack_numpy():
    import importlib.util

    path = os.path.join(RSL.CLIENT, "msgpack_numpy.py")
    spec = importlib.util.spec_from_file_location("_kai0_reference_msgpack_numpy", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_message_framing_and_session_loop():
    """openpi_client/msgpack_numpy.py + websocket_policy_server.py:48-83."""
    req = MG.requests()[0]
    payload = {"images": req["images"], "state": req["state"], "prompt": req["prompt"], "k": np.float32(1.5),
               "flag": np.True_, "nested": {"ids": np.arange(6, dtype=np.int64).reshape(2, 3)}, "plain": [1, 2.5, "x"]}
    frame = S.packb(payload)
    back = S.unpackb(frame)
    assert back["prompt"] == req["prompt"] and back["plain"] == [1, 2.5, "x"]
    _same(back["images"]["top_head"], req["images"]["top_head"], "uint8 image")
    _same(back["state"], req["state"], "state")
    _same(back["nested"]["ids"], payload["nested"]["ids"], "nested int64")
    assert type(back["k"]) is np.float32 and back["k"] == np.float32(1.5) and type(back["flag"]) is np.bool_
    with pytest.raises(ValueError, match="Unsupported dtype"):
        S.packb({"z": np.zeros(2, dtype=np.complex64)})
    with pytest.raises(ValueError, match="Unsupported dtype"):
        S.packb({"o": np.array([None, 1], dtype=object)})
    if RSL.available():  # byte-identical frames, and each side reads the other's
        ref = _reference_msgpack_numpy()
        assert ref.packb(payload) == frame
        assert ref.Packer().pack(payload) == frame
        theirs = ref.unpackb(frame)
        _same(theirs["nested"]["ids"], payload["nested"]["ids"], "reference reads this repo's frame")
        assert type(theirs["k"]) is np.float32

    pol, model = _policy()
    h = S.MessageHandler(pol, {"robot": "agilex"})
    assert S.unpackb(h.greeting()) == {"robot": "agilex"}
    wire = {"images": req["images"], "state": req["state"], "prompt": req["prompt"]}
    r1 = S.unpackb(h.handle(S.packb(wire)))
    _same(r1["actions"], pol.infer(MG.copy_request(req))["actions"], "reply over the wire")
    assert set(r1["server_timing"]) == {"infer_ms"} and "infer_ms" in r1["policy_timing"]
    r2 = S.unpackb(h.handle(S.packb(wire)))
    assert set(r2["server_timing"]) == {"infer_ms", "prev_total_ms"} and not h.closed
    bad = h.handle(S.packb({"images": {}, "state": req["state"]}))  # a failing request: traceback as text, session closed
    assert isinstance(bad, str) and "Traceback" in bad and "not found" in bad and h.closed
    # a RequestBatcher shared by several connections speaks the same protocol
    with S.RequestBatcher(pol, max_batch=4, max_wait_ms=50.0) as rb:
        h2 = S.MessageHandler(rb, rb.metadata)
        _same(S.unpackb(h2.handle(S.packb(wire)))["actions"], r1["actions"], "through the batcher")


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_the_references_own_websocket_client_talks_to_this_serving_stack():
    """openpi_client/websocket_client_policy.py executed in place against tools/serve_policy_b200.py (MessageHandler +
    RequestBatcher under the `websockets` library) on 127.0.0.1: metadata on connect, replies equal to a direct
    `Policy.infer`, a failing request surfaces as the client's RuntimeError with the server's traceback."""
    import asyncio
    import importlib.util

    pytest.importorskip("websockets")
    import serve_policy_b200 as SRV

    RSL.load()  # registers openpi_client (+ image_tools); add its two remaining pure-python modules from where they lie
    for name in ("msgpack_numpy", "base_policy", "websocket_client_policy"):
        full = f"openpi_client.{name}"
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(RSL.CLIENT, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            setattr(sys.modules["openpi_client"], name, mod)
    client_mod = sys.modules["openpi_client.websocket_client_policy"]

    pol, model = _policy()
    reqs = MG.requests()
    direct = [pol.infer(MG.copy_request(r))["actions"] for r in reqs]
    state = {}
    started = threading.Event()

    def run_server():
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
        state["loop"], state["stop"] = loop, asyncio.Event()
        with S.RequestBatcher(pol, max_batch=4, max_wait_ms=20.0) as rb:
            loop.run_until_complete(SRV.serve(rb, "127.0.0.1", 0, {"robot": "agilex"},
                                              ready=lambda port: (state.__setitem__("port", port), started.set()),
                                              stop=state["stop"]))
        loop.close()

    th = threading.Thread(target=run_server, daemon=True)
    th.start()
    assert started.wait(30)
    try:
        client = client_mod.WebsocketClientPolicy("127.0.0.1", state["port"])
        assert client.get_server_metadata() == {"robot": "agilex"}
        for i, r in enumerate(reqs):
            wire = {"images": r["images"], "state": r["state"]}
            if "prompt" in r:
                wire["prompt"] = str(np.asarray(r["prompt"]).item()) if not isinstance(r["prompt"], str) else r["prompt"]
            out = client.infer(wire)
            _same(out["actions"], direct[i], f"request {i} through the reference's client")
            assert "infer_ms" in out["server_timing"] and "infer_ms" in out["policy_timing"]
        second = client_mod.WebsocketClientPolicy("127.0.0.1", state["port"])  # a second connection shares the batcher
        _same(second.infer({"images": reqs[0]["images"], "state": reqs[0]["state"], "prompt": reqs[0]["prompt"]})["actions"],
              direct[0], "second connection")
        with pytest.raises(RuntimeError, match="Error in inference server"):
            second.infer({"images": {}, "state": reqs[0]["state"]})
        import urllib.request

        assert urllib.request.urlopen(f"http://127.0.0.1:{state['port']}/healthz", timeout=10).read() == b"OK\n"
    finally:
        state["loop"].call_soon_threadsafe(state["stop"].set)
        th.join(timeout=30)
    assert not th.is_alive()


# ------------------------------------------------------------------ checkpoint directory
def test_parameters_are_registered_in_the_references_order():
    """optimizer.pt keys its state by position in model.parameters() (train_pytorch.py:170,236-243)."""
    import reference_pin as PIN
    from kai0_b200.pi0_pytorch import AdvantageEstimator, GemmaVariant, PI0Pytorch, Pi05EngineConfig

    want = json.load(open(os.path.join(GOLD, "reference_param_order.json")))
    cfg = Pi05EngineConfig(paligemma_variant=GemmaVariant(*PIN.PG), action_expert_variant=GemmaVariant(*PIN.EX),
                           vit_depth=PIN.VIT_LAYERS, max_token_len=PIN.MAX_TOKEN_LEN, vocab_size=1024)
    for cls in (PI0Pytorch, AdvantageEstimator):
        m = cls(cfg, init_weights=False)
        assert [n for n, _ in m.named_parameters()] == want[cls.__name__]
        # the arenas keep their own layout: the never-trained expert lm_head stays last in the bf16 arena
        last = max((v for v in m._offsets.values() if v[0] == torch.bfloat16), key=lambda v: v[1])
        assert last is m._offsets["paligemma_with_expert.gemma_expert.lm_head.weight"]


def test_checkpoint_directory_round_trip(tmp_path):
    """train_pytorch.py:149-273: <dir>/<step>/{model.safetensors, optimizer.pt, metadata.pt, assets/<id>/norm_stats.json}."""
    oc = O.tiny_config()
    model, _ = H.build_pair(oc, device=None)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)
    g = torch.Generator().manual_seed(0)
    for p in model.parameters():
        p.grad = torch.randn(p.shape, generator=g).to(p.dtype)
    opt.step()
    stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    assert CK.get_latest_checkpoint_step(tmp_path / "nothing") is None
    assert CK.save_checkpoint(model, opt, 10, tmp_path, is_main=False) is None and not os.listdir(tmp_path)
    CK.save_checkpoint(model, opt, 10, tmp_path, norm_stats=stats, asset_id="agilex", config={"name": "pi05_test"})
    os.makedirs(tmp_path / "tmp_30")  # a crashed save must not be mistaken for a step
    final = CK.save_checkpoint(model, opt, 20, tmp_path, norm_stats=stats, asset_id="agilex")
    assert sorted(os.listdir(final)) == ["assets", "metadata.pt", "model.safetensors", "optimizer.pt"]
    assert os.path.exists(os.path.join(final, "assets", "agilex", "norm_stats.json"))
    assert CK.get_latest_checkpoint_step(tmp_path) == 20
    meta = torch.load(tmp_path / "10" / "metadata.pt", weights_only=False)
    assert meta["global_step"] == 10 and meta["config"] == {"name": "pi05_test"} and "timestamp" in meta

    model2, _ = H.build_pair(oc, seed=1, device=None)
    opt2 = torch.optim.AdamW(model2.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)
    assert CK.load_checkpoint(model2, opt2, tmp_path, "cpu") == 20
    for (n, a), (_, b) in zip(model.named_parameters(), model2.named_parameters()):
        assert torch.equal(a, b), n
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys()
    for k in s1:
        assert torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) and torch.equal(s1[k]["exp_avg_sq"], s2[k]["exp_avg_sq"])
    back = CK.load_norm_stats(tmp_path, "agilex")
    _same(back["actions"].q01, stats["actions"].q01, "norm stats from the checkpoint")
    _same(CK.load_norm_stats(final, "agilex")["state"].mean, stats["state"].mean, "step directory given directly")
    with pytest.raises(FileNotFoundError):
        CK.load_checkpoint(model2, opt2, tmp_path / "nothing", "cpu")


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_checkpoint_directory_interchanges_with_the_references_own_functions(tmp_path):
    """scripts/train_pytorch.py executed in place: ITS `save_checkpoint` writes a step this repo's `load_checkpoint`
    restores, and the other way round (weights, stock-AdamW state keyed by parameter position, global step, norm stats)."""
    import dataclasses
    import pathlib

    T = RSL.load_train_script()
    R = RSL.load()

    @dataclasses.dataclass
    class Cfg:  # the fields the three functions read from TrainConfig (train_pytorch.py:155-183)
        checkpoint_dir: pathlib.Path
        save_interval: int = 10
        num_train_steps: int = 100
        wandb_enabled: bool = False
        name: str = "pi05_test"

    oc = O.tiny_config()

    def fresh(seed):
        m, _ = H.build_pair(oc, seed=seed, device=None)
        return m, torch.optim.AdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)

    def step(m, opt, seed):
        g = torch.Generator().manual_seed(seed)
        for n, p in m.named_parameters():
            p.grad = None if n in m._dead_grad_names or "gemma_expert.lm_head" in n else torch.randn(p.shape, generator=g).to(p.dtype)
        opt.step()

    def same(m1, o1, m2, o2):
        for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert torch.equal(a, b), n
        s1, s2 = o1.state_dict()["state"], o2.state_dict()["state"]
        assert s1.keys() == s2.keys() and len(s1) > 50
        for k in s1:
            assert torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) and torch.equal(s1[k]["exp_avg_sq"], s2[k]["exp_avg_sq"])
            assert float(s1[k]["step"]) == float(s2[k]["step"])

    ref_stats = {k: R.normalize.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    my_stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    data_config = types.SimpleNamespace(norm_stats=ref_stats, asset_id="agilex")

    # the reference writes, this repo reads
    d1 = tmp_path / "by_reference"
    m, opt = fresh(1)
    step(m, opt, 7)
    cfg = Cfg(checkpoint_dir=d1)
    T.save_checkpoint(m, opt, 15, cfg, True, data_config)  # not on the schedule (15 % 10): nothing written
    assert not d1.exists() or not os.listdir(d1)
    T.save_checkpoint(m, opt, 20, cfg, True, data_config)
    assert sorted(os.listdir(d1 / "20")) == ["assets", "metadata.pt", "model.safetensors", "optimizer.pt"]
    m2, opt2 = fresh(2)
    assert CK.load_checkpoint(m2, opt2, d1, "cpu") == 20
    same(m, opt, m2, opt2)
    _same(CK.load_norm_stats(d1, "agilex")["actions"].q99, my_stats["actions"].q99, "norm stats written by the reference")

    # this repo writes, the reference reads
    d2 = tmp_path / "by_this_repo"
    step(m2, opt2, 8)
    CK.save_checkpoint(m2, opt2, 30, d2, norm_stats=my_stats, asset_id="agilex", config=Cfg(checkpoint_dir=d2))
    os.makedirs(d2 / "tmp_40")
    os.makedirs(d2 / "notes")
    assert T.get_latest_checkpoint_step(d2) == CK.get_latest_checkpoint_step(d2) == 30
    m3, opt3 = fresh(3)
    assert T.load_checkpoint(m3, opt3, d2, torch.device("cpu")) == 30
    same(m2, opt2, m3, opt3)
    back = R.normalize.load(d2 / "30" / "assets" / "agilex")
    _same(back["state"].q01, my_stats["state"].q01, "norm stats read by the reference")
    # the files themselves: same names, and the norm-stats JSON is byte-identical
    assert (d1 / "20" / "assets" / "agilex" / "norm_stats.json").read_text() == \
        (d2 / "30" / "assets" / "agilex" / "norm_stats.json").read_text()
    meta1 = torch.load(d1 / "20" / "metadata.pt", weights_only=False)
    meta2 = torch.load(d2 / "30" / "metadata.pt", weights_only=False)
    assert set(meta1) == set(meta2) == {"global_step", "config", "timestamp"}
    assert set(meta1["config"]) == set(meta2["config"])


@pytest.mark.skipif(not RSL.available(), reason="needs /root/reference (build container)")
def test_a_checkpoint_of_the_references_own_model_and_optimizer_resumes_on_this_module(tmp_path):
    """End to end: the REFERENCE'S `PI0Pytorch` (pin configuration, executed in place) + stock AdamW, saved by the
    reference's own `save_checkpoint`; restored with kai0_b200.checkpoint into this repo's module + stock AdamW, and into
    the fused optimiser.  Every parameter and every optimiser moment must land on the parameter of the SAME NAME --
    optimizer.pt is keyed by position in model.parameters(), hence the reference registration order of this module."""
    import dataclasses
    import pathlib

    import make_golden_reference as G
    import reference_pin as PIN
    from kai0_b200.optim import FusedClipAdamW
    from kai0_b200.pi0_pytorch import GemmaVariant, PI0Pytorch, Pi05EngineConfig

    p0, ref = G.build_reference("bfloat16")
    pw = ref.paligemma_with_expert.paligemma
    pw.lm_head.weight = pw.model.language_model.embed_tokens.weight  # tie_weights() of transformers 4.53.2 (5.5 here skips it)
    names = [n for n, _ in ref.named_parameters()]
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for _, p in ref.named_parameters():
            p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.dtype))
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)
    for n, p in ref.named_parameters():  # the unused expert head never gets a gradient (find_unused_parameters, :441-447)
        p.grad = None if "gemma_expert.lm_head" in n else (torch.randn(p.shape, generator=g) * 1e-2).to(p.dtype)
    opt.step()

    @dataclasses.dataclass
    class Cfg:
        checkpoint_dir: pathlib.Path
        save_interval: int = 10
        num_train_steps: int = 100
        wandb_enabled: bool = False

    T = RSL.load_train_script()
    T.save_checkpoint(ref, opt, 10, Cfg(tmp_path), True, types.SimpleNamespace(norm_stats=None, asset_id=None))
    cfg = Pi05EngineConfig(paligemma_variant=GemmaVariant(*PIN.PG), action_expert_variant=GemmaVariant(*PIN.EX),
                           vit_depth=PIN.VIT_LAYERS, max_token_len=PIN.MAX_TOKEN_LEN)
    mine = PI0Pytorch(cfg, init_weights=False)
    mopt = torch.optim.AdamW(mine.parameters(), lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)
    assert CK.load_checkpoint(mine, mopt, tmp_path, "cpu") == 10
    rp, mp = dict(ref.named_parameters()), dict(mine.named_parameters())
    assert set(rp) == set(mp)
    for n in mp:
        assert mp[n].dtype == rp[n].dtype and torch.equal(mp[n], rp[n]), n
    rs, ms = opt.state_dict()["state"], mopt.state_dict()["state"]
    my_names = [n for n, _ in mine.named_parameters()]
    assert my_names == names and rs.keys() == ms.keys() and len(rs) == len(names) - 1
    for i in rs:
        assert ms[i]["exp_avg"].shape == mp[names[i]].shape
        assert torch.equal(ms[i]["exp_avg"], rs[i]["exp_avg"]) and torch.equal(ms[i]["exp_avg_sq"], rs[i]["exp_avg_sq"]), names[i]
    # the fused optimiser takes the same file
    sd = torch.load(tmp_path / "10" / "optimizer.pt", weights_only=False)
    for n in mine._dead_grad_names:  # parameters this engine gives no gradient have state in THIS synthetic file only
        sd["state"].pop(names.index(n), None)
    fused = FusedClipAdamW(mine)
    fused.load_state_dict(sd)
    assert fused.step_count == 1
    for i, rec in sd["state"].items():
        dt, off, n, shape = mine._offsets[names[i]]
        arena = 0 if dt == torch.bfloat16 else 1
        assert torch.equal(fused.m[arena][off:off + n].view(shape), rec["exp_avg"]), names[i]


def test_create_trained_policy_from_a_checkpoint_directory(tmp_path):
    """policy_config.py:16-94: weights from model.safetensors, dtype map, norm stats FROM THE CHECKPOINT (not the config's
    assets), the transform chain in the reference's order; and the engine-backed policy has no CPU fallback."""
    oc = O.tiny_config()
    trained, _ = H.build_pair(oc, seed=11, device=None)
    stats = {k: S.NormStats(**v) for k, v in MG.norm_stats_arrays().items()}
    step_dir = CK.save_checkpoint(trained, None, 500, tmp_path, norm_stats=stats, asset_id="agilex")
    assert not os.path.exists(os.path.join(step_dir, "optimizer.pt"))  # optimizer=None: weights + metadata + assets only
    from kai0_b200.pi0_pytorch import PI0Pytorch

    fresh = PI0Pytorch(H.engine_config(oc))
    tok = S.PaligemmaTokenizer(oc.max_token_len, model_path=MG.SPM)
    pol = S.create_trained_policy(fresh, step_dir, asset_id="agilex", tokenizer=tok, default_prompt="fold the cloth",
                                  pytorch_device="cpu", metadata={"reset_pose": [0.0] * 14}, sample_kwargs={"num_steps": 10})
    for (n, a), (_, b) in zip(trained.named_parameters(), fresh.named_parameters()):
        assert torch.equal(a, b), n
    assert pol.metadata == {"reset_pose": [0.0] * 14}
    ins = [type(t).__name__ for t in pol._input_transform.__closure__[0].cell_contents]
    outs = [type(t).__name__ for t in pol._output_transform.__closure__[0].cell_contents]
    assert ins == ["InjectDefaultPrompt", "AgilexInputs", "DeltaActions", "Normalize", "InjectDefaultPrompt", "ResizeImages",
                   "TokenizePrompt", "PadStatesAndActions"]                      # policy_config.py:75-81
    assert outs == ["Unnormalize", "AbsoluteActions", "AgilexOutputs"]            # :82-87
    norm = pol._input_transform.__closure__[0].cell_contents[3]
    assert norm.use_quantiles and np.array_equal(norm.norm_stats["state"].q01, stats["state"].q01)
    req = MG.requests()[2]
    with pytest.raises(RuntimeError, match="no CPU path"):
        pol.infer(req)  # the request went through every transform and reached the engine, which refuses to run on a CPU
    with pytest.raises(ValueError, match="Asset id is required"):
        S.create_trained_policy(fresh, step_dir, asset_id=None, tokenizer=tok, pytorch_device="cpu")
    with pytest.raises(FileNotFoundError, match="model.safetensors"):
        S.create_trained_policy(fresh, tmp_path / "nowhere", asset_id="agilex", tokenizer=tok, pytorch_device="cpu")
    with pytest.raises(FileNotFoundError, match="Norm stats file not found"):
        S.create_trained_policy(fresh, step_dir, asset_id="other_robot", tokenizer=tok, pytorch_device="cpu")


def test_model_safetensors_loads_whichever_tied_name_the_file_carries(tmp_path):
    """safetensors.save_model keeps ONE name of the tied embed_tokens / lm_head pair (train_pytorch.py:167); a file written
    by either side, with either name, must restore the shared table."""
    import safetensors.torch
    from safetensors import safe_open

    E = "paligemma_with_expert.paligemma.model.language_model.embed_tokens.weight"
    L = "paligemma_with_expert.paligemma.lm_head.weight"
    oc = O.tiny_config()
    model, _ = H.build_pair(oc, device=None)
    sd = model.state_dict()
    safetensors.torch.save_model(model, str(tmp_path / "ours.safetensors"))
    with safe_open(str(tmp_path / "ours.safetensors"), "pt") as f:
        keys = set(f.keys())
    assert (E in keys) != (L in keys) and len(keys) == len(sd) - 1
    for drop in (E, L):
        path = str(tmp_path / f"without_{drop.split('.')[-2]}.safetensors")
        safetensors.torch.save_file({k: v.contiguous() for k, v in sd.items() if k != drop}, path)
        other, _ = H.build_pair(oc, seed=9, device=None)
        missing, unexpected = safetensors.torch.load_model(other, path)
        assert not missing and not unexpected
        named = dict(other.named_parameters())
        assert torch.equal(named[E], sd[E])
        assert other.paligemma_with_expert.paligemma.lm_head.weight is named[E]  # still ONE parameter


def test_fused_optimizer_state_interchanges_with_stock_adamw():
    """`FusedClipAdamW.state_dict(format="torch")` is what torch.optim.AdamW over model.parameters() would hold, and a
    stock optimizer.pt loads into the fused optimiser (moments land in the right arena slices)."""
    from kai0_b200.optim import FusedClipAdamW

    oc = O.tiny_config()
    model, _ = H.build_pair(oc, device=None)
    stock = torch.optim.AdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=1e-10)
    g = torch.Generator().manual_seed(3)
    named = dict(model.named_parameters())
    dead = set(model._dead_grad_names) | {"paligemma_with_expert.gemma_expert.lm_head.weight"}
    for n, p in named.items():  # the reference's autograd leaves these without a gradient: no optimiser state
        p.grad = None if n in dead else torch.randn(p.shape, generator=g).to(p.dtype)
    stock.step()
    stock.step()
    sd = stock.state_dict()
    fused = FusedClipAdamW(model, lr=1.0)
    fused.load_state_dict(sd)
    assert fused.step_count == 2 and fused.param_groups[0]["lr"] == 3e-4 and fused.param_groups[0]["max_norm"] == 1.0
    names = [n for n, _ in model.named_parameters()]
    for i, rec in sd["state"].items():
        dt, off, n, shape = model._offsets[names[i]]
        arena = 0 if dt == torch.bfloat16 else 1
        assert torch.equal(fused.m[arena][off:off + n].view(shape), rec["exp_avg"]), names[i]
        assert torch.equal(fused.v[arena][off:off + n].view(shape), rec["exp_avg_sq"]), names[i]
    # everything outside the loaded slices is zero (alignment gaps, parameters without state)
    total = sum(float(rec["exp_avg"].float().abs().sum()) for rec in sd["state"].values())
    assert abs(sum(float(m.float().abs().sum()) for m in fused.m) - total) <= 1e-3 * total
    # and back: the exported record loads into a fresh stock optimiser and equals the original
    out = fused.state_dict(format="torch")
    assert out["state"].keys() == sd["state"].keys()
    assert out["param_groups"][0]["params"] == sd["param_groups"][0]["params"]
    assert {k: v for k, v in out["param_groups"][0].items() if k != "params"} == \
        {k: v for k, v in sd["param_groups"][0].items() if k != "params"}
    fresh = torch.optim.AdamW(model.parameters(), lr=1.0)
    fresh.load_state_dict(out)
    s2 = fresh.state_dict()["state"]
    for k, rec in sd["state"].items():
        assert torch.equal(s2[k]["exp_avg"], rec["exp_avg"]) and float(s2[k]["step"]) == 2.0
    # the compact layout still round-trips
    flat = fused.state_dict()
    other = FusedClipAdamW(model)
    other.load_state_dict(flat)
    assert other.step_count == 2 and torch.equal(other.m[0], fused.m[0]) and torch.equal(other.v[1], fused.v[1])
    # a record that cannot be represented is refused
    bad = {"state": {**sd["state"], 0: {**sd["state"][0], "step": torch.tensor(5.0)}}, "param_groups": sd["param_groups"]}
    with pytest.raises(ValueError, match="step counts differ"):
        fused.load_state_dict(bad)
    dead_idx = names.index(next(iter(model._dead_grad_names)))
    bad = {"state": {dead_idx: sd["state"][0]}, "param_groups": sd["param_groups"]}
    with pytest.raises(ValueError, match="does not update"):
        fused.load_state_dict(bad)
    assert FusedClipAdamW(model).state_dict(format="torch")["state"] == {}  # no step taken yet: no state, like torch
