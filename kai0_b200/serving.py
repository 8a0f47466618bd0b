"""Serving side of the pi0.5 path (second half): everything between a client's observation dict and
`PI0Pytorch.sample_actions`, and between the action chunk and the reply.

Mirrors, for the PyTorch pi0.5 policy of the reference:
  * `openpi.policies.policy.Policy`           (src/openpi/policies/policy.py:23-129)      -> `Policy`
  * `policy_config.create_trained_policy`     (src/openpi/policies/policy_config.py:16-94) -> `create_trained_policy`
  * the transforms that policy chains          (src/openpi/transforms.py)                   -> same class names, below
  * `AgilexInputs` / `AgilexOutputs`           (src/openpi/policies/agilex_policy.py)
  * `PaligemmaTokenizer`                       (src/openpi/models/tokenizer.py:13-47)
  * the norm-stats wire format                 (src/openpi/shared/normalize.py:123-146)     -> `NormStats`, `save`, `load`
  * the server's message format and loop       (openpi_client/msgpack_numpy.py, serving/websocket_policy_server.py:48-83)
                                                                                             -> `packb` / `unpackb`, `MessageHandler`

What is B200-native here is the request path around the engine, not the arithmetic (which is a few hundred host flops):
the reference serves ONE observation per model call and moves every leaf to the device with its own blocking copy
(policy.py:78); `Policy.infer_batch` stages any number of concurrent requests into one pinned host block per field, issues
the copies asynchronously on the engine's stream, runs ONE CUDA-graph replay of `sample_actions` at batch B (the decode is
bound by streaming the expert's weights, so B requests cost about as much as one) and reads the B action chunks back
with one copy.  `RequestBatcher` turns that into a drop-in `infer(obs)` for a multi-client server: callers block on a
future while a worker thread collects up to `max_batch` requests within `max_wait_ms`.

All transforms work on the host with numpy exactly as the reference's do (same operations in the same order, so the
tests compare bit for bit against the reference's own functions: tests/test_serving_cpu.py).
"""
from __future__ import annotations

import dataclasses
import json
import os
import queue
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, Mapping, Sequence

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------------------------
# nested dict <-> "a/b/c" paths  (transforms.py:372-379: flax.traverse_util.flatten_dict(sep="/") and its inverse)
# ---------------------------------------------------------------------------------------------------------------


def flatten_dict(tree: Mapping) -> dict:
    """Leaves of a nested dict keyed by their '/'-joined path, in depth-first key order; empty sub-dicts vanish."""
    out: dict = {}

    def walk(node, prefix):
        for key, val in node.items():
            path = f"{prefix}/{key}" if prefix else str(key)
            if isinstance(val, Mapping):
                walk(val, path)
            else:
                out[path] = val

    walk(tree, "")
    return out


def unflatten_dict(flat: Mapping[str, Any]) -> dict:
    root: dict = {}
    for path, val in flat.items():
        node = root
        *parents, leaf = path.split("/")
        for p in parents:
            node = node.setdefault(p, {})
        node[leaf] = val
    return root


def apply_tree(tree: Mapping, selector: Mapping, fn: Callable, *, strict: bool = False) -> dict:
    """transforms.py:436-452: fn(leaf, selector_leaf) on every path the two trees share; `strict` demands that every
    selector path exists in the tree."""
    flat, sel = flatten_dict(tree), flatten_dict(selector)
    if strict:
        for path in sel:
            if path not in flat:
                raise ValueError(f"Selector key {path} not found in tree")
    return unflatten_dict({p: (fn(v, sel[p]) if p in sel else v) for p, v in flat.items()})


def pad_to_dim(x: np.ndarray, target_dim: int, axis: int = -1, value: float = 0.0) -> np.ndarray:
    """transforms.py:455-462: constant-pad `axis` up to `target_dim`; longer inputs pass through untouched."""
    missing = target_dim - x.shape[axis]
    if missing <= 0:
        return x
    widths = [(0, 0)] * x.ndim
    widths[axis] = (0, missing)
    return np.pad(x, widths, constant_values=value)


def make_bool_mask(*dims: int) -> tuple:
    """transforms.py:465-484: make_bool_mask(2, -2, 2) == (T, T, F, F, T, T); a zero contributes nothing."""
    bits: list = []
    for d in dims:
        bits += [d > 0] * abs(d)
    return tuple(bits)


# ---------------------------------------------------------------------------------------------------------------
# normalisation statistics and their wire format: <checkpoint>/assets/<asset_id>/norm_stats.json
# ---------------------------------------------------------------------------------------------------------------


@dataclasses.dataclass
class NormStats:
    """shared/normalize.py:9-14.  Arrays are float64 after a JSON round trip, as in the reference."""

    mean: np.ndarray
    std: np.ndarray
    q01: np.ndarray | None = None
    q99: np.ndarray | None = None

    def __post_init__(self):
        for f in ("mean", "std", "q01", "q99"):
            v = getattr(self, f)
            if v is not None and not isinstance(v, np.ndarray):
                setattr(self, f, np.asarray(v))


def serialize_json(norm_stats: Mapping[str, NormStats]) -> str:
    """normalize.py:123-125: {"norm_stats": {key: {"mean": [...], "std": [...], "q01": [...]|null, "q99": ...}}},
    two-space indentation.  Readable by the reference's `deserialize_json` (and vice versa)."""

    def enc(a):
        return None if a is None else np.asarray(a).tolist()

    body = {k: {"mean": enc(s.mean), "std": enc(s.std), "q01": enc(s.q01), "q99": enc(s.q99)} for k, s in norm_stats.items()}
    return json.dumps({"norm_stats": body}, indent=2)


def deserialize_json(data: str) -> dict:
    """normalize.py:128-130."""
    body = json.loads(data)["norm_stats"]
    out = {}
    for key, rec in body.items():
        unknown = set(rec) - {"mean", "std", "q01", "q99"}
        if unknown or "mean" not in rec or "std" not in rec:
            raise ValueError(f"norm_stats[{key!r}]: expected mean/std(/q01/q99), got {sorted(rec)}")

        def dec(name):
            v = rec.get(name)
            return None if v is None else np.asarray(v)

        out[key] = NormStats(dec("mean"), dec("std"), dec("q01"), dec("q99"))
    return out


def save(directory, norm_stats: Mapping[str, NormStats]) -> None:
    """normalize.py:133-137."""
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "norm_stats.json"), "w") as f:
        f.write(serialize_json(norm_stats))


def load(directory) -> dict:
    """normalize.py:140-145."""
    path = os.path.join(directory, "norm_stats.json")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Norm stats file not found at: {path}")
    with open(path) as f:
        return deserialize_json(f.read())


def load_norm_stats(assets_dir, asset_id: str) -> dict:
    """training/checkpoints.py:110-114."""
    return load(os.path.join(assets_dir, asset_id))


class RunningStats:
    """shared/normalize.py:17-116: streaming mean / std and histogram quantiles (5000 bins per dimension, re-binned when
    the observed range grows).  Same arithmetic in the same order, so results equal the reference's bit for bit."""

    BINS = 5000

    def __init__(self):
        self._n = 0
        self._mean = self._sq = self._lo = self._hi = None
        self._hist: list = []
        self._edges: list = []

    def update(self, batch: np.ndarray) -> None:
        rows = batch.reshape(-1, batch.shape[-1])
        k, dim = rows.shape
        if self._n == 0:
            self._mean, self._sq = rows.mean(axis=0), (rows**2).mean(axis=0)
            self._lo, self._hi = rows.min(axis=0), rows.max(axis=0)
            self._hist = [np.zeros(self.BINS) for _ in range(dim)]
            self._edges = [np.linspace(self._lo[i] - 1e-10, self._hi[i] + 1e-10, self.BINS + 1) for i in range(dim)]
        else:
            if dim != self._mean.size:
                raise ValueError("The length of new vectors does not match the initialized vector length.")
            hi, lo = rows.max(axis=0), rows.min(axis=0)
            grew = bool(np.any(hi > self._hi)) or bool(np.any(lo < self._lo))
            self._hi, self._lo = np.maximum(self._hi, hi), np.minimum(self._lo, lo)
            if grew:
                self._rebin()
        self._n += k
        w = k / self._n
        self._mean += (rows.mean(axis=0) - self._mean) * w
        self._sq += ((rows**2).mean(axis=0) - self._sq) * w
        for i in range(dim):
            self._hist[i] += np.histogram(rows[:, i], bins=self._edges[i])[0]

    def _rebin(self) -> None:
        for i, old in enumerate(self._edges):
            new = np.linspace(self._lo[i], self._hi[i], self.BINS + 1)
            self._hist[i] = np.histogram(old[:-1], bins=new, weights=self._hist[i])[0]
            self._edges[i] = new

    def get_statistics(self) -> NormStats:
        if self._n < 2:
            raise ValueError("Cannot compute statistics for less than 2 vectors.")
        std = np.sqrt(np.maximum(0, self._sq - self._mean**2))
        qs = []
        for q in (0.01, 0.99):
            target = q * self._n
            qs.append(np.array([e[np.searchsorted(np.cumsum(h), target)] for h, e in zip(self._hist, self._edges)]))
        return NormStats(mean=self._mean, std=std, q01=qs[0], q99=qs[1])


# ---------------------------------------------------------------------------------------------------------------
# transforms (host, numpy; unbatched dicts in, unbatched dicts out -- transforms.py:23-37)
# ---------------------------------------------------------------------------------------------------------------


@dataclasses.dataclass(frozen=True)
class Group:
    """transforms.py:39-60: `push` appends input transforms and PREPENDS output transforms."""

    inputs: Sequence[Callable] = ()
    outputs: Sequence[Callable] = ()

    def push(self, *, inputs: Sequence[Callable] = (), outputs: Sequence[Callable] = ()) -> "Group":
        return Group(inputs=(*self.inputs, *inputs), outputs=(*outputs, *self.outputs))


def compose(transforms: Sequence[Callable]) -> Callable:
    chain = tuple(transforms)

    def run(data):
        for t in chain:
            data = t(data)
        return data

    return run


def _tree_map_leaves(fn, tree):
    if isinstance(tree, Mapping):
        return {k: _tree_map_leaves(fn, v) for k, v in tree.items()}
    return fn(tree)


@dataclasses.dataclass(frozen=True)
class RepackTransform:
    """transforms.py:79-101: new nested structure whose leaves name '/'-paths of the incoming dict."""

    structure: Mapping

    def __call__(self, data):
        flat = flatten_dict(data)
        return _tree_map_leaves(lambda path: flat[path], self.structure)


@dataclasses.dataclass(frozen=True)
class InjectDefaultPrompt:
    """transforms.py:104-111."""

    prompt: str | None

    def __call__(self, data):
        if self.prompt is not None and "prompt" not in data:
            data["prompt"] = np.asarray(self.prompt)
        return data


@dataclasses.dataclass(frozen=True)
class InsertAdvantageIntoPrompt:
    """transforms.py:113-121 (AWBC, BASELINE.json configs[4]): the advantage label becomes part of the prompt text,
    `<prompt>, Advantage: <a with 4 decimals>`; first input transform when the data config asks for it
    (training/config.py:431-432)."""

    def __call__(self, data):
        if "advantage" not in data:
            raise AssertionError(f"advantage is not in data, data_keys: {data.keys()}")
        if "prompt" not in data:
            raise AssertionError(f"prompt is not in data, data_keys: {data.keys()}")
        data["prompt"] = data["prompt"] + f", Advantage: {data['advantage']:.4f}"
        return data


def _require_quantiles(norm_stats) -> None:
    for path, s in flatten_dict(norm_stats).items():
        if s.q01 is None or s.q99 is None:
            raise ValueError(
                f"quantile stats must be provided if use_quantile_norm is True. Key {path} is missing q01 or q99."
            )


@dataclasses.dataclass(frozen=True)
class Normalize:
    """transforms.py:124-155: z-score `(x - mean) / (std + 1e-6)` or quantile `(x - q01) / (q99 - q01 + 1e-6) * 2 - 1`;
    statistics are cut to the leaf's last dimension."""

    norm_stats: Mapping | None
    use_quantiles: bool = False
    strict: bool = False

    def __post_init__(self):
        if self.norm_stats is not None and self.use_quantiles:
            _require_quantiles(self.norm_stats)

    def __call__(self, data):
        if self.norm_stats is None:
            return data

        def z(x, s):
            n = x.shape[-1]
            return (x - s.mean[..., :n]) / (s.std[..., :n] + 1e-6)

        def q(x, s):
            n = x.shape[-1]
            lo, hi = s.q01[..., :n], s.q99[..., :n]
            return (x - lo) / (hi - lo + 1e-6) * 2.0 - 1.0

        return apply_tree(data, self.norm_stats, q if self.use_quantiles else z, strict=self.strict)


@dataclasses.dataclass(frozen=True)
class Unnormalize:
    """transforms.py:158-191: inverse of `Normalize`; every statistics key must be present in the data.  Dimensions of
    the leaf beyond the statistics' length are padded with mean 0 / std 1 (z-score) or passed through (quantile)."""

    norm_stats: Mapping | None
    use_quantiles: bool = False

    def __post_init__(self):
        if self.norm_stats is not None and self.use_quantiles:
            _require_quantiles(self.norm_stats)

    def __call__(self, data):
        if self.norm_stats is None:
            return data

        def z(x, s):
            n = x.shape[-1]
            return x * (pad_to_dim(s.std, n, value=1.0) + 1e-6) + pad_to_dim(s.mean, n, value=0.0)

        def q(x, s):
            lo, hi = s.q01, s.q99
            d = lo.shape[-1]
            if d < x.shape[-1]:
                return np.concatenate([(x[..., :d] + 1.0) / 2.0 * (hi - lo + 1e-6) + lo, x[..., d:]], axis=-1)
            return (x + 1.0) / 2.0 * (hi - lo + 1e-6) + lo

        return apply_tree(data, self.norm_stats, q if self.use_quantiles else z, strict=True)


def resize_with_pad(images: np.ndarray, height: int, width: int) -> np.ndarray:
    """packages/openpi-client/src/openpi_client/image_tools.py:15-58: aspect-preserving PIL bilinear resize of uint8
    [..., H, W, C] images, centred on a zero canvas (the host-side resize of the serving path; the engine has its own
    device-side resize for tensors, pi05_preprocess_image)."""
    from PIL import Image

    if images.shape[-3:-1] == (height, width):
        return images
    lead = images.shape[:-3]
    out = []
    for im in images.reshape(-1, *images.shape[-3:]):
        pil = Image.fromarray(im)
        w0, h0 = pil.size
        ratio = max(w0 / width, h0 / height)
        h1, w1 = int(h0 / ratio), int(w0 / ratio)
        canvas = Image.new(pil.mode, (width, height), 0)
        canvas.paste(pil.resize((w1, h1), resample=Image.BILINEAR), (max(0, int((width - w1) / 2)), max(0, int((height - h1) / 2))))
        out.append(np.asarray(canvas))
    out = np.stack(out)
    return out.reshape(*lead, *out.shape[-3:])


@dataclasses.dataclass(frozen=True)
class ResizeImages:
    """transforms.py:194-201."""

    height: int
    width: int

    def __call__(self, data):
        data["image"] = {k: resize_with_pad(v, self.height, self.width) for k, v in data["image"].items()}
        return data


def _shift_by_state(data, mask, sign):
    if "actions" not in data or mask is None:
        return data
    m = np.asarray(mask)
    d = m.shape[-1]
    offs = np.expand_dims(np.where(m, data["state"][..., :d], 0), axis=-2)
    actions = data["actions"]
    if sign > 0:
        actions[..., :d] += offs
    else:
        actions[..., :d] -= offs
    data["actions"] = actions
    return data


@dataclasses.dataclass(frozen=True)
class DeltaActions:
    """transforms.py:213-232: masked action dimensions become offsets from the current state (in place)."""

    mask: Sequence[bool] | None

    def __call__(self, data):
        return _shift_by_state(data, self.mask, -1)


@dataclasses.dataclass(frozen=True)
class AbsoluteActions:
    """transforms.py:235-254: the inverse (in place)."""

    mask: Sequence[bool] | None

    def __call__(self, data):
        return _shift_by_state(data, self.mask, +1)


@dataclasses.dataclass(frozen=True)
class PadStatesAndActions:
    """transforms.py:359-369."""

    model_action_dim: int

    def __call__(self, data):
        data["state"] = pad_to_dim(data["state"], self.model_action_dim, axis=-1)
        if "actions" in data:
            data["actions"] = pad_to_dim(data["actions"], self.model_action_dim, axis=-1)
        return data


class PaligemmaTokenizer:
    """models/tokenizer.py:13-47.  The reference downloads `paligemma_tokenizer.model` from a bucket; here the
    SentencePiece model is handed in (path or serialized proto) because a serving box has no egress.

    pi0.5 format (`state` given): the normalised state is cut into 256 bins over [-1, 1) and written into the prompt,
    `Task: <text>, State: <b0 b1 ...>;\\nAction: `; pi0 format: `<text>` + "\\n".  Output is padded with id 0 / mask
    False to `max_len` (or truncated)."""

    def __init__(self, max_len: int = 48, *, model_path: str | None = None, model_proto: bytes | None = None):
        import sentencepiece

        if (model_path is None) == (model_proto is None):
            raise ValueError("PaligemmaTokenizer needs exactly one of model_path / model_proto (paligemma_tokenizer.model)")
        if model_proto is None:
            with open(model_path, "rb") as f:
                model_proto = f.read()
        self._max_len = int(max_len)
        self._sp = sentencepiece.SentencePieceProcessor(model_proto=model_proto)

    def tokenize(self, prompt: str, state: np.ndarray | None = None):
        text = prompt.strip().replace("_", " ").replace("\n", " ")
        if state is not None:
            bins = np.digitize(state, bins=np.linspace(-1, 1, 256 + 1)[:-1]) - 1
            ids = self._sp.encode(f"Task: {text}, State: {' '.join(map(str, bins))};\nAction: ", add_bos=True)
        else:
            ids = self._sp.encode(text, add_bos=True) + self._sp.encode("\n")
        n = len(ids)
        if n < self._max_len:
            fill = [False] * (self._max_len - n)
            return np.asarray(ids + fill), np.asarray([True] * n + fill)
        return np.asarray(ids[: self._max_len]), np.asarray([True] * self._max_len)


@dataclasses.dataclass(frozen=True)
class TokenizePrompt:
    """transforms.py:279-298."""

    tokenizer: Any
    discrete_state_input: bool = False

    def __call__(self, data):
        prompt = data.pop("prompt", None)
        if prompt is None:
            raise ValueError("Prompt is required")
        state = None
        if self.discrete_state_input:
            state = data.get("state")
            if state is None:
                raise ValueError("State is required.")
        if not isinstance(prompt, str):
            prompt = prompt.item()
        tokens, mask = self.tokenizer.tokenize(prompt, state)
        return {**data, "tokenized_prompt": tokens, "tokenized_prompt_mask": mask}


@dataclasses.dataclass(frozen=True)
class AgilexInputs:
    """policies/agilex_policy.py:14-153: camera renaming to the model's keys, uint8 HWC images, state/actions padded to
    the model's action width with out-of-range joint values (|x| > pi) zeroed.  `pi05=False` adds the pi0 `action_mask`."""

    action_dim: int
    pi05: bool = True
    mask_state: bool = False
    zero_out_of_range_state: bool = True  # agilex_policy.py:92-94; the ARX variant (arx_policy.py) leaves the state alone

    REQUIRED = {"top_head": "base_0_rgb", "hand_left": "left_wrist_0_rgb", "hand_right": "right_wrist_0_rgb"}
    OPTIONAL = {"his_-100_top_head": "base_-100_rgb", "his_-100_hand_left": "left_wrist_-100_rgb",
                "his_-100_hand_right": "right_wrist_-100_rgb"}
    PASSTHROUGH = ("frame_index", "episode_length", "progress", "image_original", "episode_index")

    def __call__(self, data):
        cams = data["images"]
        known = {**self.REQUIRED, **self.OPTIONAL}
        if set(cams) - set(known):
            raise ValueError(f"Expected images to contain {tuple(self.REQUIRED)}, got {tuple(cams)}")
        state = pad_to_dim(data["state"], self.action_dim).squeeze()
        images, masks = {}, {}
        for cam, key in known.items():
            if cam not in cams:
                if cam in self.OPTIONAL:
                    continue
                raise ValueError(f"Camera {cam} not found in data")
            img = cams[cam]
            if isinstance(img, torch.Tensor):
                img = img.cpu().numpy()
            if np.issubdtype(img.dtype, np.floating):
                img = (255 * img).astype(np.uint8)
            if img.shape[0] == 3:
                img = np.transpose(img, (1, 2, 0))
            images[key], masks[key] = img, np.True_
        if self.zero_out_of_range_state:
            state = np.where(state > np.pi, 0, state)
            state = np.where(state < -np.pi, 0, state)
        out = {"image": images, "image_mask": masks, "state": np.zeros_like(state) if self.mask_state else state}
        if "actions" in data:
            actions = pad_to_dim(data["actions"], self.action_dim)
            actions = np.where(actions > np.pi, 0, actions)
            actions = np.where(actions < -np.pi, 0, actions)
            if not self.pi05:
                am = np.ones_like(actions, dtype=bool)
                am[:, self.action_dim:] = False
                out["action_mask"] = am
            out["actions"] = actions.squeeze()
        if "prompt" in data:
            out["prompt"] = data["prompt"]
        for k in self.PASSTHROUGH:
            if k in data:
                out[k] = data[k]
        for k in ("action_advantage", "action_advantage_original"):
            if k in data:
                v = data[k]
                if isinstance(v, np.ndarray):
                    out[k] = torch.from_numpy(v)
                elif isinstance(v, torch.Tensor):
                    out[k] = v.detach().clone()
                else:
                    raise NotImplementedError(f"Unsupported type: {type(v)}")
        return out


def ARXInputs(action_dim: int, pi05: bool = True, mask_state: bool = False) -> AgilexInputs:
    """policies/arx_policy.py:14-135: the Agilex transform without the out-of-range filter on the STATE (actions are
    still filtered); kai0's second robot (training/config.py:457-545)."""
    return AgilexInputs(action_dim=action_dim, pi05=pi05, mask_state=mask_state, zero_out_of_range_state=False)


@dataclasses.dataclass(frozen=True)
class AgilexOutputs:
    """agilex_policy.py:156-162: the 14 real joint dimensions of the chunk."""

    def __call__(self, data):
        return {"actions": np.asarray(data["actions"][:, :14])}


ARXOutputs = AgilexOutputs  # arx_policy.py:138-144: the same 14 dimensions


# ---------------------------------------------------------------------------------------------------------------
# Policy
# ---------------------------------------------------------------------------------------------------------------

_RTC_KEYS = ("prev_action_chunk", "inference_delay", "execute_horizon")  # policy.py:84-90


def _copy_structure(tree):
    """policy.py:70 (`jax.tree.map(lambda x: x, obs)`): new containers, shared leaves."""
    return _tree_map_leaves(lambda x: x, tree)


class _NullContext:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _Staging:
    """Pinned host blocks, one per field, reused across calls: a request batch is written into them with numpy and leaves
    for the device as ONE asynchronous copy per field (the reference issues a pageable, blocking copy per leaf)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.pinned = self.device.type == "cuda" and torch.cuda.is_available()
        self._blocks: dict = {}

    def put(self, name: str, arrays: Sequence[np.ndarray]) -> torch.Tensor:
        first = np.asarray(arrays[0])
        for a in arrays[1:]:
            if np.asarray(a).shape != first.shape:
                raise ValueError(f"{name}: requests of one batch must agree in shape, got {first.shape} and {np.asarray(a).shape}")
        dtype = torch.from_numpy(np.zeros((), dtype=first.dtype)).dtype
        shape = (len(arrays), *first.shape)
        blk = self._blocks.get(name)
        if blk is None or blk.dtype != dtype or tuple(blk.shape[1:]) != shape[1:] or blk.shape[0] < shape[0]:
            blk = torch.empty(shape, dtype=dtype, pin_memory=self.pinned)
            self._blocks[name] = blk
        host = blk[: shape[0]]
        view = host.numpy()
        for i, a in enumerate(arrays):
            view[i] = a
        if self.device.type == "cpu":
            return host.clone()
        return host.to(self.device, non_blocking=True)


class _Prepared:
    """One request after its input transforms: what `Policy.infer_prepared` batches."""

    __slots__ = ("raw", "data", "noise", "key")

    def __init__(self, raw, data, noise, key):
        self.raw, self.data, self.noise, self.key = raw, data, noise, key


class Policy:
    """`openpi.policies.policy.Policy` for the PyTorch pi0.5 engine (policy.py:23-129).

    infer(obs)           one request, the reference's call: returns {"state", "actions", "policy_timing"} after the
                         output transforms.
    infer_batch([obs])   any number of requests in one model call (not in the reference).  Requests carrying the
                         real-time-chunking keys (`prev_action_chunk`, `inference_delay`, `execute_horizon`) are grouped
                         by their scalar settings; each group is one `sample_actions` call.
    """

    def __init__(self, model, *, transforms: Sequence[Callable] = (), output_transforms: Sequence[Callable] = (),
                 sample_kwargs: dict | None = None, metadata: dict | None = None, pytorch_device: str = "cuda",
                 keep_uint8: bool = True, max_batch: int | None = None):
        self._model = model.to(pytorch_device)
        self._model.eval()
        # max_batch: the largest number of requests one model call may carry.  The engine plans its inference workspace
        # for that batch up front (instead of re-planning the first time a larger batch shows up) and longer request
        # lists are served in slices of this size.
        self._max_batch = int(max_batch) if max_batch else None
        if self._max_batch and hasattr(model, "_max_batch_hint"):
            model._max_batch_hint = max(int(model._max_batch_hint or 0), self._max_batch)
        self._input_transform = compose(transforms)
        self._output_transform = compose(output_transforms)
        self._sample_kwargs = dict(sample_kwargs or {})
        self._metadata = dict(metadata or {})
        self._device = pytorch_device
        self._keep_uint8 = keep_uint8
        self._staging = _Staging(pytorch_device)
        self._lock = threading.Lock()  # the engine is single-stream, one model call at a time (SURVEY §8b threading)

    @property
    def metadata(self) -> dict:
        return self._metadata

    def infer(self, obs: dict, *, noise: np.ndarray | None = None) -> dict:
        return self.infer_batch([obs], noise=None if noise is None else [noise])[0]

    def infer_batch(self, observations: Sequence[dict], *, noise: Sequence[np.ndarray | None] | None = None) -> list:
        if noise is not None and len(noise) != len(observations):
            raise ValueError("noise must hold one entry (or None) per observation")
        return self.infer_prepared([self.prepare(o, None if noise is None else noise[i]) for i, o in enumerate(observations)])

    def prepare(self, obs: dict, noise: np.ndarray | None = None) -> "_Prepared":
        """The host half of a request that needs no device and no lock: copy + input transforms (policy.py:70-71).  Safe to
        call from any thread; `RequestBatcher` runs it in the client's thread so that the transforms of the next batch overlap
        the model call of the current one."""
        data = self._input_transform(_copy_structure(obs))
        key = tuple((k, None if k not in obs or k == "prev_action_chunk" or obs[k] is None else int(obs[k])) for k in _RTC_KEYS)
        key += ("prev_action_chunk" in obs, noise is not None)
        return _Prepared(obs, data, noise, key)

    def infer_prepared(self, items: Sequence["_Prepared"]) -> list:
        """The device half: requests that can share a `sample_actions` call (same RTC scalars, noise given or not) are
        staged, copied, run and read back together, in slices of at most `max_batch`."""
        if len(items) == 0:
            return []
        groups: dict = {}
        for i, it in enumerate(items):
            groups.setdefault(it.key, []).append(i)
        results: list = [None] * len(items)
        # a server thread that never touched CUDA starts on device 0: make the engine's device current for the call
        dev = torch.device(self._device)
        on_dev = torch.cuda.device(dev) if dev.type == "cuda" and dev.index is not None else _NullContext()
        with self._lock, on_dev:
            for members in groups.values():
                step = self._max_batch or len(members)
                for lo in range(0, len(members), step):
                    idx = members[lo:lo + step]
                    has_noise = items[idx[0]].noise is not None
                    outs = self._run_group([items[i].raw for i in idx], [items[i].data for i in idx],
                                           [items[i].noise for i in idx] if has_noise else None)
                    for i, o in zip(idx, outs):
                        results[i] = o
        return results

    def _run_group(self, raw, prepared, noise):
        from .model import Observation

        B = len(prepared)
        st = self._staging
        first = prepared[0]
        batch = {
            "image": {k: st.put(f"image/{k}", [p["image"][k] for p in prepared]) for k in first["image"]},
            "image_mask": {k: st.put(f"image_mask/{k}", [np.asarray(p["image_mask"][k]) for p in prepared])
                           for k in first["image_mask"]},
            "state": st.put("state", [np.asarray(p["state"]) for p in prepared]),
        }
        for k in ("tokenized_prompt", "tokenized_prompt_mask", "token_ar_mask", "token_loss_mask"):
            if k in first:
                batch[k] = st.put(k, [np.asarray(p[k]) for p in prepared])
        kwargs = dict(self._sample_kwargs)
        if "prev_action_chunk" in raw[0]:
            kwargs["prev_action_chunk"] = st.put("prev_action_chunk",
                                                 [np.asarray(r["prev_action_chunk"], dtype=np.float32) for r in raw])
        for k in ("inference_delay", "execute_horizon"):
            if k in raw[0]:
                kwargs[k] = raw[0][k]
        if noise is not None and noise[0] is not None:
            ns = [np.asarray(n, dtype=np.float32) for n in noise]
            ns = [n[0] if n.ndim == 3 else n for n in ns]  # policy.py:100-101 accepts [H, A] or [1, H, A]
            kwargs["noise"] = st.put("noise", ns)
        observation = Observation.from_dict(batch, keep_uint8=self._keep_uint8)
        t0 = time.monotonic()
        actions = self._model.sample_actions(self._device, observation, **kwargs)
        actions = actions.detach().to("cpu").numpy()  # one read-back for the whole batch; synchronises
        ms = (time.monotonic() - t0) * 1e3
        out = []
        for i, p in enumerate(prepared):
            reply = self._output_transform({"state": np.asarray(p["state"]), "actions": actions[i]})
            reply["policy_timing"] = {"infer_ms": ms, "batch": B}
            out.append(reply)
        return out


class RequestBatcher:
    """Drop-in `infer(obs)` for a server with many clients: requests arriving within `max_wait_ms` of each other are
    served by one model call (at most `max_batch` per call).  A lone request waits at most `max_wait_ms`.  The input
    transforms of a request run in the thread that submits it (`Policy.prepare`), the worker thread only stages, runs
    and unpacks batches (`Policy.infer_prepared`)."""

    _STOP = object()

    def __init__(self, policy: Policy, *, max_batch: int = 8, max_wait_ms: float = 2.0):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        if getattr(policy, "_max_batch", None):
            max_batch = min(max_batch, policy._max_batch)
        self._policy = policy
        self._max_batch = int(max_batch)
        self._max_wait = float(max_wait_ms) / 1e3
        self._q: queue.Queue = queue.Queue()
        self._closed = False
        self.batches_served = 0
        self.requests_served = 0
        self._worker = threading.Thread(target=self._loop, name="kai0-request-batcher", daemon=True)
        self._worker.start()

    @property
    def metadata(self) -> dict:
        return self._policy.metadata

    def submit(self, obs: dict) -> Future:
        if self._closed:
            raise RuntimeError("RequestBatcher is closed")
        fut: Future = Future()
        try:
            item = self._policy.prepare(obs)  # in the CLIENT'S thread: transforms of many clients run side by side, and
        except Exception as e:  # noqa: BLE001  overlap the model call in flight; a bad request fails here, alone
            fut.set_exception(e)
            return fut
        self._q.put((item, fut))
        return fut

    def infer(self, obs: dict) -> dict:
        return self.submit(obs).result()

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(self._STOP)
            self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _loop(self):
        while True:
            item = self._q.get()
            if item is self._STOP:
                break
            pending = [item]
            deadline = time.monotonic() + self._max_wait
            stop = False
            while len(pending) < self._max_batch:
                left = deadline - time.monotonic()
                try:
                    nxt = self._q.get(timeout=left) if left > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is self._STOP:
                    stop = True
                    break
                pending.append(nxt)
            self._serve(pending)
            if stop:
                break
        # fail whatever raced with close()
        while True:
            try:
                item = self._q.get_nowait()
            except queue.Empty:
                return
            if item is not self._STOP:
                item[1].set_exception(RuntimeError("RequestBatcher is closed"))

    def _serve(self, pending):
        live = [(it, f) for it, f in pending if f.set_running_or_notify_cancel()]
        if not live:
            return
        try:
            replies = self._policy.infer_prepared([it for it, _ in live])
        except Exception:  # noqa: BLE001  -- one bad request must not take its neighbours down: retry one by one
            for it, f in live:
                try:
                    f.set_result(self._policy.infer_prepared([it])[0])
                except Exception as e:  # noqa: BLE001
                    f.set_exception(e)
        else:
            for (_, f), r in zip(live, replies):
                f.set_result(r)
        self.batches_served += 1
        self.requests_served += len(live)


# ---------------------------------------------------------------------------------------------------------------
# wire format of the policy server: msgpack with numpy framing, and the per-connection message loop
# ---------------------------------------------------------------------------------------------------------------


def pack_array(obj):
    """packages/openpi-client/src/openpi_client/msgpack_numpy.py:21-41: an ndarray travels as
    {__ndarray__: True, data: raw bytes, dtype: dtype.str, shape}, a numpy scalar as {__npgeneric__: True, data, dtype};
    void / object / complex dtypes are refused."""
    if isinstance(obj, (np.ndarray, np.generic)) and obj.dtype.kind in ("V", "O", "c"):
        raise ValueError(f"Unsupported dtype: {obj.dtype}")
    if isinstance(obj, np.ndarray):
        return {b"__ndarray__": True, b"data": obj.tobytes(), b"dtype": obj.dtype.str, b"shape": obj.shape}
    if isinstance(obj, np.generic):
        return {b"__npgeneric__": True, b"data": obj.item(), b"dtype": obj.dtype.str}
    return obj


def unpack_array(obj):
    """msgpack_numpy.py:44-51."""
    if b"__ndarray__" in obj:
        return np.ndarray(buffer=obj[b"data"], dtype=np.dtype(obj[b"dtype"]), shape=obj[b"shape"])
    if b"__npgeneric__" in obj:
        return np.dtype(obj[b"dtype"]).type(obj[b"data"])
    return obj


def packb(obj) -> bytes:
    import msgpack

    return msgpack.packb(obj, default=pack_array)


def unpackb(data: bytes):
    import msgpack

    return msgpack.unpackb(data, object_hook=unpack_array)


class MessageHandler:
    """The per-connection protocol of `WebsocketPolicyServer._handler` (src/openpi/serving/websocket_policy_server.py:
    48-83) without the socket: `greeting()` is the first frame a client receives (the policy metadata); `handle(frame)`
    turns one request frame into one reply frame -- unpack, `policy.infer`, attach `server_timing` (`infer_ms`, and
    `prev_total_ms` of the previous exchange), pack.  A failing request yields the traceback as a TEXT frame and marks the
    session closed, as the reference sends it before closing with INTERNAL_ERROR.  `policy` is anything with
    `infer(obs)` -- a `Policy` or a `RequestBatcher` shared by all connections.  Put it under any transport, e.g.

        async def handler(ws):                                   # websockets.asyncio.server, as the reference uses
            h = MessageHandler(policy, metadata)
            await ws.send(h.greeting())
            async for frame in ws:
                reply = await asyncio.get_running_loop().run_in_executor(None, h.handle, frame)
                await ws.send(reply)
                if h.closed: await ws.close(code=1011, reason="Internal server error. Traceback included in previous frame."); break
    """

    def __init__(self, policy, metadata: dict | None = None):
        self._policy = policy
        self._metadata = metadata or {}
        self._prev_total = None
        self.closed = False

    def greeting(self) -> bytes:
        return packb(self._metadata)

    def handle(self, frame: bytes):
        import traceback

        start = time.monotonic()
        try:
            obs = unpackb(frame)
            t0 = time.monotonic()
            action = self._policy.infer(obs)
            timing = {"infer_ms": (time.monotonic() - t0) * 1000}
            if self._prev_total is not None:
                timing["prev_total_ms"] = self._prev_total * 1000
            action["server_timing"] = timing
            reply = packb(action)
            self._prev_total = time.monotonic() - start
            return reply
        except Exception:  # noqa: BLE001
            self.closed = True
            return traceback.format_exc()


# ---------------------------------------------------------------------------------------------------------------
# policy construction from a checkpoint directory
# ---------------------------------------------------------------------------------------------------------------


def agilex_pi05_transforms(*, action_dim: int, max_token_len: int, tokenizer, norm_stats, default_prompt: str | None = None,
                           use_delta_joint_actions: bool = True, image_size: int = 224, mask_state: bool = False,
                           use_quantile_norm: bool = True, repack: Group | None = None,
                           insert_advantage_into_prompt: bool = False):
    """The transform chain `create_trained_policy` assembles for a pi0.5 Agilex config (policy_config.py:75-90 over
    training/config.py:129-141,420-452): returns (input transforms, output transforms)."""
    repack = repack or Group()
    data = Group(inputs=[AgilexInputs(action_dim=action_dim, pi05=True, mask_state=mask_state)], outputs=[AgilexOutputs()])
    if insert_advantage_into_prompt:  # training/config.py:431-432 (AWBC prompts)
        data = Group(inputs=[InsertAdvantageIntoPrompt(), *data.inputs], outputs=data.outputs)
    if use_delta_joint_actions:
        m = make_bool_mask(6, -1, 6, -1)  # joints relative to the state, the two grippers absolute
        data = data.push(inputs=[DeltaActions(m)], outputs=[AbsoluteActions(m)])
    model_in = [InjectDefaultPrompt(default_prompt), ResizeImages(image_size, image_size),
                TokenizePrompt(tokenizer, discrete_state_input=True), PadStatesAndActions(action_dim)]
    ins = [*repack.inputs, InjectDefaultPrompt(default_prompt), *data.inputs,
           Normalize(norm_stats, use_quantiles=use_quantile_norm), *model_in]
    outs = [Unnormalize(norm_stats, use_quantiles=use_quantile_norm), *data.outputs, *repack.outputs]
    return ins, outs


def create_trained_policy(model, checkpoint_dir, *, asset_id: str | None, tokenizer, default_prompt: str | None = None,
                          norm_stats: Mapping | None = None, sample_kwargs: dict | None = None,
                          pytorch_device: str | None = None, metadata: dict | None = None,
                          use_delta_joint_actions: bool = True, repack_transforms: Group | None = None) -> Policy:
    """policy_config.py:16-94 for a PyTorch pi0.5 checkpoint directory (`model.safetensors` +
    `assets/<asset_id>/norm_stats.json`, as train_pytorch.py:149-189 writes it).  `model` is a constructed `PI0Pytorch`
    (the reference builds it from its TrainConfig, which needs the jax-side config registry); the weights are loaded with
    `safetensors.torch.load_model` and cast with `to_bfloat16_for_selected_params` exactly as :53-55."""
    import safetensors.torch

    weights = os.path.join(str(checkpoint_dir), "model.safetensors")
    if not os.path.exists(weights):
        raise FileNotFoundError(f"No PyTorch checkpoint (model.safetensors) in {checkpoint_dir}")
    safetensors.torch.load_model(model, weights)
    model.paligemma_with_expert.to_bfloat16_for_selected_params("bfloat16")
    if norm_stats is None:
        if asset_id is None:
            raise ValueError("Asset id is required to load norm stats.")
        norm_stats = load_norm_stats(os.path.join(str(checkpoint_dir), "assets"), asset_id)
    if pytorch_device is None:
        pytorch_device = "cuda" if torch.cuda.is_available() else "cpu"
    cfg = model.ecfg
    ins, outs = agilex_pi05_transforms(action_dim=cfg.action_dim, max_token_len=cfg.max_token_len, tokenizer=tokenizer,
                                       norm_stats=norm_stats, default_prompt=default_prompt,
                                       use_delta_joint_actions=use_delta_joint_actions, image_size=cfg.image_size,
                                       repack=repack_transforms)
    return Policy(model, transforms=ins, output_transforms=outs, sample_kwargs=sample_kwargs, metadata=metadata,
                  pytorch_device=pytorch_device)
