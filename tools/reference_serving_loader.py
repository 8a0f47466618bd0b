"""Execute the REFERENCE'S OWN serving-side host code in place (test infrastructure; needs /root/reference).

Loads, from where they lie, `openpi/shared/normalize.py`, `openpi/transforms.py`, `openpi/models/tokenizer.py`,
`openpi/policies/agilex_policy.py` and `packages/openpi-client/src/openpi_client/image_tools.py`, so that
tests/test_serving_cpu.py and tools/make_golden_serving.py can pin kai0_b200/serving.py to the reference itself.
Nothing is copied.  Third-party modules those files import that this image lacks are replaced by the smallest stand-in
that reproduces the behaviour the files use (pinned versions from the reference's uv.lock):

  flax 0.10.2  `traverse_util.flatten_dict(tree, sep=)/unflatten_dict(flat, sep=)`: depth-first walk over nested dicts,
               keys joined with `sep`, empty sub-dicts dropped (keep_empty_nodes=False); inverse splits on `sep`.
  jax          only `jax.tree.map(fn, tree)` over nested dicts of leaves.
  numpydantic 1.6.9  `NDArray` as a pydantic field type: validates to `np.ndarray`, JSON-serialises as nested lists.
  orbax, transformers.AutoProcessor, fsq_tokenizer: imported by tokenizer.py for classes this path never touches.
  openpi.shared.download.maybe_download: returns the SentencePiece model path registered with `set_tokenizer_model`
               (the reference fetches gs://big_vision/paligemma_tokenizer.model; no egress here).
  openpi.models.model: only the `ModelType` enum (models/model.py:30-37) for `load()`; `load_policy()` executes the real
               models/model.py (for `Observation.from_dict`, :122-157) and policies/policy.py (`Policy.infer`, :68-124)
               in place, with augmax / flax.nnx / flax.struct / orbax / nnx_utils / the PyTorch model module stubbed
               (`struct.dataclass` -> `dataclasses.dataclass`; only import-time names are needed, the JAX branches of
               `Policy` are never taken with `is_pytorch=True`).
"""
from __future__ import annotations

import enum
import importlib.util
import os
import pathlib
import sys
import types
import typing

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STAGED = os.path.join(_ROOT, "baseline", "_ref")


def _first(*candidates):
    """The read-only checkout of the build container, else the byte-for-byte staged copy under baseline/_ref
    (tools/stage_reference.py; it travels to the GPU box, where /root/reference does not exist)."""
    if os.environ.get("KAI0_REFERENCE_STAGED_ONLY"):  # exercise the GPU-box situation in the build container
        candidates = candidates[1:]
    return next((c for c in candidates if os.path.exists(c)), candidates[0])


REPO = "/root/reference"
SRC = _first(os.path.join(REPO, "src", "openpi"), os.path.join(_STAGED, "openpi"))
CLIENT = _first(os.path.join(REPO, "packages", "openpi-client", "src", "openpi_client"), os.path.join(_STAGED, "openpi_client"))
TRAIN_SCRIPT = _first(os.path.join(REPO, "scripts", "train_pytorch.py"), os.path.join(_STAGED, "scripts", "train_pytorch.py"))
_loaded = None
_tokenizer_model: dict = {"path": None}


def available() -> bool:
    return os.path.isfile(os.path.join(SRC, "transforms.py")) and os.path.isfile(os.path.join(CLIENT, "image_tools.py"))


def set_tokenizer_model(path: str) -> None:
    _tokenizer_model["path"] = path


def _exec_as(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _module(name: str, **attrs):
    """Create the module unless a real/stub one is already registered; attributes are added either way when missing."""
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        if not hasattr(m, k):
            setattr(m, k, v)
    return m


def _flatten(tree, sep="/", **_):
    out = {}

    def walk(node, prefix):
        for k, v in node.items():
            p = (*prefix, k)
            if isinstance(v, dict):
                walk(v, p)
            else:
                out[sep.join(p) if sep is not None else p] = v

    walk(tree, ())
    return out


def _unflatten(flat, sep="/"):
    root = {}
    for path, v in flat.items():
        parts = path.split(sep) if sep is not None else path
        node = root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return root


def _tree_map(fn, tree, *rest):
    if isinstance(tree, dict):
        return {k: _tree_map(fn, v, *(r[k] for r in rest)) for k, v in tree.items()}
    return fn(tree, *rest)


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __getitem__(self, item):
        return self

    def __or__(self, other):
        return self

    def __ror__(self, other):
        return self

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything()


def numpydantic_stub():
    """numpydantic 1.6.9's `NDArray` as normalize.py uses it: a pydantic field type that validates to `np.ndarray` and
    JSON-serialises as nested lists."""
    import pydantic

    m = types.ModuleType("numpydantic")
    m.NDArray = typing.Annotated[
        typing.Any,
        pydantic.BeforeValidator(lambda v: v if v is None else np.asarray(v)),
        pydantic.PlainSerializer(lambda a: None if a is None else np.asarray(a).tolist(), when_used="json"),
    ]
    return m


class ModelType(enum.Enum):  # models/model.py:30-37 (values only)
    PI0 = "pi0"
    PI0_FAST = "pi0_fast"
    PI05 = "pi05"
    PI0_RTC = "pi0_rtc"
    PI05_RTC = "pi05_rtc"


_policy_loaded = None
_train_loaded = None
_policy_config_loaded = None


def load_policy_config():
    """`openpi/policies/policy_config.py` executed in place on top of `load_policy()`: the reference's own
    `create_trained_policy` (:16-94).  Supplied: `openpi.shared.download.maybe_download` (returns the local path),
    `openpi.training.checkpoints.load_norm_stats` (training/checkpoints.py:110-114: `normalize.load(assets_dir / asset_id)`,
    here through the reference's own normalize module; the real file imports orbax / jax at its top) and an empty
    `openpi.training.config` (type annotations only).  Returns (policy_config module, policy module, model module)."""
    global _policy_config_loaded
    if _policy_config_loaded is not None:
        return _policy_config_loaded
    policy, model = load_policy()
    R = load()

    class _StubModule(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything()

    names = ("jax", "jax.numpy", "openpi.shared.download", "openpi.training", "openpi.training.checkpoints",
             "openpi.training.config", "openpi.models.model", "openpi.policies.policy")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        jnp = _StubModule("jax.numpy")
        jx = _StubModule("jax")
        jx.numpy = jnp
        dl = _StubModule("openpi.shared.download")
        dl.maybe_download = lambda url, **kw: pathlib.Path(str(url))
        ck = _StubModule("openpi.training.checkpoints")
        ck.load_norm_stats = lambda assets_dir, asset_id: R.normalize.load(pathlib.Path(assets_dir) / asset_id)
        tr = _StubModule("openpi.training")
        tr.__path__ = []
        tr.checkpoints, tr.config = ck, _StubModule("openpi.training.config")
        sys.modules.update({"jax": jx, "jax.numpy": jnp, "openpi.shared.download": dl, "openpi.training": tr,
                            "openpi.training.checkpoints": ck, "openpi.training.config": tr.config,
                            "openpi.models.model": model, "openpi.policies.policy": policy})
        sys.modules["openpi.shared"].download = dl
        sys.modules["openpi"].training = tr
        sys.modules["openpi.models"].model = model
        sys.modules["openpi.policies"].policy = policy
        spec = importlib.util.spec_from_file_location("_kai0_reference_policy_config", os.path.join(SRC, "policies", "policy_config.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _policy_config_loaded = (mod, policy, model)
    return _policy_config_loaded


def load_train_script():
    """`scripts/train_pytorch.py` of the reference executed in place as a module (its `save_checkpoint` :149-189,
    `load_checkpoint` :192-273, `get_latest_checkpoint_step` :276-283 are what kai0_b200/checkpoint.py mirrors).  jax,
    wandb and the openpi config / data-loader / model modules it imports at the top are stubbed: none of them is touched
    by those three functions, which use torch, safetensors and the reference's own `openpi.shared.normalize` (real, loaded
    in place by `load()`)."""
    global _train_loaded
    if _train_loaded is not None:
        return _train_loaded
    load()

    class _StubModule(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything()

    names = ("jax", "wandb", "openpi.models.pi0_config", "openpi.models_pytorch", "openpi.models_pytorch.pi0_pytorch",
             "openpi.training", "openpi.training.config", "openpi.training.data_loader")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in names:
            m = _StubModule(k)
            m.__path__ = []
            sys.modules[k] = m
        op = sys.modules["openpi"]
        op.models_pytorch, op.training = sys.modules["openpi.models_pytorch"], sys.modules["openpi.training"]
        sys.modules["openpi.models"].pi0_config = sys.modules["openpi.models.pi0_config"]
        sys.modules["openpi.models_pytorch"].pi0_pytorch = sys.modules["openpi.models_pytorch.pi0_pytorch"]
        sys.modules["openpi.training"].config = sys.modules["openpi.training.config"]
        sys.modules["openpi.training"].data_loader = sys.modules["openpi.training.data_loader"]
        spec = importlib.util.spec_from_file_location("_kai0_reference_train_pytorch", TRAIN_SCRIPT)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _train_loaded = mod
    return mod


def load_policy():
    """Returns (policy module, model module) of the reference: `openpi/policies/policy.py` and `openpi/models/model.py`
    executed in place on top of `load()`."""
    global _policy_loaded
    if _policy_loaded is not None:
        return _policy_loaded
    import dataclasses

    R = load()

    class _StubModule(types.ModuleType):
        def __getattr__(self, name):  # any other attribute (type names in annotations, unused helpers)
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything()

    def stub(name, **attrs):
        m = _StubModule(name)
        m.__dict__.update(attrs)
        return m

    class _NnxModule:
        pass

    saved = {k: sys.modules.get(k) for k in ("augmax", "flax.nnx", "flax.struct", "jax", "jax.numpy", "openpi.models_pytorch",
                                             "openpi.models_pytorch.pi0_pytorch", "openpi.shared.image_tools",
                                             "openpi.shared.nnx_utils", "openpi.models.model", "openpi_client.base_policy",
                                             "orbax.checkpoint")}
    flax = sys.modules["flax"]
    nnx = stub("flax.nnx", Module=_NnxModule)
    struct = stub("flax.struct", dataclass=dataclasses.dataclass)
    jnp = stub("jax.numpy")
    # for these two files `jax` is: tree.map (used), plus names that only appear in annotations / JAX-only branches
    jx = stub("jax", numpy=jnp, tree=types.SimpleNamespace(map=_tree_map))
    try:
        sys.modules.update({"augmax": stub("augmax"), "flax.nnx": nnx, "flax.struct": struct, "jax": jx, "jax.numpy": jnp,
                            "orbax.checkpoint": stub("orbax.checkpoint")})
        flax.nnx, flax.struct = nnx, struct
        sys.modules["orbax"].checkpoint = sys.modules["orbax.checkpoint"]
        mp = sys.modules.get("openpi.models_pytorch") or stub("openpi.models_pytorch")
        mp.__path__ = getattr(mp, "__path__", [])
        sys.modules["openpi.models_pytorch"] = mp
        if "openpi.models_pytorch.pi0_pytorch" not in sys.modules:
            sys.modules["openpi.models_pytorch.pi0_pytorch"] = stub("openpi.models_pytorch.pi0_pytorch")
        mp.pi0_pytorch = sys.modules["openpi.models_pytorch.pi0_pytorch"]
        sh = sys.modules["openpi.shared"]
        if "openpi.shared.image_tools" not in sys.modules:
            sys.modules["openpi.shared.image_tools"] = stub("openpi.shared.image_tools")
        sh.image_tools = sys.modules["openpi.shared.image_tools"]
        sys.modules["openpi.shared.nnx_utils"] = stub("openpi.shared.nnx_utils")
        sh.nnx_utils = sys.modules["openpi.shared.nnx_utils"]
        at = sys.modules["openpi.shared.array_typing"]
        for name in ("Float", "Bool", "Int", "Real", "UInt8", "Array", "PyTree", "KeyArrayLike", "Params"):
            if not hasattr(at, name):
                setattr(at, name, _Anything())
        model = _exec_as("openpi.models.model", os.path.join(SRC, "models", "model.py"))
        sys.modules["openpi.models"].model = model
        oc = sys.modules["openpi_client"]
        oc.base_policy = _exec_as("openpi_client.base_policy", os.path.join(CLIENT, "base_policy.py"))
        policy = _exec_as("openpi.policies.policy", os.path.join(SRC, "policies", "policy.py"))
        sys.modules["openpi.policies"].policy = policy
    finally:
        for k, v in saved.items():  # leave the stubs of other loaders (tools/reference_loader.py) as they were
            if v is not None:
                sys.modules[k] = v
    _policy_loaded = (policy, model)
    return _policy_loaded


def load():
    """Returns a namespace with the reference's modules: .normalize, .transforms, .tokenizer, .agilex_policy,
    .image_tools, .ModelType."""
    global _loaded
    if _loaded is not None:
        return _loaded
    import pydantic

    # third-party stand-ins
    nd = _module("numpydantic")
    if not hasattr(nd, "NDArray"):
        nd.NDArray = numpydantic_stub().NDArray
    tu = _module("flax.traverse_util", flatten_dict=_flatten, unflatten_dict=_unflatten)
    fl = _module("flax", traverse_util=tu)
    fl.__path__ = getattr(fl, "__path__", [])
    jx = _module("jax")
    if not hasattr(jx, "tree"):
        jx.tree = types.SimpleNamespace(map=_tree_map)
    _module("orbax").__path__ = []
    _module("orbax.checkpoint")
    sys.modules["orbax"].checkpoint = sys.modules["orbax.checkpoint"]
    oc = _module("openpi_client")
    oc.__path__ = getattr(oc, "__path__", [])
    oc.image_tools = _exec_as("openpi_client.image_tools", os.path.join(CLIENT, "image_tools.py"))

    # openpi package skeleton (tools/reference_loader.py may have created parts of it already)
    op = _module("openpi")
    op.__path__ = getattr(op, "__path__", [])
    sh = _module("openpi.shared")
    sh.__path__ = getattr(sh, "__path__", [])
    op.shared = sh
    at = _module("openpi.shared.array_typing", typecheck=lambda f: f, PyTree=_Anything(), UInt8=_Anything(),
                 Float=_Anything(), Array=_Anything())
    if not hasattr(at, "PyTree"):
        at.PyTree = _Anything()
    sh.array_typing = at
    dl = _module("openpi.shared.download")
    dl.maybe_download = lambda url, **kw: pathlib.Path(_tokenizer_model["path"]) if str(url).endswith(
        "paligemma_tokenizer.model") else pathlib.Path(str(url))
    sh.download = dl
    sh.normalize = _exec_as("openpi.shared.normalize", os.path.join(SRC, "shared", "normalize.py"))
    md = _module("openpi.models")
    md.__path__ = getattr(md, "__path__", [])
    op.models = md
    ut = _module("openpi.models.utils")
    ut.__path__ = getattr(ut, "__path__", [])
    md.utils = ut
    ut.fsq_tokenizer = _module("openpi.models.utils.fsq_tokenizer")
    mm = _module("openpi.models.model", ModelType=ModelType)
    md.model = mm
    md.tokenizer = _exec_as("openpi.models.tokenizer", os.path.join(SRC, "models", "tokenizer.py"))
    op.transforms = _exec_as("openpi.transforms", os.path.join(SRC, "transforms.py"))
    po = _module("openpi.policies")
    po.__path__ = getattr(po, "__path__", [])
    op.policies = po
    po.agilex_policy = _exec_as("openpi.policies.agilex_policy", os.path.join(SRC, "policies", "agilex_policy.py"))
    _loaded = types.SimpleNamespace(normalize=sh.normalize, transforms=op.transforms, tokenizer=md.tokenizer,
                                    agilex_policy=po.agilex_policy, image_tools=oc.image_tools,
                                    ModelType=getattr(mm, "ModelType", ModelType))
    return _loaded
