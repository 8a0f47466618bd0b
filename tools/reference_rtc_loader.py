"""Execute the REFERENCE'S OWN real-time-chunking sampler in place (test infrastructure; needs /root/reference).

The reference implements RTC only in its JAX model: `src/openpi/models/pi0_rtc.py` (`get_prefix_weights` :47-61,
`make_attn_mask` :19-44, `posemb_sincos` :64-80, `Pi0RTC.embed_prefix` :121-154, `Pi0RTC.embed_suffix` :156-199,
`Pi0RTC.sample_actions` :234-360).  jax and flax are not installed here, so that file cannot run as it is.  What this
loader does instead:

  * the file is executed UNMODIFIED from where it lies, with `jax` / `jax.numpy` replaced by a small stand-in that maps
    the array functions the file uses onto torch (same semantics: weak-typed python scalars become float32, `jnp.minimum`
    propagates NaN, `x.at[idx].set(v)` is a functional update, `jax.vjp` is torch.autograd, `jax.lax.scan` a loop);
    flax / the openpi model modules are stubbed (only `Pi0RTC.__init__`, which builds the flax networks, needs them, and
    it is never called);
  * a `Pi0RTC` instance is created without `__init__` and given the four sub-networks its methods call
    (`PaliGemma.img`, `PaliGemma.llm`, `action_in_proj`, `time_mlp_in/out`, `action_out_proj`) as thin adapters over the
    PyTorch-path network of oracle/pi05_oracle.py -- which is itself pinned to the reference's PyTorch model.

Everything else -- prefix/suffix embedding order, masks, positions, the KV-cache protocol, the Euler scan, and the whole
guidance computation (prefix weights, delay masking, error, vector-Jacobian product, guidance weight, clipping, NaN
handling) -- is the reference's own code running.  tests/test_rtc_oracle_cpu.py compares oracle/rtc_oracle.py (and the
engine's host-side tables) with it; tools/make_golden_rtc.py commits its outputs for boxes without the checkout.
"""
from __future__ import annotations

import contextlib
import importlib.util
import math
import os
import sys
import types

import torch
import torch.nn.functional as F

PATH = "/root/reference/src/openpi/models/pi0_rtc.py"
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_loaded = None


def available() -> bool:
    return os.path.isfile(PATH)


# ------------------------------------------------------------------------------------------------------------
# jax.numpy on torch
# ------------------------------------------------------------------------------------------------------------
class JT(torch.Tensor):
    """torch.Tensor with the two jax.Array members the file uses that torch lacks."""

    def astype(self, dtype):
        return self.to(_dtype(dtype))

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, t):
        self.t = t

    def __getitem__(self, idx):
        return _AtIdx(self.t, idx)


class _AtIdx:
    def __init__(self, t, idx):
        self.t, self.idx = t, idx

    def set(self, value):
        out = self.t.clone()
        out[self.idx] = value
        return out


def _dtype(d):
    if d is bool:
        return torch.bool
    if d is float:
        return torch.float32
    if d is int:
        return torch.int64
    return d


def _w(x, dtype=None):
    """jnp.asarray: python floats are weakly typed float32, ints int (values only matter here)."""
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
    if dtype is not None:
        t = t.to(_dtype(dtype))
    return t.as_subclass(JT)


def _shape(s):
    return (s,) if isinstance(s, int) else tuple(s)


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    m.float32, m.bool_, m.int32, m.e, m.pi = torch.float32, torch.bool, torch.int64, math.e, math.pi
    m.asarray = m.array = _w
    m.broadcast_to = lambda x, shape: torch.broadcast_to(_w(x), _shape(shape))
    m.cumsum = lambda x, axis=None: torch.cumsum(_w(x), dim=axis)
    m.sum = lambda x, axis=None: torch.sum(_w(x)) if axis is None else torch.sum(_w(x), dim=axis)
    m.mean = lambda x, axis=None: torch.mean(_w(x)) if axis is None else torch.mean(_w(x), dim=axis)
    m.square = lambda x: torch.square(_w(x))
    m.logical_and = lambda a, b: torch.logical_and(_w(a), _w(b))
    m.minimum = lambda a, b: torch.minimum(*torch.broadcast_tensors(*_promote(_w(a), _w(b))))
    m.clip = lambda x, lo=None, hi=None: torch.clamp(_w(x), lo, hi)
    m.ones = lambda shape, dtype=None: _w(torch.ones(_shape(shape), dtype=_dtype(dtype) or torch.float32))
    m.zeros = lambda shape, dtype=None: _w(torch.zeros(_shape(shape), dtype=_dtype(dtype) or torch.float32))
    m.arange = lambda n: _w(torch.arange(n))
    m.linspace = lambda a, b, n: _w(torch.linspace(a, b, n, dtype=torch.float32))
    m.where = lambda c, a, b: torch.where(_w(c), a, b)
    m.expm1 = lambda x: torch.expm1(_w(x))
    m.sin = lambda x: torch.sin(_w(x))
    m.cos = lambda x: torch.cos(_w(x))
    m.einsum = lambda expr, *ops, precision=None: torch.einsum(expr, *[_w(o) for o in ops])
    m.concatenate = lambda xs, axis=0: torch.cat([_w(x) for x in xs], dim=axis)
    m.nan_to_num = lambda x, nan=0.0, posinf=None, neginf=None: torch.nan_to_num(_w(x), nan=nan, posinf=posinf, neginf=neginf)
    return m


def _promote(a, b):
    dt = torch.promote_types(a.dtype, b.dtype)
    return a.to(dt), b.to(dt)


def _vjp(fn, x, has_aux=False):
    """jax.vjp(fn, x, has_aux=True) -> (out, vjp_fun, aux); vjp_fun(ct) -> (J^T ct,)."""
    xl = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        res = fn(xl)
    out, aux = res if has_aux else (res, None)

    def vjp_fun(ct):
        (g,) = torch.autograd.grad(out, xl, grad_outputs=ct)
        return (g,)

    return (out.detach(), vjp_fun, aux.detach()) if has_aux else (out.detach(), vjp_fun)


def _scan(f, init, xs=None, length=None):
    """jax.lax.scan with xs=None: python scalars in the carry become float32 arrays, as under jit."""
    carry = tuple(_w(c, torch.float32) if isinstance(c, float) else c for c in init)
    for _ in range(length):
        carry, _ = f(carry, None)
    return carry, None


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


@contextlib.contextmanager
def _temporary_modules(mods: dict):
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load():
    """The reference's `pi0_rtc` module, executed in place (cached).  sys.modules is left as it was found."""
    global _loaded
    if _loaded is not None:
        return _loaded

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__getattr__ = lambda attr: _Anything()  # PEP 562: any other name (type annotations) resolves to a dummy
        return m

    jnp = _make_jnp()
    jax = stub("jax", numpy=jnp, vjp=_vjp,
               lax=types.SimpleNamespace(scan=_scan, Precision=types.SimpleNamespace(HIGHEST=None)))
    nnx = stub("flax.nnx", swish=F.silu)
    bridge = stub("flax.nnx.bridge")
    nnx.bridge = bridge
    flax = stub("flax", nnx=nnx)
    flax.__path__ = []
    nnx.__path__ = []

    class BaseModel:  # models/model.py: only the base-class slot and three names are needed at import time
        pass

    model = stub("openpi.models.model", BaseModel=BaseModel, preprocess_observation=lambda rng, obs, train=False: obs)
    at = stub("openpi.shared.array_typing", typecheck=lambda f: f)
    cfgm, gemma, siglip = stub("openpi.models.pi0_config"), stub("openpi.models.gemma"), stub("openpi.models.siglip")
    models = stub("openpi.models", model=model, pi0_config=cfgm, gemma=gemma, siglip=siglip)
    models.__path__ = []
    shared = stub("openpi.shared", array_typing=at)
    shared.__path__ = []
    openpi = stub("openpi", models=models, shared=shared)
    openpi.__path__ = []
    mods = {"jax": jax, "jax.numpy": jnp, "flax": flax, "flax.nnx": nnx, "flax.nnx.bridge": bridge, "openpi": openpi,
            "openpi.models": models, "openpi.models.model": model, "openpi.models.pi0_config": cfgm,
            "openpi.models.gemma": gemma, "openpi.models.siglip": siglip, "openpi.shared": shared,
            "openpi.shared.array_typing": at}
    with _temporary_modules(mods):
        spec = importlib.util.spec_from_file_location("_kai0_reference_pi0_rtc", PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _loaded = mod
    return mod


# ------------------------------------------------------------------------------------------------------------
# a Pi0RTC instance whose sub-networks are the PyTorch-path oracle network
# ------------------------------------------------------------------------------------------------------------
class _Linear:
    def __init__(self, w, b):
        self.w, self.b, self.out_features = w, b, w.shape[0]

    def __call__(self, x):
        return F.linear(x, self.w, self.b)


class _Nets(dict):
    """`self.PaliGemma`: nnx.Dict(llm=..., img=...) -- attribute access on a dict."""

    __getattr__ = dict.__getitem__


def make_model(params: dict, cfg):
    """Pi0RTC instance (no __init__) over the oracle network with parameters `params` (pi0.5 branch)."""
    from oracle import pi05_oracle as O

    mod = load()
    lm = "paligemma_with_expert.paligemma.model.language_model."

    def img(image, train=False):
        return O.siglip_embed_image(params, cfg, image), None

    def llm(tokens, mask=None, positions=None, kv_cache=None, adarms_cond=None, method=None):
        if method == "embed":  # Embedder.encode: table lookup scaled by sqrt(width) (the PyTorch path: pi0_pytorch.py:213-216)
            emb = F.embedding(tokens, params[lm + "embed_tokens.weight"])
            return emb * math.sqrt(emb.shape[-1])
        prefix, suffix = tokens
        mask4d = O.prepare_attention_masks_4d(mask.to(torch.bool))
        if suffix is None:  # prefix pass that fills the cache (pi0_rtc.py:263-266)
            out, cache = O.single_stream_forward(params, cfg, "prefix", prefix, mask4d, positions, use_cache=True)
            return (out, None), cache
        assert prefix is None and kv_cache is not None
        out, _ = O.single_stream_forward(params, cfg, "suffix", suffix, mask4d, positions, past_kv=kv_cache,
                                         use_cache=False, adarms_cond=adarms_cond[1])
        return (None, out), None

    m = object.__new__(mod.Pi0RTC)
    m.pi05 = True
    m.action_dim, m.action_horizon, m.max_token_len = cfg.action_dim, cfg.action_horizon, cfg.max_token_len
    m.PaliGemma = _Nets(llm=llm, img=img)
    m.action_in_proj = _Linear(params["action_in_proj.weight"], params["action_in_proj.bias"])
    m.time_mlp_in = _Linear(params["time_mlp_in.weight"], params["time_mlp_in.bias"])
    m.time_mlp_out = _Linear(params["time_mlp_out.weight"], params["time_mlp_out.bias"])
    out_w, out_b = params["action_out_proj.weight"], params["action_out_proj.bias"]
    m.action_out_proj = lambda x: F.linear(x.to(torch.float32), out_w, out_b)  # pi0_pytorch.py:364-371 casts first
    m.deterministic = True
    return m


class Obs:
    """What `Pi0RTC.embed_prefix` / `sample_actions` read from an observation (models/model.py:84-119)."""

    def __init__(self, batch, keys=("base_0_rgb", "left_wrist_0_rgb", "right_wrist_0_rgb")):
        self.images = {k: _w(batch["images"][i]) for i, k in enumerate(keys)}
        self.image_masks = {k: _w(batch["img_masks"][i]) for i, k in enumerate(keys)}
        self.state = _w(torch.zeros(batch["tokens"].shape[0], 32))
        self.tokenized_prompt = _w(batch["tokens"])
        self.tokenized_prompt_mask = _w(batch["token_mask"])


@torch.no_grad()
def sample_actions(params, cfg, batch, noise, *, oracle_suffix_embedding: bool = False, **kw):
    """`Pi0RTC.sample_actions(rng=None, observation, noise=noise, **kw)` of the reference, as a plain torch tensor.
    oracle_suffix_embedding=True replaces `Pi0RTC.embed_suffix` (whose sincos embedding is computed in float32 by the JAX
    code, in float64 by the PyTorch path) with the PyTorch path's, so that the sampler logic can be compared exactly."""
    from oracle import pi05_oracle as O

    m = make_model(params, cfg)
    if oracle_suffix_embedding:
        m.embed_suffix = lambda obs, x_t, t: O.embed_suffix(params, cfg, x_t, t)
    out = m.sample_actions(None, Obs(batch), noise=_w(noise), **kw)
    return out.as_subclass(torch.Tensor)
