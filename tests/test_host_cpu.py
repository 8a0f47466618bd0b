"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/pi05.h, the module mirrors
the reference's state_dict / dtype contract, and error behaviour matches the reference's (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

import helpers as H
from kai0_b200 import _lib
from kai0_b200.model import Observation
from kai0_b200.pi0_pytorch import PI0Pytorch, Pi05EngineConfig, get_gemma_config
from oracle import pi05_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "pi05.h")).read()
    declared = set(re.findall(r"\b(pi05_[a-z0-9_]+)\s*\(", header))
    declared -= {"pi05_engine"}
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"libpi05.so does not export {sym}"
    for sym in _lib.EXPORTS:
        assert sym in declared, f"{sym} is bound in _lib.py but not declared in include/pi05.h"
    assert lib.pi05_abi_version() == 2


def test_abi_struct_sizes_match_header():
    """The ctypes mirrors must have the C layout (compile a tiny C program against include/pi05.h)."""
    src = r"""
    #include <stdio.h>
    #include "pi05.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(pi05_config), sizeof(pi05_param), sizeof(pi05_batch),
                           sizeof(pi05_gemm_desc), sizeof(pi05_gemma_cfg)); return 0; }
    """
    d = os.path.join(ROOT, "build")
    os.makedirs(d, exist_ok=True)
    c = os.path.join(d, "abi_sizes.c")
    open(c, "w").write(src)
    exe = os.path.join(d, "abi_sizes")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    got = [C.sizeof(_lib.Config), C.sizeof(_lib.Param), C.sizeof(_lib.Batch), C.sizeof(_lib.GemmDesc),
           C.sizeof(_lib.GemmaCfg)]
    assert [int(x) for x in out] == got


def test_workspace_bytes_is_pure_and_validates():
    lib = _lib.lib()
    c = _lib.Config()
    assert lib.pi05_workspace_bytes(C.byref(c)) == 0  # all-zero config is rejected
    assert "width" in _lib.last_error() or "geometry" in _lib.last_error()


def test_state_dict_contract_matches_reference_names_and_dtypes():
    oc = O.tiny_config()
    model, params = H.build_pair(oc, device=None)
    sd = model.state_dict()
    spec = O.param_specs(oc)
    for k, (shape, dt) in spec.items():
        assert tuple(sd[k].shape) == shape and sd[k].dtype == dt, k
    # gemma_pytorch.py:72-79 dtype map
    assert sd["paligemma_with_expert.paligemma.model.vision_tower.vision_model.embeddings.patch_embedding.weight"].dtype == torch.float32
    assert sd["paligemma_with_expert.paligemma.model.language_model.layers.0.input_layernorm.weight"].dtype == torch.float32
    assert sd["paligemma_with_expert.gemma_expert.model.layers.0.input_layernorm.dense.weight"].dtype == torch.float32
    assert sd["paligemma_with_expert.paligemma.model.language_model.layers.0.mlp.up_proj.weight"].dtype == torch.bfloat16
    assert sd["action_in_proj.weight"].dtype == torch.float32
    # tied lm_head (same storage), unused expert lm_head present for checkpoint round-trips
    assert sd["paligemma_with_expert.paligemma.lm_head.weight"].data_ptr() == \
        sd["paligemma_with_expert.paligemma.model.language_model.embed_tokens.weight"].data_ptr()
    assert "paligemma_with_expert.gemma_expert.lm_head.weight" in sd
    # weights loaded from the oracle dict are bit-identical
    for k, v in params.items():
        assert torch.equal(sd[k], v), k


def test_full_size_parameter_count():
    from kai0_b200.pi0_pytorch import parameter_table

    cfg = Pi05EngineConfig()
    table = parameter_table(cfg, get_gemma_config("gemma_2b"), get_gemma_config("gemma_300m"))
    total = sum(torch.Size(s).numel() for _, s, _, _ in table)
    unused = 257152 * 1024
    assert abs((total - unused) / 1e9 - 3.353) < 0.002  # 3.353 B trainable


def test_fused_arena_layout_is_contiguous():
    model, _ = H.build_pair(O.tiny_config(), device=None)
    l0 = model.paligemma_with_expert.paligemma.model.language_model.layers[0]
    q, k, v = l0.self_attn.q_proj.weight, l0.self_attn.k_proj.weight, l0.self_attn.v_proj.weight
    assert q.data_ptr() + q.numel() * 2 == k.data_ptr() and k.data_ptr() + k.numel() * 2 == v.data_ptr()
    g, u = l0.mlp.gate_proj.weight, l0.mlp.up_proj.weight
    assert g.data_ptr() + g.numel() * 2 == u.data_ptr()
    for p in model.parameters():
        assert p.data_ptr() % 16 == 0


def test_module_surface_and_error_behaviour():
    oc = O.tiny_config()
    model, _ = H.build_pair(oc, device=None)
    assert hasattr(model, "gradient_checkpointing_enable")
    model.gradient_checkpointing_enable()
    assert model.is_gradient_checkpointing_enabled()
    assert model.paligemma_with_expert.to_bfloat16_for_selected_params("bfloat16") is model.paligemma_with_expert
    with pytest.raises(ValueError):
        model.paligemma_with_expert.to_bfloat16_for_selected_params("float32")
    with pytest.raises(TypeError):
        model.float()
    batch = O.synthetic_batch(oc, 2)
    obs = H.Obs(batch)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(obs, batch["actions"])  # the product path never falls back to the CPU
    del obs.images["base_0_rgb"]
    with pytest.raises(ValueError, match="missing keys"):
        model._preprocess_observation(obs)
    with pytest.raises(ValueError):
        PI0Pytorch(Pi05EngineConfig(pi05=False))


def test_observation_from_dict_matches_reference_conversion():
    u8 = torch.randint(0, 256, (2, 8, 8, 3), dtype=torch.uint8)
    d = {"image": {"base_0_rgb": u8}, "image_mask": {"base_0_rgb": torch.ones(2, dtype=torch.bool)},
         "state": torch.zeros(2, 32), "tokenized_prompt": torch.zeros(2, 4, dtype=torch.int64),
         "tokenized_prompt_mask": torch.ones(2, 4, dtype=torch.bool)}
    obs = Observation.from_dict(d)
    ref = u8.to(torch.float32).permute(0, 3, 1, 2) / 255.0 * 2.0 - 1.0  # models/model.py:132-133
    assert torch.equal(obs.images["base_0_rgb"], ref)
    with pytest.raises(ValueError):
        Observation.from_dict({"image": {}, "image_mask": {}, "state": None, "tokenized_prompt": 1})


def test_product_code_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kai0_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "pi05_oracle" in txt:
                    bad.append(f)
    assert not bad, bad


def test_bench_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_advantage_estimator_surface():
    """pi0_pytorch.py:464-498 and preprocessing_pytorch.py:196-204: extra value_head.* parameters (fp32), loss weights
    read from the config, image keys sorted by (timestep, part) whatever the dict order."""
    from kai0_b200.pi0_pytorch import AdvantageEstimator

    oc = O.tiny_config(value_head=True)
    model, params = H.build_pair(oc, device=None, cls=AdvantageEstimator)
    sd = model.state_dict()
    for k in ("value_head.0.weight", "value_head.0.bias", "value_head.2.weight", "value_head.2.bias",
              "value_head.4.weight", "value_head.4.bias"):
        assert sd[k].dtype == torch.float32 and sd[k].shape == params[k].shape
    assert model.loss_value_weight == 0.0 and model.loss_action_weight == 1.0  # getattr defaults, :467-468
    keys = {"right_wrist_0_rgb": 0, "base_-100_rgb": 0, "left_wrist_0_rgb": 0, "base_0_rgb": 0, "left_wrist_-100_rgb": 0}
    assert model._image_keys(keys) == ["base_-100_rgb", "left_wrist_-100_rgb", "base_0_rgb", "left_wrist_0_rgb",
                                       "right_wrist_0_rgb"]
    with pytest.raises(ValueError, match="not of the form"):
        model._image_keys({"top_head": 0})
    batch = O.synthetic_batch(oc, 2)
    obs = H.Obs(batch)
    with pytest.raises(ValueError, match="progress"):
        model(obs, batch["actions"])


def test_every_kernel_honours_the_pdl_contract():
    """launch.h: a kernel launched with the programmatic-stream-serialization attribute must wait for its predecessor
    (pdl_enter / pdl_wait) before touching global memory.  Static check of the sources: every __global__ kernel body
    contains the call, and no launch bypasses launch_pdl with a raw <<<>>>."""
    csrc = os.path.join(ROOT, "kai0_b200", "csrc")
    missing, raw = [], []
    for fn in sorted(os.listdir(csrc)):
        if not fn.endswith((".cu", ".cuh")):
            continue
        src = open(os.path.join(csrc, fn)).read()
        if "<<<" in re.sub(r"//[^\n]*", "", src):
            raw.append(fn)
        for m in re.finditer(r"__global__[^{;]*\{", src):
            # body = up to the matching closing brace
            depth, i = 1, m.end()
            while depth and i < len(src):
                depth += {"{": 1, "}": -1}.get(src[i], 0)
                i += 1
            body = src[m.end():i]
            if "pdl_enter()" not in body and "pdl_wait()" not in body:
                name = re.findall(r"(\w+)\s*\(", m.group(0))
                missing.append(f"{fn}:{name[-1] if name else '?'}")
    assert not missing, missing
    assert not raw, raw


def test_submodule_moves_and_assigning_loads_are_refused():
    """`.to()` on a sub-module or `load_state_dict(assign=True)` would silently detach parameters from the
    flat arenas the engine reads; both raise instead."""
    import pytest
    import torch

    from kai0_b200.pi0_pytorch import GemmaVariant, PI0Pytorch, Pi05EngineConfig

    cfg = Pi05EngineConfig(paligemma_variant=GemmaVariant(64, 2, 128, 8, 1, 16),
                           action_expert_variant=GemmaVariant(32, 2, 64, 8, 1, 16), vit_width=32, vit_depth=1,
                           vit_mlp_dim=64, vit_heads=2, image_size=28, vocab_size=64)
    m = PI0Pytorch(cfg)
    with pytest.raises(RuntimeError, match="whole PI0Pytorch"):
        m.paligemma_with_expert.to(torch.float32)
    with pytest.raises(ValueError, match="assign"):
        m.load_state_dict(m.state_dict(), assign=True)
    # the supported forms still work and keep the parameters views of the arenas
    m.load_state_dict(m.state_dict())
    m.to("cpu")
    p = m.action_in_proj.weight
    lo = m._flat[torch.float32].data_ptr()
    assert lo <= p.data_ptr() < lo + m._flat[torch.float32].numel() * 4


def test_effective_token_len_rules():
    """Host logic of the prompt padding removal (pi05_batch.token_len): longest left-aligned prompt rounded up to 8, the
    full length for masks with holes, while taps are recorded, or when switched off."""
    import torch

    from kai0_b200.pi0_pytorch import GemmaVariant, PI0Pytorch, Pi05EngineConfig

    cfg = Pi05EngineConfig(paligemma_variant=GemmaVariant(64, 2, 128, 8, 1, 16),
                           action_expert_variant=GemmaVariant(32, 2, 64, 8, 1, 16), vit_width=32, vit_depth=1,
                           vit_mlp_dim=64, vit_heads=2, image_size=28, vocab_size=64, max_token_len=40)
    m = PI0Pytorch(cfg)
    mask = torch.zeros(3, 40, dtype=torch.bool)
    mask[0, :19] = True
    mask[1, :3] = True
    assert m._effective_token_len(mask) == 24
    assert m._effective_token_len(torch.zeros(2, 40, dtype=torch.bool)) == 8
    assert m._effective_token_len(torch.ones(2, 40, dtype=torch.bool)) == 40
    hole = mask.clone()
    hole[0, 5] = False
    assert m._effective_token_len(hole) == 40
    m.set_taps(True)
    assert m._effective_token_len(mask.clone()) == 40
    m.set_taps(False)
    m.skip_prompt_padding = False
    assert m._effective_token_len(mask.clone()) == 40


def _load_bench():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_bookkeeping():
    """bench.py's numbers that are not measured: both arms print the same `config`; executed FLOPs reduce to the documented
    dense figure at 968 prefix rows and scale down with prompt padding removal; the ncu traffic table falls back to the same
    GEMM class with the capture shape stated."""
    b = _load_bench()
    assert b.bench_config(8, 32, False) == b.bench_config(8, 32, False)
    assert b.bench_config(1, 32, False)["parallelism"] == "dp1" and b.bench_config(8, 32, False)["global_batch"] == 256
    dense = b.executed_train_tflop_per_sample(968)
    assert abs(dense - b.EXECUTED_TRAIN_TFLOP_PER_SAMPLE) < 5e-3 and dense < b.TRAIN_TFLOP_PER_SAMPLE
    assert 12.0 < b.executed_train_tflop_per_sample(864) < dense
    v, note = b.ncu_traffic(32768, 2048, 30976, 0, 3)
    assert v == b.NCU_TRAFFIC_BYTES[(32768, 2048, 30976, 0, 3)] and "this shape" in note
    v2, note2 = b.ncu_traffic(32768, 2048, 27648, 0, 3)
    assert v2 == v and "dense prefix" in note2
    assert b.ncu_traffic(7, 7, 7, 0, 0)[0] is None
    # the serving leg runs in a child process and can only ever ADD a record: without a GPU it reports why, nothing raises
    if not torch.cuda.is_available():
        leg = b.serving_leg(timeout_s=120)
        assert set(leg) == {"unavailable"} and "serving_probe.py" in leg["unavailable"]


def test_staged_reference_is_byte_identical_to_the_checkout():
    """baseline/_ref (git-ignored; what bench.py's reference legs execute on the GPU box) must be the UNMODIFIED package."""
    import filecmp

    import pytest

    src, dst = "/root/reference/src/openpi", os.path.join(ROOT, "baseline", "_ref", "openpi")
    if not (os.path.isdir(src) and os.path.isdir(dst)):
        pytest.skip("needs the reference checkout and a staged copy (tools/stage_reference.py)")
    n = 0
    for d, _, files in os.walk(src):
        if "__pycache__" in d:
            continue
        for f in files:
            if f.endswith(".pyc"):
                continue
            rel = os.path.relpath(os.path.join(d, f), src)
            assert filecmp.cmp(os.path.join(d, f), os.path.join(dst, rel), shallow=False), rel
            n += 1
    assert n > 20
    staged_script = os.path.join(ROOT, "baseline", "_ref", "scripts", "train_pytorch.py")
    assert filecmp.cmp("/root/reference/scripts/train_pytorch.py", staged_script, shallow=False)


def test_shipped_library_is_blackwell_native_sass():
    """Static evidence in the built library (B200_PROFILING.md "What proves a Blackwell-native kernel"): every embedded
    cubin is sm_100a; the tensor-core work is tcgen05 (`UTCHMMA`, incl. the 2-CTA form) fed by TMA (`UTMALDG`) with TMEM
    read-back (`LDTM`); no legacy `mma.sync` / `wmma` (`HMMA`) and no Hopper `wgmma` (`HGMMA`) anywhere; no kernel spills to
    local memory and stack frames stay tiny."""
    import re
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    so = os.path.join(ROOT, "kai0_b200", "libpi05.so")
    if not os.path.exists(cuobjdump) or not os.path.exists(so):
        pytest.skip("needs cuobjdump and the built kai0_b200/libpi05.so")
    elfs = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True, check=True).stdout
    names = re.findall(r"ELF file\s+\d+:\s+(\S+)", elfs)
    assert len(names) >= 15 and all(n.endswith(".sm_100a.cubin") for n in names), names
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    ops = set(re.findall(r"\b([A-Z][A-Z0-9]+(?:\.[A-Za-z0-9_]+)*)\b", sass))
    base = {o.split(".")[0] for o in ops}
    assert "UTCHMMA" in base and any(o.startswith("UTCHMMA.2CTA") for o in ops)       # tcgen05.mma, cta_group::1 and ::2
    assert any(o.startswith("UTMALDG") for o in ops) and any(o.startswith("UTMALDG") and "2CTA" in o for o in ops)  # TMA loads
    assert any(o.startswith("LDTM") for o in ops)                                      # tcgen05.ld (TMEM -> registers)
    assert any(o.startswith("UTCBAR") and "MULTICAST" in o for o in ops)               # tcgen05.commit multicast to the CTA pair
    assert not base & {"HMMA", "HGMMA", "QGMMA", "IGMMA", "IMMA", "BMMA"}, base & {"HMMA", "HGMMA", "QGMMA", "IGMMA", "IMMA", "BMMA"}
    usage = subprocess.run([cuobjdump, "-res-usage", so], capture_output=True, text=True, check=True).stdout
    recs = re.findall(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert len(recs) >= 150
    assert all(int(local) == 0 for _, _, _, local in recs)         # nothing spills
    assert max(int(stack) for _, stack, _, _ in recs) <= 64        # parameter-copy frames only
    assert max(int(reg) for reg, _, _, _ in recs) <= 255

