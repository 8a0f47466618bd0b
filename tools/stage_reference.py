"""Offline install of the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box).

The contract's `python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target
baseline/_ref /root/reference` fails in this image: the reference's build backend (`hatchling`, pyproject.toml:131-133)
is neither installed nor in /opt/wheelhouse.  The wheel hatchling would build is a plain copy of `src/openpi/**` (pure
Python), so this script performs exactly that copy by hand: every file of /root/reference/src/openpi goes, byte for byte,
to baseline/_ref/openpi, and the commit it came from is recorded.  Nothing under baseline/_ref is tracked by git, nothing
of it is imported by the product (kai0_b200/): it exists so that `bench.py --impl reference`, `cpu_baseline` and the
`reference_gpu` anchor can run the reference's OWN PI0Pytorch (through tools/reference_loader.py) on the GPU box, where
/root/reference does not exist.

    python tools/stage_reference.py
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/openpi"
DST = os.path.join(ROOT, "baseline", "_ref", "openpi")
SCRIPT_SRC = "/root/reference/scripts/train_pytorch.py"
SCRIPTS_DST = os.path.join(ROOT, "baseline", "_ref", "scripts")
CLIENT_SRC = "/root/reference/packages/openpi-client/src/openpi_client"
CLIENT_DST = os.path.join(ROOT, "baseline", "_ref", "openpi_client")


def stage() -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # the training entry point itself (not part of the wheel, staged beside it): tests/test_zzzz_train_script_gpu.py runs ITS
    # `train_loop` around the engine on the GPU box (tools/reference_train_harness.py)
    os.makedirs(SCRIPTS_DST, exist_ok=True)
    shutil.copyfile(SCRIPT_SRC, os.path.join(SCRIPTS_DST, "train_pytorch.py"))
    # the client-side helper package the serving code imports (image_tools, msgpack_numpy, base_policy, websocket client)
    if os.path.isdir(CLIENT_DST):
        shutil.rmtree(CLIENT_DST)
    shutil.copytree(CLIENT_SRC, CLIENT_DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    h = hashlib.sha256()
    n = 0
    for d, _, files in sorted(os.walk(DST)):
        for f in sorted(files):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
            n += 1
    head = ""
    for cand in ("/root/reference/.git/HEAD",):
        if os.path.exists(cand):
            head = open(cand).read().strip()
    with open(os.path.join(ROOT, "baseline", "_ref", "STAGED_FROM.txt"), "w") as f:
        f.write(f"source: {SRC}\nfiles: {n}\nsha256(all files, sorted): {h.hexdigest()}\ngit HEAD: {head}\n"
                "how: tools/stage_reference.py (pip --target install impossible: hatchling absent)\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else "no reference checkout and no staged copy", DST)
    sys.exit(0 if ok else 1)
