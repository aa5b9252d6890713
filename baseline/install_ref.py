#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE — puts an UNMODIFIED copy of the reference's env path where it can travel.

    python baseline/install_ref.py            # needs /root/reference (the build container)

The reference is pure Python with no build step; `pip install --target baseline/_ref /root/reference` would install
only the `merging_gym` package (setup.py:1-6) and leave out `scripts/helper.py`, which the package imports
(`from helper import mpc_1d`, merging_env.py:16).  This recipe therefore copies, byte for byte,
    /root/reference/merging_gym/**          -> baseline/_ref/merging_gym/**
    /root/reference/scripts/helper.py       -> baseline/_ref/scripts/helper.py
into the git-ignored `baseline/_ref/` (never committed; it ships to the GPU box with the working tree like the
built `.so` files) and writes `baseline/_ref/MANIFEST.json` with the sha256 of every file so that a run can show
the copy is the reference's own code.  `oracle/ref_loader.py` falls back to this tree when `/root/reference` is
absent (the GPU box), which is what lets `bench.py` time the reference's own `MergeEnv.step` on the box's host cores.
gym / pygame / shapely / qpsolvers are not installed anywhere: the copy runs against `oracle/ref_shims/` exactly as
the golden-vector generator does.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MERGING_GYM_REFERENCE", "/root/reference")


def install(src: str = SRC, dest: str = DEST) -> dict:
    if not os.path.isfile(os.path.join(src, "merging_gym", "envs", "merging_env.py")):
        raise RuntimeError(f"reference tree not found at {src}")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(os.path.join(dest, "scripts"))
    shutil.copytree(os.path.join(src, "merging_gym"), os.path.join(dest, "merging_gym"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy2(os.path.join(src, "scripts", "helper.py"), os.path.join(dest, "scripts", "helper.py"))
    manifest = {}
    for root, _, files in os.walk(dest):
        for f in sorted(files):
            p = os.path.join(root, f)
            rel = os.path.relpath(p, dest)
            with open(p, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
            with open(os.path.join(src, rel), "rb") as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == manifest[rel], rel
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = install()
    print(f"copied {len(m)} reference files into {DEST}")
    sys.exit(0)
