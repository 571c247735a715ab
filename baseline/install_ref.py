#!/usr/bin/env python
"""Stage the UNMODIFIED reference (HaoZhongkai/DPOT) under baseline/_ref/ so that it travels to the GPU box.

    python baseline/install_ref.py            # needs /root/reference (build container only)

The contract's `pip install --target baseline/_ref /root/reference` was tried first and fails:
"Directory '/root/reference' is not installable. Neither 'setup.py' nor 'pyproject.toml' found." -- the reference is
a flat script tree, not a package.  The files its hot path consists of (models/dpot.py, utils/optimizer.py,
utils/criterion.py and the two package __init__ files) are therefore copied byte for byte; baseline/_ref/ is
git-ignored (no reference source enters the history) but NOT gpurun-ignored, so `bench.py --impl reference` and the
`gpu_eager_baseline` leg can import the real reference on the GPU box, where /root/reference does not exist.
A SHA-256 manifest is written next to the files; bench.py reports it so that a reader can check they are unmodified.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("DPOT_REFERENCE", "/root/reference")
FILES = ["models/__init__.py", "models/dpot.py", "utils/__init__.py", "utils/optimizer.py", "utils/criterion.py"]


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"install_ref: {SRC} not present (GPU box?) -- keeping whatever is under {DST}")
        return os.path.exists(os.path.join(DST, "models", "dpot.py"))
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"install_ref: {len(FILES)} reference files staged under {DST}")
    return True


def import_reference():
    """(DPOTNet, Adam, SimpleLpLoss, manifest) of the staged reference, or None when it is absent."""
    if not os.path.exists(os.path.join(DST, "models", "dpot.py")):
        return None
    sys.dont_write_bytecode = True
    # the reference's package names (models, utils) are generic: import them from _ref only, then restore sys.path
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.") or
             k == "utils" or k.startswith("utils.")}
    sys.path.insert(0, DST)
    try:
        from models.dpot import DPOTNet
        from utils.criterion import SimpleLpLoss
        from utils.optimizer import Adam
    finally:
        sys.path.remove(DST)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            sys.modules["_dpot_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    manifest = json.load(open(os.path.join(DST, "MANIFEST.json"))) if os.path.exists(os.path.join(DST, "MANIFEST.json")) else {}
    return DPOTNet, Adam, SimpleLpLoss, manifest


if __name__ == "__main__":
    ok = install()
    sys.exit(0 if ok else 1)
