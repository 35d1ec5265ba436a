"""Stages the UNMODIFIED reference under oracle/_ref/ so that it can travel to the GPU box.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference (Zachary-Luk/3D-Denoising-Diffusion-Model) is pure Python;
its sampling path lives in nine modules of `guided_diffusion/` that import cleanly with the torch / numpy of this image
(dist_util / image_datasets / train_util need mpi4py, blobfile, SimpleITK and are NOT staged).
`build_ref()` copies those modules byte for byte from the reference checkout into `oracle/_ref/guided_diffusion/` and
writes their SHA-256 into `oracle/_ref/MANIFEST.json`.  `oracle/_ref/` is a git-ignored build output (like the compiled
library): no reference source enters the repository's history, but the directory ships with the `gpurun` snapshot, so

  * `bench.py --impl reference` times the reference's own `sr_create_model_and_diffusion` + `p_sample` on the host cores,
  * the `cpu_baseline` / `library_bar` legs of `bench.py` run the reference's own nn.Module (CPU fp32; GPU fp16 + cuDNN),

instead of the oracle's restatement.  Where `oracle/_ref/` is absent (a fresh clone without /root/reference) those legs
fall back to the restatement and say `kind: "port"`.

    python -m oracle.build_ref [--src /root/reference] [--check]
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
DEFAULT_SRC = "/root/reference"
# the import closure of guided_diffusion.script_util.sr_create_model_and_diffusion
MODULES = ["__init__.py", "script_util.py", "unet.py", "gaussian_diffusion.py", "respace.py", "nn.py", "fp16_util.py",
           "logger.py", "losses.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build_ref(src: str = DEFAULT_SRC, quiet: bool = False):
    """Copies the modules; returns REF_DIR, or None when there is no reference checkout (the GPU box: the staged copy,
    if any, is used as it is)."""
    pkg = os.path.join(src, "guided_diffusion")
    if not os.path.isdir(pkg):
        return REF_DIR if available() else None
    dst = os.path.join(REF_DIR, "guided_diffusion")
    os.makedirs(dst, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for m in MODULES:
        shutil.copyfile(os.path.join(pkg, m), os.path.join(dst, m))
        manifest["files"]["guided_diffusion/" + m] = _sha(os.path.join(dst, m))
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(MODULES)} reference modules under {REF_DIR}")
    return REF_DIR


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "MANIFEST.json"))


def verify(src: str = DEFAULT_SRC) -> bool:
    """The staged files still have the recorded hashes (and equal the checkout, when one is present)."""
    if not available():
        return False
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for rel, sha in manifest["files"].items():
        if _sha(os.path.join(REF_DIR, rel)) != sha:
            return False
        orig = os.path.join(src, rel)
        if os.path.isfile(orig) and _sha(orig) != sha:
            return False
    return True


def load_ref():
    """Imports the staged reference package (as `guided_diffusion`) and returns its script_util module, or None."""
    if not available():
        return None
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    mod = importlib.import_module("guided_diffusion.script_util")
    if not os.path.abspath(mod.__file__).startswith(REF_DIR):
        raise RuntimeError(f"guided_diffusion resolved to {mod.__file__}, not to the staged reference")
    return mod


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else DEFAULT_SRC
    if "--check" in sys.argv:
        print("staged reference verified" if verify(src) else "staged reference missing or modified")
    else:
        build_ref(src)
