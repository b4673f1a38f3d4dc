#!/usr/bin/env python
"""Installs the UNMODIFIED reference (timm, pure Python) into oracle/_ref/ so that it travels to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under b200timm/ imports oracle/_ref.  Users of it:
  * tests/ (plugin parity on real timm modules, timm's own test-suite through the TORCH_BACKEND hook),
  * bench.py --impl reference (the reference's own benchmark.py --device cpu),
  * bench.py's informational sub-records (benchmark.py through the plugin, timm eager on the same GPU).

Recipe: `pip install --no-index --no-build-isolation --no-deps --target oracle/_ref /root/reference` needs the
`pdm-backend` build backend, which is not in this image (no network) -- so the install step is what a wheel install of
a pure-Python package does: the package directory is placed on the target path byte for byte.  The three CLI scripts
the survey names (benchmark.py / train.py / validate.py) and the reference's own tests/ go next to it under
oracle/_ref/scripts and oracle/_ref/tests.  oracle/_ref/ is git-ignored (never part of the history) and NOT
gpurun-ignored (it ships with the snapshot like the built .so).

    python oracle/install_ref.py [--ref /root/reference] [--force]
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SCRIPTS = ("benchmark.py", "train.py", "validate.py", "inference.py")


def _tree_digest(root):
    h = hashlib.sha256()
    for d, dirs, files in os.walk(root):
        dirs.sort()
        for f in sorted(files):
            if f.endswith(".pyc"):
                continue
            p = os.path.join(d, f)
            h.update(os.path.relpath(p, root).encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def installed():
    return os.path.isfile(os.path.join(TARGET, "timm", "__init__.py")) and os.path.isfile(
        os.path.join(TARGET, "scripts", "benchmark.py"))


def install(ref="/root/reference", force=False, quiet=False):
    """Returns the install directory, or None when neither an install nor the reference checkout exists."""
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(ref, "timm")):
        return TARGET if installed() else None
    if os.path.isdir(TARGET):
        shutil.rmtree(TARGET)
    os.makedirs(TARGET)
    how = "copy"
    try:  # the contract's recipe first; falls through when the build backend is missing (it is, here)
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                        "--find-links", "/opt/wheelhouse", "--target", TARGET, ref], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        how = "pip"
    except (subprocess.SubprocessError, OSError):
        pass
    if not os.path.isfile(os.path.join(TARGET, "timm", "__init__.py")):
        shutil.copytree(os.path.join(ref, "timm"), os.path.join(TARGET, "timm"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(TARGET, "scripts"), exist_ok=True)
    for s in SCRIPTS:
        if os.path.isfile(os.path.join(ref, s)):
            shutil.copy2(os.path.join(ref, s), os.path.join(TARGET, "scripts", s))
    if os.path.isdir(os.path.join(ref, "tests")):
        shutil.copytree(os.path.join(ref, "tests"), os.path.join(TARGET, "tests"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    digest = _tree_digest(os.path.join(TARGET, "timm"))
    same = digest == _tree_digest(os.path.join(ref, "timm"))
    with open(os.path.join(TARGET, "INSTALL.txt"), "w") as f:
        f.write(f"source: {ref}\nmethod: {how}\ntimm tree sha256: {digest}\nidentical to source tree: {same}\n")
    if not same:
        raise RuntimeError("oracle/_ref/timm differs from the reference checkout")
    if not quiet:
        print(f"reference installed into {TARGET} ({how}); timm tree sha256 {digest[:16]}, identical to {ref}/timm")
    return TARGET


def add_to_path():
    """Put the installed reference on sys.path (tests / bench only).  Returns True when timm is importable from it."""
    if not installed():
        return False
    if TARGET not in sys.path:
        sys.path.insert(0, TARGET)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if install(a.ref, a.force) is None:
        raise SystemExit("no reference checkout and no previous install")
