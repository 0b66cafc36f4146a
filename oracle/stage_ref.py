"""ORACLE / CPU BASELINE staging - test and benchmark infrastructure only.

Stages the reference's own, UNMODIFIED implementation of the hot path (joliang17/FLYP clip/loss.py: gather_features and
ClipLoss) from the read-only reference checkout into the git-ignored directory oracle/_ref/, so that it travels to the
GPU box with the snapshot (where /root/reference does not exist) and bench.py can time the reference ITSELF on the
host cores ("cpu_baseline.kind" = "reference") and the golden vectors can be regenerated.  Nothing is copied into the
tracked tree; nothing under flyp_b200/ ever imports it.

    python oracle/stage_ref.py            # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("FLYP_REFERENCE_ROOT", "/root/reference")
DST_DIR = os.path.join(HERE, "_ref", "clip")
FILES = ["clip/loss.py"]          # the path of SURVEY.md section 8(a); it imports only torch


def stage(verbose: bool = False) -> bool:
    """Copy the reference files byte for byte.  Returns False when the reference checkout is not present."""
    src_ok = all(os.path.exists(os.path.join(REF_ROOT, f)) for f in FILES)
    if not src_ok:
        return False
    os.makedirs(DST_DIR, exist_ok=True)
    meta = {}
    for f in FILES:
        src = os.path.join(REF_ROOT, f)
        dst = os.path.join(DST_DIR, os.path.basename(f))
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            meta[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "STAGED.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "sha256": meta}, fh, indent=1)
    if verbose:
        print("staged", meta)
    return True


def load_reference_module():
    """The staged clip/loss.py as a module (None when it was never staged)."""
    path = os.path.join(DST_DIR, "loss.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("flyp_reference_clip_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else f"{REF_ROOT} not present: nothing staged")
