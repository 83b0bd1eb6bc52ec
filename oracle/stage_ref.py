#!/usr/bin/env python
"""Stage the UNMODIFIED reference checkout under ``oracle/_ref/`` so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (like everything under ``oracle/``).  ``/root/reference`` exists only in the build
container; ``gpurun`` ships ``/root/repo`` minus what ``.gpurunignore`` lists, and ``oracle/_ref/`` is git-ignored but
NOT gpurun-ignored.  Nothing is copied into the history: this script only mirrors the reference's own files (its
``tblup`` package and ``main.py``) into the ignored directory, byte for byte, and records their sha256 so a reader can
check that the staged tree is the reference and nothing else.

Used by: ``bench.py --impl reference`` / ``cpu_baseline`` (the reference's own BlupParallelEvaluator with its worker
pool, tblup/evaluator.py:116-131,380-405), ``tests/test_gpu_main_loop.py`` and ``scripts/main_c1.py`` (the reference's
``main.py`` driving the GPU evaluator).  Run by ``__graft_entry__.build()`` whenever the reference is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def stage(src="/root/reference", dst=DST):
    if not os.path.isdir(os.path.join(src, "tblup")):
        return False
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for rel_root in ("tblup",):
        for root, dirs, files in os.walk(os.path.join(src, rel_root)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if not f.endswith(".py"):
                    continue
                s = os.path.join(root, f)
                rel = os.path.relpath(s, src)
                d = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    for f in ("main.py",):
        s = os.path.join(src, f)
        if os.path.isfile(s):
            shutil.copyfile(s, os.path.join(dst, f))
            manifest[f] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": manifest}, fh, indent=1, sort_keys=True)
    return True


def ref_path():
    """Directory to put on sys.path to import the reference's ``tblup`` package: the live checkout when present
    (build container), else the staged mirror (GPU box), else None."""
    env = os.environ.get("TBLUP_REFERENCE")
    for cand in (env, "/root/reference", DST):
        if cand and os.path.isdir(os.path.join(cand, "tblup")):
            return cand
    return None


if __name__ == "__main__":
    ok = stage(*(sys.argv[1:2] or ["/root/reference"]))
    print("staged" if ok else "reference not found; nothing staged", DST)
