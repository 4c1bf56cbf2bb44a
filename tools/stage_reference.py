"""Stages the reference (read-only at /root/reference in the authoring container) into baseline/_ref/ so that it
travels to the GPU box with the gpurun snapshot (baseline/_ref/ is git-ignored, NOT gpurun-ignored).

    python tools/stage_reference.py [--src /root/reference]

Only the files the hot path imports are copied (Python sources + YAML configs + LICENSE); nothing is modified.  The
staged tree is used by
  * tools/gpu_reference_golden.py   — the reference itself on a B200 (CUDA autocast bf16 / fp32) -> tests/golden/full_*.pt
  * bench.py --impl reference       — the reference's own CPU path timed on the box's host cores
It is never imported by the product package.  Reference sources are NOT committed (see .gitignore).
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = {".py", ".yaml", ".yml", ".json", ".toml", ".txt", ".md"}


def stage(src, dst=DST):
    if not os.path.isdir(os.path.join(src, "hunyuanvideo_foley")):
        raise SystemExit(f"{src} does not look like the reference tree")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__", "example_workflows", ".github")]
        rel = os.path.relpath(base, src)
        for f in files:
            if os.path.splitext(f)[1].lower() in KEEP_EXT or f == "LICENSE":
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copy2(os.path.join(base, f), os.path.join(dst, rel, f))
                n += 1
    return n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("FOLEY_REFERENCE_SRC", "/root/reference"))
    a = ap.parse_args()
    print(f"staged {stage(a.src)} files into {DST}")
    sys.exit(0)
