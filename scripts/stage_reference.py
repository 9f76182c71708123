#!/usr/bin/env python3
"""Stage the reference's Python tree where the GPU box can see it: /root/reference -> baseline/_ref/.

`baseline/_ref/` is git-ignored (no reference source enters the history) but travels with `gpurun`, so that
tests/test_runners_gpu.py can execute the reference's UNMODIFIED runner files (Louvre_Evacuation/runners/*.py) on the
B200 classes through `python -m dqn_marl_b200.compat`.  Nothing in the product imports from there.

    python scripts/stage_reference.py [--src /root/reference]
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("MARL_REFERENCE_ROOT", "/root/reference"))
    args = ap.parse_args()
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(args.src, "Louvre_Evacuation")):
        raise SystemExit(f"{args.src} holds no Louvre_Evacuation/ tree")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "dqn_results")
    for name in ("Louvre_Evacuation", "configs"):
        shutil.copytree(os.path.join(args.src, name), os.path.join(dst, name), ignore=ignore)
    n = sum(len(f) for _, _, f in os.walk(dst))
    print(f"staged {n} files under {dst}")


if __name__ == "__main__":
    main()
