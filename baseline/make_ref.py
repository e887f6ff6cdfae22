#!/usr/bin/env python
"""Stage the UNMODIFIED reference for ``bench.py --impl reference`` under ``baseline/_ref/`` (git-ignored; it
travels to the GPU box with the gpurun snapshot like the built .so files do).

The reference (Peterande/SAST) has no setup.py / pyproject: ``pip install /root/reference`` has nothing to build,
so the "install" is a verbatim copy of the packages its backbone path imports -- ``models/``, ``data/utils/types.py``, ``data/genx_utils/labels.py``,
``utils/{timers,padding,helpers}.py`` -- plus the 40-line ``omegaconf`` stand-in of ``oracle/_shim`` (omegaconf,
hydra, pytorch_lightning and fvcore are absent from the image, so ``benchmark.py`` itself cannot be imported;
``bench.py`` restates its timing loop, benchmark.py:33-42, around ``build_recurrent_backbone(cfg).forward``).
Run in the build container:   python baseline/make_ref.py
Nothing under baseline/_ref is product code; no file of it is tracked by git."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SAST_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not present: baseline/_ref left as it is", file=sys.stderr)
        return 0 if os.path.isdir(DST) else 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "models"), os.path.join(DST, "models"))
    for rel in ("data/utils/types.py", "data/genx_utils/labels.py", "utils/timers.py", "utils/padding.py",
                "utils/helpers.py", "LICENSE"):
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copy2(os.path.join(SRC, rel), os.path.join(DST, rel))
    shutil.copytree(os.path.join(ROOT, "oracle", "_shim", "omegaconf"), os.path.join(DST, "omegaconf"))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"staged {n} files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
