"""Stage the UNMODIFIED reference Python sources of the hot path under ``baseline/_ref/``
(git-ignored, NOT gpurun-ignored) so that the reference's own eager-PyTorch path can be
timed on the GPU box (``bench.py --impl reference-gpu``), where /root/reference does not
exist.  TEST / BENCH INFRASTRUCTURE -- nothing under histogan_b200/ reads this directory.

    python -m oracle.make_baseline_ref

The reference has no setup.py / pyproject (pure scripts), so ``pip install --target`` has
nothing to build; the "install" is a verbatim copy of the three packages the training step
imports: histoGAN/, histogram_classes/, utils/*.py (byte-identical; checked below)."""
from __future__ import annotations

import filecmp
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("HISTOGAN_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("histoGAN", "histogram_classes", "utils")


def main() -> bool:
    if not os.path.isdir(os.path.join(SRC, "histoGAN")):
        return os.path.isdir(os.path.join(DST, "histoGAN"))
    for pkg in PACKAGES:
        for dirpath, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in files:
                if not f.endswith(".py"):
                    continue                     # skips utils/shape_predictor_68_face_landmarks.dat (96 MB)
                s = os.path.join(dirpath, f)
                d = os.path.join(DST, os.path.relpath(s, SRC))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
                assert filecmp.cmp(s, d, shallow=False)
    return True


if __name__ == "__main__":
    print("baseline/_ref staged:", main())
