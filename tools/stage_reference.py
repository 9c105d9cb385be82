#!/usr/bin/env python
"""Stage the reference's python package next to the repo for ONE GPU-box run of the real-caller test
(tests/test_gpu_round2.py::test_unmodified_reference_intracodec_on_installed_classes): copies
$IVCLAB_REFERENCE/ivclab (default /root/reference) into baseline/_ref/ivclab, which is git-ignored (never committed)
but travels with a gpurun snapshot.  `--clean` removes it again."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "baseline", "_ref")
if "--clean" in sys.argv:
    shutil.rmtree(os.path.join(ROOT, "baseline"), ignore_errors=True)
    print("removed", dst)
else:
    src = os.path.join(os.environ.get("IVCLAB_REFERENCE", "/root/reference"), "ivclab")
    shutil.rmtree(dst, ignore_errors=True)
    shutil.copytree(src, os.path.join(dst, "ivclab"), ignore=shutil.ignore_patterns("__pycache__"))
    print("staged", src, "->", dst)
