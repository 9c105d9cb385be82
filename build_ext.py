"""Build ivclab_b200/_C/libivcb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch
extension machinery), and the oracle's C restatement with gcc.

    python build_ext.py [--force] [--verbose]

Lives outside the package on purpose: importing ``ivclab_b200`` requires the built library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "ivclab_b200")
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
OUT = os.path.join(OUT_DIR, "libivcb200.so")
SOURCES = ["ivc_abi.cu", "ivc_transform.cu", "ivc_motion.cu", "ivc_metrics.cu", "ivc_zerorun.cu", "ivc_color.cu"]
HEADERS = ["ivc_dct.cuh", "ivc_common.cuh", "ivc_color.cuh", "ivc_tile.cuh", os.path.join("..", "..", "include", "ivclab_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # belt and braces: the arithmetic also uses explicit _rn intrinsics
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps(src: str):
    return [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return any(_stale(OUT, _deps(s)) for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per translation unit, compiled in parallel and only when its source (or a header) changed,
    then one link step: a kernel edit costs the compile time of its own file."""
    if not force and not needs_build():
        return OUT
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or _stale(obj, _deps(src)):
            cmd = [_nvcc()] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, proc in jobs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed compiling {', '.join(failed)}")
    objs = [os.path.join(obj_dir, s.replace(".cu", ".o")) for s in SOURCES]
    res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                          "-o", OUT + ".tmp"] + objs, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed linking libivcb200.so")
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
