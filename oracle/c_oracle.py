"""ctypes wrapper of oracle/c/ivc_oracle.c (TEST INFRASTRUCTURE / CPU baseline).

Same results as ``ivc_oracle.py`` (checked bit-for-bit in tests/test_oracle_cpu.py), two to three
orders of magnitude faster on motion estimation, so full-size (1080p / 4K) parity checks and a
compiled multi-threaded CPU baseline are affordable.  ``threads`` fans block rows out over host
threads (ctypes releases the GIL)."""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libivc_oracle.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle/c` (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        p, i64, i = C.c_void_p, C.c_int64, C.c_int
        _lib.ivc_o_intra_forward.argtypes = [p, i64, i64, i, p, p, i64, i64]
        _lib.ivc_o_intra_inverse.argtypes = [p, i64, i64, i, p, p, i64, i64]
        _lib.ivc_o_me_f64.argtypes = [p, p, i64, i64, i, p, i64, i64]
        _lib.ivc_o_me_f32.argtypes = [p, p, i64, i64, i, p, i64, i64]
        _lib.ivc_o_mc_f64.argtypes = [p, i64, i64, i64, p, i, p]
        for f in ("ivc_o_intra_forward", "ivc_o_intra_inverse", "ivc_o_me_f64", "ivc_o_me_f32", "ivc_o_mc_f64"):
            getattr(_lib, f).restype = None
    return _lib


def _rows(n, threads):
    threads = max(1, min(threads, n))
    edges = np.linspace(0, n, threads + 1).astype(int)
    return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def _fan(fn, n_rows, threads):
    parts = _rows(n_rows, threads)
    if len(parts) == 1:
        fn(*parts[0])
        return
    with ThreadPoolExecutor(len(parts)) as ex:
        list(ex.map(lambda ab: fn(*ab), parts))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def intra_forward(img, table, threads=1):
    img = np.ascontiguousarray(img, dtype=np.float64)
    if img.ndim == 2:
        img = img[..., None]
    H, W, Cc = img.shape
    tab = np.ascontiguousarray(table, dtype=np.float64).reshape(3, 64)
    out = np.empty((H // 8, W // 8, 3, 64), dtype=np.int32)
    L = lib()
    _fan(lambda a, b: L.ivc_o_intra_forward(_ptr(img), H, W, Cc, _ptr(tab), _ptr(out), a, b), H // 8, threads)
    return out


def intra_inverse(zz, table, threads=1):
    zz = np.ascontiguousarray(zz, dtype=np.int32)
    Hp, Wp, Cc, _ = zz.shape
    tab = np.ascontiguousarray(table, dtype=np.float64).reshape(3, 64)
    out = np.empty((Hp * 8, Wp * 8, 3), dtype=np.float64)
    L = lib()
    _fan(lambda a, b: L.ivc_o_intra_inverse(_ptr(zz), Hp, Wp, Cc, _ptr(tab), _ptr(out), a, b), Hp, threads)
    return out


def me_full_search(ref, cur, sr, threads=1):
    f32 = ref.dtype == np.float32 and cur.dtype == np.float32
    dt = np.float32 if f32 else np.float64
    ref = np.ascontiguousarray(ref, dtype=dt)
    cur = np.ascontiguousarray(cur, dtype=dt)
    H, W = ref.shape
    mv = np.empty((H // 8, W // 8, 1), dtype=np.int64)
    L = lib()
    fn = L.ivc_o_me_f32 if f32 else L.ivc_o_me_f64
    _fan(lambda a, b: fn(_ptr(ref), _ptr(cur), H, W, int(sr), _ptr(mv), a, b), H // 8, threads)
    return mv


def mc_reconstruct(ref, mv, sr):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    H, W, Cc = ref.shape
    m = np.ascontiguousarray(np.asarray(mv)[:, :, 0], dtype=np.int64)
    out = np.empty_like(ref)
    lib().ivc_o_mc_f64(_ptr(ref), H, W, Cc, _ptr(m), int(sr), _ptr(out))
    return out
