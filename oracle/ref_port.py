"""CPU port that makes the SAME library calls as the reference (TEST INFRASTRUCTURE / CPU baseline).

``ivc_oracle.py`` spells the arithmetic out operation by operation (so that it doubles as the
kernel specification); this module instead restates each reference method with the third-party
calls the reference itself makes -- ``scipy.fft.dct/idct`` (ivclab/signal/dct.py:24,26,42,44),
numpy broadcasting / ``np.round`` / ``astype`` (ivclab/quantization/patchquant.py:59-60,77-78),
fancy-index scatter/gather (ivclab/utils/shape.py:26,32) and the four nested Python loops of
``compute_motion_vector`` (ivclab/video/motion.py:28-57).  Its run time is therefore the
reference's run time, which is what ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
measure on the GPU box (where /root/reference does not exist).  tests/test_oracle_cpu.py checks
it bit-for-bit against ``ivc_oracle`` (itself pinned against the real reference)."""
from __future__ import annotations

import numpy as np
from scipy.fft import dct as _dct, idct as _idct

from .ivc_oracle import ZIGZAG_ORDER, quant_table


def transform(p):                      # dct.py:12-28
    return _dct(_dct(p, axis=-1, norm="ortho"), axis=-2, norm="ortho")


def inverse_transform(c):              # dct.py:30-46
    return _idct(_idct(c, axis=-1, norm="ortho"), axis=-2, norm="ortho")


def quantize(x, table):                # patchquant.py:56-60
    return np.round(x / table[None, None]).astype(np.int32)


def dequantize(q, table):              # patchquant.py:74-78
    return (q * table[None, None]).astype(np.int32)


def flatten(p):                        # shape.py:21-28
    f = p.reshape(p.shape[:3] + (64,))
    out = np.zeros_like(f)
    out[:, :, :, ZIGZAG_ORDER] = f
    return out


def unflatten(z):                      # shape.py:30-36
    return z[:, :, :, ZIGZAG_ORDER].reshape(z.shape[:3] + (8, 8))


def patch(img):                        # shape.py:45-54 (strided view, like einops)
    H, W, C = img.shape
    return img.reshape(H // 8, 8, W // 8, 8, C).transpose(0, 2, 4, 1, 3)


def unpatch(p):                        # shape.py:56-65
    h, w, c = p.shape[:3]
    return np.ascontiguousarray(p.transpose(0, 3, 1, 4, 2)).reshape(h * 8, w * 8, c)


def compute_motion_vector(ref, cur, sr):          # motion.py:8-58, same loop nest
    H, W = ref.shape
    span = 2 * sr + 1
    mv = np.zeros((H // 8, W // 8, 1), dtype=int)
    for y in range(0, H, 8):
        for x in range(0, W, 8):
            blk = cur[y:y + 8, x:x + 8]
            lo, bdy, bdx = float("inf"), 0, 0
            for dy in range(-sr, sr + 1):
                for dx in range(-sr, sr + 1):
                    yy, xx = y + dy, x + dx
                    if yy < 0 or yy + 8 > H or xx < 0 or xx + 8 > W:
                        continue
                    s = np.sum((blk - ref[yy:yy + 8, xx:xx + 8]) ** 2)
                    if s < lo:
                        lo, bdy, bdx = s, dy, dx
            mv[y // 8, x // 8, 0] = (bdy + sr) * span + (bdx + sr)
    return mv


def reconstruct_with_motion_vector(ref, mv, sr):  # motion.py:60-97
    H, W, _ = ref.shape
    span = 2 * sr + 1
    out = np.zeros_like(ref)
    for y in range(0, H, 8):
        for x in range(0, W, 8):
            idx = mv[y // 8, x // 8, 0]
            yy, xx = y + idx // span - sr, x + idx % span - sr
            if yy < 0 or yy + 8 > H or xx < 0 or xx + 8 > W:
                continue
            out[y:y + 8, x:x + 8, :] = ref[yy:yy + 8, xx:xx + 8, :]
    return out


# ---- the coding-loop step that bench.py times (one frame) -------------------------------------
def intra_loop(img_hwc, table):
    """transform -> quantize -> flatten -> unflatten -> dequantize -> inverse_transform
    (intracodec.py:66-75 then :115-124), returns (scan indices, reconstruction)."""
    zz = flatten(quantize(transform(patch(img_hwc)), table))
    rec = unpatch(inverse_transform(dequantize(unflatten(zz), table)))
    return zz, rec


def pframe_loop(cur, ref, sr, table, mv=None):
    """ME -> MC -> residual -> transform path -> recon (videocodec.py:52-75 / E4-1.py:257-306)."""
    if mv is None:
        mv = compute_motion_vector(ref, cur, sr)
    pred = reconstruct_with_motion_vector(ref[..., None], mv, sr)[..., 0]
    resid = cur - pred
    zz = flatten(quantize(transform(patch(resid[..., None])), table))
    rec = unpatch(inverse_transform(dequantize(unflatten(zz[:, :, :1]), table)))[..., 0]
    return mv, zz, pred + rec


__all__ = ["transform", "inverse_transform", "quantize", "dequantize", "flatten", "unflatten", "patch", "unpatch",
           "compute_motion_vector", "reconstruct_with_motion_vector", "intra_loop", "pframe_loop", "quant_table"]
