"""``ivclab.utils.calc_mse`` / ``calc_psnr`` on the B200 (reference: ivclab/utils/metrics.py:3-40;
SURVEY.md section 8f row N3), plus the batched per-frame form the RD sweeps need."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib
from .._runtime import code, dev_index, stream_ptr, to_device

__all__ = ["calc_mse", "calc_psnr", "frame_sse", "frame_sse_rgb8_vs_ycbcr"]


def frame_sse(orig, rec):
    """Per-unit sum of squared differences: ``orig``/``rec`` are ``[N, ...]`` (same shape, or gray
    ``[N,H,W]`` against RGB ``[N,H,W,3]``) -> float64 CUDA tensor ``[N]``.  One device pass over both
    arrays, deterministic, no host synchronisation."""
    a, _ = to_device(orig)
    b, _ = to_device(rec, a.device)
    if a.ndim < 1 or b.ndim < 1 or a.shape[0] != b.shape[0]:
        raise ValueError(f"leading (unit) axes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.ndim == b.ndim - 1 and tuple(b.shape[:-1]) == tuple(a.shape) and b.shape[-1] == 3:
        bc = 3
    elif b.ndim == a.ndim - 1 and tuple(a.shape[:-1]) == tuple(b.shape) and a.shape[-1] == 3:
        a, b, bc = b, a, 3                      # squared difference is symmetric
    elif tuple(a.shape) == tuple(b.shape):
        bc = 1
    else:
        raise AssertionError(f"Image shapes don't match after processing: {tuple(a.shape)} vs {tuple(b.shape)}")
    fix = lambda t: t.to(torch.float64) if t.dtype not in (torch.uint8, torch.int32, torch.float32, torch.float64, torch.int64) else t
    a, b = fix(a).contiguous(), fix(b).contiguous()
    n = b.shape[0]
    unit = b.numel() // n if n else 0
    out = torch.empty(n, dtype=torch.float64, device=b.device)
    wsb = _lib.lib.ivc_sse_workspace_bytes(n, unit)
    ws = torch.empty(max(wsb, 8), dtype=torch.uint8, device=b.device)
    st = _lib.lib.ivc_sum_squared_error(dev_index(b), stream_ptr(b.device), a.data_ptr(), code(a.dtype), b.data_ptr(),
                                        code(b.dtype), n, unit, bc, ws.data_ptr(), ws.numel(), out.data_ptr())
    _lib.check(st, "ivc_sum_squared_error")
    return out


def frame_sse_rgb8_vs_ycbcr(rgb8, rec_ycbcr):
    """``frame_sse(rgb2ycbcr(rgb8), rec_ycbcr)`` without materialising the float64 YCbCr original: uint8 RGB
    ``[N,H,W,3]`` against float64 YCbCr of the same shape -> float64 ``[N]``, the same bits as the two-step form."""
    a, _ = to_device(rgb8)
    b, _ = to_device(rec_ycbcr, a.device)
    if a.dtype != torch.uint8 or b.dtype != torch.float64 or tuple(a.shape) != tuple(b.shape) or a.shape[-1] != 3:
        raise ValueError(f"expected uint8 RGB and float64 YCbCr of one shape [N,...,3], got {a.dtype} {tuple(a.shape)} / {b.dtype} {tuple(b.shape)}")
    a, b = a.contiguous(), b.contiguous()
    n = b.shape[0]
    unit = b.numel() // n if n else 0
    out = torch.empty(n, dtype=torch.float64, device=b.device)
    wsb = _lib.lib.ivc_sse_workspace_bytes(n, unit)
    ws = torch.empty(max(wsb, 8), dtype=torch.uint8, device=b.device)
    st = _lib.lib.ivc_sum_squared_error(dev_index(b), stream_ptr(b.device), a.data_ptr(), _lib.U8, b.data_ptr(), _lib.F64,
                                        n, unit, _lib.SSE_RGB8_AS_YCBCR, ws.data_ptr(), ws.numel(), out.data_ptr())
    _lib.check(st, "ivc_sum_squared_error")
    return out


def calc_mse(orig, rec):
    """Mean squared error over all samples, float64 (metrics.py:3-23): a gray image is compared with
    every channel of an RGB one."""
    a = orig if isinstance(orig, torch.Tensor) else np.asarray(orig)
    b = rec if isinstance(rec, torch.Tensor) else np.asarray(rec)
    sse = frame_sse(a[None], b[None])
    n = max(int(np.prod(a.shape)), int(np.prod(b.shape)))
    return float(sse.item()) / n


def calc_psnr(orig, rec, maxval=255):
    """``20*log10(maxval / sqrt(mse))`` (metrics.py:25-40)."""
    mse = calc_mse(orig, rec)
    return 20 * math.log10(maxval / math.sqrt(mse)) if mse > 0 else float("inf")
