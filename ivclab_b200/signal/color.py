"""``ivclab.signal.rgb2ycbcr`` / ``ycbcr2rgb`` on the B200 (reference: ivclab/signal/color.py:15-63;
SURVEY.md section 8f row N1).  Bit-identical to numpy: the forward transform replays the FMA chain of
numpy's BLAS matmul, the inverse is elementwise with individually rounded operations."""
from __future__ import annotations

import torch

from .. import _lib
from .._runtime import code, dev_index, stream_ptr, to_device, to_host

__all__ = ["rgb2ycbcr", "ycbcr2rgb", "luma8_from_rgb8"]


def rgb2ycbcr(image):
    """[..., 3] RGB (uint8 / int32 / float32 / float64) -> float64 YCbCr of the same shape."""
    t, was_np = to_device(image)
    if t.ndim < 1 or t.shape[-1] != 3:
        raise ValueError(f"matmul: input operand has a mismatch in its core dimension: expected [..., 3], got {tuple(t.shape)}")
    if t.dtype not in (torch.uint8, torch.int32, torch.float32, torch.float64):
        t = t.to(torch.float64)
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.float64, device=t.device)
    st = _lib.lib.ivc_rgb2ycbcr(dev_index(t), stream_ptr(t.device), t.data_ptr(), code(t.dtype), t.numel() // 3, out.data_ptr())
    _lib.check(st, "ivc_rgb2ycbcr")
    return to_host(out, was_np)


def ycbcr2rgb(image):
    """[H, W, 3] (or [..., 3]) YCbCr -> float64 RGB clipped to [0, 255]."""
    t, was_np = to_device(image)
    if t.ndim < 1 or t.shape[-1] != 3:
        raise ValueError(f"expected [..., 3], got {tuple(t.shape)}")
    t = t.to(torch.float64).contiguous()
    out = torch.empty_like(t)
    st = _lib.lib.ivc_ycbcr2rgb(dev_index(t), stream_ptr(t.device), t.data_ptr(), t.numel() // 3, out.data_ptr())
    _lib.check(st, "ivc_ycbcr2rgb")
    return to_host(out, was_np)


def luma8_from_rgb8(image, out=None, out_f64=None):
    """uint8 RGB [..., 3] -> uint8 luma plane ``clip(round(rgb2ycbcr(image)[..., 0]), 0, 255)`` (np.round: half to even):
    the Y channel the video codecs code (ivclab/video/videocodec.py:38) in 8-bit form, computed on the device.
    ``out`` / ``out_f64``: caller-owned device tensors for the plane and (optionally) its float64 copy."""
    t, was_np = to_device(image)
    if t.ndim < 1 or t.shape[-1] != 3 or t.dtype != torch.uint8:
        raise ValueError(f"expected uint8 [..., 3], got {t.dtype} {tuple(t.shape)}")
    t = t.contiguous()
    npix = t.numel() // 3
    if out is None:
        out = torch.empty(t.shape[:-1], dtype=torch.uint8, device=t.device)
    for o, dt in ((out, torch.uint8), (out_f64, torch.float64)):
        if o is not None and (o.dtype != dt or o.numel() != npix or not o.is_contiguous() or o.device != t.device):
            raise ValueError("out / out_f64 must be contiguous uint8 / float64 tensors with one element per pixel on the input's device")
    st = _lib.lib.ivc_rgb8_to_luma8(dev_index(t), stream_ptr(t.device), t.data_ptr(), npix, out.data_ptr(),
                                    out_f64.data_ptr() if out_f64 is not None else None)
    _lib.check(st, "ivc_rgb8_to_luma8")
    return to_host(out, was_np)
