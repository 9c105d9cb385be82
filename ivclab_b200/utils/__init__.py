from .shape import ZigZag, Patcher  # noqa: F401
from .metrics import calc_mse, calc_psnr, frame_sse, frame_sse_rgb8_vs_ycbcr  # noqa: F401

__all__ = ["ZigZag", "Patcher", "calc_mse", "calc_psnr", "frame_sse", "frame_sse_rgb8_vs_ycbcr"]
