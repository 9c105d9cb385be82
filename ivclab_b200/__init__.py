"""ivclab_b200 -- B200-native (sm_100a) implementation of ivclab's per-block coding loop.

Drop-in classes with the reference's names, signatures and array conventions:

    DiscreteCosineTransform   (ivclab.signal)         transform / inverse_transform
    PatchQuant                (ivclab.quantization)   get_quantization_table / quantize / dequantize
    ZigZag, Patcher           (ivclab.utils)          flatten / unflatten, patch / unpatch
    MotionCompensator         (ivclab.video)          compute_motion_vector / reconstruct_with_motion_vector

plus the fused coders (:class:`IntraBlockCoder`, :class:`PFrameBlockCoder`), the sharding
helpers (:mod:`ivclab_b200.shard`) and :func:`install` / :func:`inject` to run the unmodified
reference codecs on top.  Everything executes in hand-written CUDA kernels behind the C ABI of
``include/ivclab_b200.h``; there is no CPU fallback -- importing this package fails if the
extension has not been built, and calling it fails if no CUDA device is visible.
"""
from . import _lib  # noqa: F401  (loads libivcb200.so or raises)
from .codec import IntraBlockCoder, PFrameBlockCoder, forward_rgb_multi
from .entropy import ZeroRunCoder, stats_marg, symbol_histogram, symbol_minmax, zerorun_symbol_histogram
from .image import IntraCodec
from ._runtime import set_device_memo
from .install import inject, install
from .quantization import PatchQuant
from .signal import DiscreteCosineTransform, luma8_from_rgb8, rgb2ycbcr, ycbcr2rgb
from .streaming import StreamedCoder
from .sweep import RateDistortionSweep
from .utils import Patcher, ZigZag, calc_mse, calc_psnr, frame_sse, frame_sse_rgb8_vs_ycbcr
from .video import ClosedLoopLumaCoder, MotionCompensator

__version__ = "0.1.0"
__all__ = ["DiscreteCosineTransform", "PatchQuant", "ZigZag", "Patcher", "MotionCompensator",
           "IntraBlockCoder", "PFrameBlockCoder", "forward_rgb_multi", "IntraCodec", "ClosedLoopLumaCoder", "ZeroRunCoder", "calc_mse", "calc_psnr",
           "frame_sse", "frame_sse_rgb8_vs_ycbcr", "rgb2ycbcr", "ycbcr2rgb", "luma8_from_rgb8", "StreamedCoder", "stats_marg", "symbol_minmax", "symbol_histogram", "zerorun_symbol_histogram",
           "install", "inject", "set_device_memo", "RateDistortionSweep"]
