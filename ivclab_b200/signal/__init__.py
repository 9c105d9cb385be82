from .dct import DiscreteCosineTransform  # noqa: F401
from .color import luma8_from_rgb8, rgb2ycbcr, ycbcr2rgb  # noqa: F401

__all__ = ["DiscreteCosineTransform", "rgb2ycbcr", "ycbcr2rgb", "luma8_from_rgb8"]
