from .dct import DiscreteCosineTransform  # noqa: F401
from .color import rgb2ycbcr, ycbcr2rgb  # noqa: F401

__all__ = ["DiscreteCosineTransform", "rgb2ycbcr", "ycbcr2rgb"]
