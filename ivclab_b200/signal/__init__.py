from .dct import DiscreteCosineTransform  # noqa: F401

__all__ = ["DiscreteCosineTransform"]
