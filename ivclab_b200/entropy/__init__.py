from .zerorun import ZeroRunCoder  # noqa: F401

__all__ = ["ZeroRunCoder"]
