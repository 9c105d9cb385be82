from .shape import ZigZag, Patcher  # noqa: F401

__all__ = ["ZigZag", "Patcher"]
