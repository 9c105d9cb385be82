from .motion import MotionCompensator  # noqa: F401

__all__ = ["MotionCompensator"]
