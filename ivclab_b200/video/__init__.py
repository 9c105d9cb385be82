from .motion import MotionCompensator  # noqa: F401
from .closed_loop import ClosedLoopLumaCoder  # noqa: F401

__all__ = ["MotionCompensator", "ClosedLoopLumaCoder"]
