from .entropy import stats_marg, symbol_histogram, symbol_minmax, zerorun_symbol_histogram  # noqa: F401
from .zerorun import ZeroRunCoder  # noqa: F401

__all__ = ["ZeroRunCoder", "stats_marg", "symbol_minmax", "symbol_histogram", "zerorun_symbol_histogram"]
