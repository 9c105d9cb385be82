"""CPU oracle for the ivclab per-block coding loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline -- never on the path that is measured or shipped as the
GPU implementation.  ``ivclab_b200`` never imports this package.

Parity status: PINNED.  ``oracle/gen_golden.py`` executes the reference's own
modules (``/root/reference/ivclab/{signal/dct,quantization/patchquant,
utils/shape,video/motion}.py``, loaded by file path) on seeded inputs, checks
that every oracle function is bit-identical to them, and freezes the
input/output vectors under ``tests/golden/``.
"""
from .ivc_oracle import *  # noqa: F401,F403
