"""Swap the B200 classes into the reference package so that the UNMODIFIED ``IntraCodec`` and
the ch4 exercise codecs run on them.

``ivclab.image.intracodec`` binds ``DiscreteCosineTransform``, ``PatchQuant``, ``ZigZag`` and
``Patcher`` by name at import time (intracodec.py:3-5), so either call :func:`install` before
importing it, or use :func:`inject` on an existing codec instance."""
from __future__ import annotations

import importlib
import sys

from .quantization import PatchQuant
from .signal import DiscreteCosineTransform
from .utils import Patcher, ZigZag
from .video import MotionCompensator

_TARGETS = {
    "ivclab.signal": {"DiscreteCosineTransform": DiscreteCosineTransform},
    "ivclab.signal.dct": {"DiscreteCosineTransform": DiscreteCosineTransform},
    "ivclab.quantization": {"PatchQuant": PatchQuant},
    "ivclab.quantization.patchquant": {"PatchQuant": PatchQuant},
    "ivclab.utils": {"ZigZag": ZigZag, "Patcher": Patcher},
    "ivclab.utils.shape": {"ZigZag": ZigZag, "Patcher": Patcher},
    "ivclab.video": {"MotionCompensator": MotionCompensator},
    "ivclab.video.motion": {"MotionCompensator": MotionCompensator},
    # modules that copied the names at import time
    "ivclab.image.intracodec": {"DiscreteCosineTransform": DiscreteCosineTransform, "PatchQuant": PatchQuant,
                                "ZigZag": ZigZag, "Patcher": Patcher},
    "ivclab.video.videocodec": {"MotionCompensator": MotionCompensator},
}


def install(import_missing: bool = False) -> list:
    """Replace the five classes in every already-imported ``ivclab`` module (and, with
    ``import_missing=True``, import the leaf modules first).  Returns the patched module names."""
    done = []
    for modname, names in _TARGETS.items():
        mod = sys.modules.get(modname)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                mod = None
        if mod is None:
            continue
        for k, v in names.items():
            setattr(mod, k, v)
        done.append(modname)
    return done


def inject(codec):
    """Attribute-inject into an existing reference codec object: an ``IntraCodec`` gets
    ``dct/quant/zigzag/patcher``; a ``VideoCodec``-like object gets ``motion_comp`` and its
    ``intra_codec`` / ``residual_codec`` members are injected recursively."""
    if hasattr(codec, "dct") and hasattr(codec, "quant"):
        q = codec.quant
        codec.dct = DiscreteCosineTransform(getattr(codec.dct, "norm", "ortho"))
        codec.quant = PatchQuant(getattr(q, "quantization_scale", 1.0), getattr(q, "luminance", None),
                                 getattr(q, "chrominance", None))
        codec.zigzag = ZigZag()
        if hasattr(codec, "patcher"):
            codec.patcher = Patcher(getattr(codec.patcher, "window_size", (8, 8)))
    if hasattr(codec, "motion_comp"):
        codec.motion_comp = MotionCompensator(getattr(codec.motion_comp, "search_range", 4))
    for name in ("intra_codec", "residual_codec"):
        if hasattr(codec, name):
            inject(getattr(codec, name))
    return codec
