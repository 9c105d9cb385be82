from .patchquant import PatchQuant  # noqa: F401

__all__ = ["PatchQuant"]
