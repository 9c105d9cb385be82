from .intracodec import IntraCodec  # noqa: F401

__all__ = ["IntraCodec"]
