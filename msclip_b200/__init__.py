"""msclip_b200 — B200-native MS-CLIP-S encode-and-contrast path behind the reference's model API."""
from .config import MSCLIPConfig, from_reference_config  # noqa: F401

__all__ = ["MSCLIPConfig", "from_reference_config", "CLIP", "get_clip_model"]


def __getattr__(name):
    # torch and the shared library are only needed once a model is built
    if name in ("CLIP", "get_clip_model"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
