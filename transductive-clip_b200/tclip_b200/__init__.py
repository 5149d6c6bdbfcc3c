"""tclip_b200 — B200-native (sm_100a) implementation of transductive-CLIP's batched EM inference loop.

Host side mirrors the reference's ``src/methods`` method-class API; all numeric work runs in
``libtclip_b200.so`` (hand-written CUDA, C ABI in ``include/tclip_b200.h``).  No CPU fallback."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
