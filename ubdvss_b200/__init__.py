"""ubdvss_b200 -- the segment + connected-component hot path of asmekal/ubdvss on B200 (sm_100a).

Python mirrors of the reference's boundary (``NetConfig`` / ``NetManager`` / Keras-shaped model,
``ModelRunner.predict``, ``SegmapManager.postprocess``) over the C ABI of ``csrc/libubd.so``
(``include/ubd.h``).  There is no CPU fallback: importing the binding without the built library
raises ImportError, creating a handle without a B200 raises ``UbdError``."""
from .data_markup import ClassifiedObjectMarkup, ObjectMarkup  # noqa: F401

__all__ = ["ObjectMarkup", "ClassifiedObjectMarkup", "net", "model_runner", "segmap_manager", "utils", "losses"]
