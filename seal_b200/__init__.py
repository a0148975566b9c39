"""seal_b200 — B200-native constrained beam-search decode for SEAL.

Mirrors ``seal/__init__.py:7-9`` for the hot path: ``FMIndex``, ``fm_index_generate``,
``IndexBasedLogitsProcessor``.  Importing this package loads libsealb200.so (CUDA, sm_100a); there
is no CPU fallback.
"""
from .index import FMIndex  # noqa: F401
from .beam_search import fm_index_generate, IndexBasedLogitsProcessor, SealBartEngine  # noqa: F401

__all__ = ["FMIndex", "fm_index_generate", "IndexBasedLogitsProcessor", "SealBartEngine"]
