"""B200-native request router + model replicas (drop-in for the litellm.Router hot path of
aws-samples/sample-resilient-llm-inference).  See DESIGN.md.

Importing this package loads librr_b200.so; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA extension is missing)
from ._lib import RRError, lib  # noqa: F401

__all__ = ["lib", "RRError"]
