"""Importable alias for the package directory `sample-resilient-llm-inference_b200/`
(hyphens are not valid in a Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("sample-resilient-llm-inference_b200")
sys.modules[__name__] = _pkg
