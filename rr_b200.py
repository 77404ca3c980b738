"""Importable alias for the package directory `sample-resilient-llm-inference_b200/` (hyphens are not valid in a
Python identifier).  `import rr_b200`, `from rr_b200.router import Router`, `import rr_b200.server` all resolve to
the ONE set of module objects of the real package (no duplicate classes)."""
import importlib
import importlib.abc
import importlib.machinery
import sys

_REAL = "sample-resilient-llm-inference_b200"
_ALIAS = __name__


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_ALIAS + "."):
            return None
        return importlib.machinery.ModuleSpec(fullname, self)

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


_pkg = importlib.import_module(_REAL)
sys.meta_path.insert(0, _AliasFinder())
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[_ALIAS + _name[len(_REAL):]] = _mod
sys.modules[_ALIAS] = _pkg
