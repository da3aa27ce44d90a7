"""Import shim: ``import b200ssl`` loads the package that lives in ``gipmed-project-self-supervised-vit_b200/``
(a directory name Python's import statement cannot spell)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gipmed-project-self-supervised-vit_b200")
_spec = importlib.util.spec_from_file_location("b200ssl", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200ssl"] = _mod
_spec.loader.exec_module(_mod)
