"""Import alias for the product package.

The package directory is named ``diffab-pytorch_b200/`` (hyphen, as the repo layout requires),
which Python cannot import by name.  Importing this module registers that directory as the
package ``diffab_pytorch_b200`` so that ``import diffab_pytorch_b200.so3`` etc. work from the
repo root (tests, bench.py and __graft_entry__ put the repo root on ``sys.path``).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffab-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "diffab_pytorch_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_module = importlib.util.module_from_spec(_spec)
sys.modules["diffab_pytorch_b200"] = _module
_spec.loader.exec_module(_module)
