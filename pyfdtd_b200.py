"""Import shim: exposes the package directory ``py-fdtd_pic_b200/`` as the module ``pyfdtd_b200``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "py-fdtd_pic_b200")
_spec = importlib.util.spec_from_file_location(
    "pyfdtd_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pyfdtd_b200"] = _mod
_spec.loader.exec_module(_mod)
