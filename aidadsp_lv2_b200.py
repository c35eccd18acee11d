"""Import shim: the package directory is named ``aidadsp-lv2_b200`` (hyphen, as the project is),
which Python cannot import by name.  ``import aidadsp_lv2_b200`` loads it from that directory."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "aidadsp-lv2_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
