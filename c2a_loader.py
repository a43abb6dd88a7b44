"""Import helper: the package directory is `circom-2-arithc_b200/` (hyphen, as the task names it), which is not
a valid Python identifier, so load it under the module name `circom_2_arithc_b200`."""
import importlib.util
import os
import sys

_NAME = "circom_2_arithc_b200"
_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "circom-2-arithc_b200")


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


c2a = _load()
