"""Import helper: the package directory `zkinterface-ir_b200/` has a hyphen in its
name, so it is loaded by path and registered as module `zkir_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "zkinterface-ir_b200")


def load():
    if "zkir_b200" in sys.modules:
        return sys.modules["zkir_b200"]
    spec = importlib.util.spec_from_file_location("zkir_b200", os.path.join(_PKG, "__init__.py"),
                                                  submodule_search_locations=[_PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["zkir_b200"] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules["zkir_b200"]
        raise
    return mod
