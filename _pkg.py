"""Loader for the product package.  Its directory is `ecdna-evo_b200` (the name the layout asks
for), which is not a Python identifier, so it is registered under the module name `ecdna_evo_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if "ecdna_evo_b200" in sys.modules:
        return sys.modules["ecdna_evo_b200"]
    d = os.path.join(_ROOT, "ecdna-evo_b200")
    spec = importlib.util.spec_from_file_location("ecdna_evo_b200", os.path.join(d, "__init__.py"),
                                                  submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ecdna_evo_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
