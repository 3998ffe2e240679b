"""Imports the package directory ``reinforcementlearningzoo.jl_b200`` (its name is not a valid Python identifier)
under the module name ``rlzoo_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "reinforcementlearningzoo.jl_b200")


def load():
    if "rlzoo_b200" in sys.modules:
        return sys.modules["rlzoo_b200"]
    spec = importlib.util.spec_from_file_location("rlzoo_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rlzoo_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
