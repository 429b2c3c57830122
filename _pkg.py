"""Loads the package directory `hm-opencl_b200/` (hyphenated: not a valid identifier) as module
`hm_opencl_b200`.   Usage:  from _pkg import hm"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_ROOT, "hm-opencl_b200")


def load():
    if "hm_opencl_b200" in sys.modules:
        return sys.modules["hm_opencl_b200"]
    spec = importlib.util.spec_from_file_location("hm_opencl_b200", os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hm_opencl_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


hm = load()
