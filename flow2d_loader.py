"""Imports the product package, whose directory name (`cuda-flow2d_b200`) is not a Python identifier."""
import importlib.util
import os
import sys

_NAME = "cuda_flow2d_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-flow2d_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(root, "__init__.py"), submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
