"""Import shim: makes the on-disk package directory `sketchlib.rust_b200/` importable as
`sketchlib.rust_b200` (a dot cannot appear in a Python package directory name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sketchlib.rust_b200")
_spec = importlib.util.spec_from_file_location(
    "sketchlib.rust_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
rust_b200 = importlib.util.module_from_spec(_spec)
sys.modules["sketchlib.rust_b200"] = rust_b200
_spec.loader.exec_module(rust_b200)
