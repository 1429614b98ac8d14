"""Import shim: makes the in-tree directory `tensorqec.jl_b200/` importable as `tensorqec.jl_b200`.

The product package lives in the directory literally named `tensorqec.jl_b200` at the repo root; a dot is not legal
in a Python package name, so this one-file namespace registers that directory as the submodule `jl_b200`.
"""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tensorqec.jl_b200")
if "tensorqec.jl_b200" not in _sys.modules:
    _spec = _u.spec_from_file_location("tensorqec.jl_b200", _os.path.join(_dir, "__init__.py"),
                                       submodule_search_locations=[_dir])
    jl_b200 = _u.module_from_spec(_spec)
    _sys.modules["tensorqec.jl_b200"] = jl_b200
    _spec.loader.exec_module(jl_b200)
else:
    jl_b200 = _sys.modules["tensorqec.jl_b200"]
