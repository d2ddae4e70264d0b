"""Import shim: the package directory is named ``nabla.jl_b200`` (after the reference repository,
invenia/Nabla.jl), which is not a valid Python identifier.  ``import nabla_jl_b200`` loads that
directory as a regular package under this name (sub-modules import as ``nabla_jl_b200.core`` ...)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nabla.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "nabla_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nabla_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
