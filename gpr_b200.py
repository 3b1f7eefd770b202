"""Import shim: the package directory is named ``gaussian-object-modelling_b200`` (with hyphens, as the
project layout requires), which Python cannot import by name.  ``import gpr_b200`` loads it."""
import importlib.util
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gaussian-object-modelling_b200")
_NAME = "gaussian_object_modelling_b200"
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG, "__init__.py"),
                                                   submodule_search_locations=[_PKG])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
workloads = importlib.import_module(_NAME + ".workloads")
