"""Import alias: the package directory is named after the reference repository
(``metadata-augmented-unet-for-lst-ndvi_b200/``), which is not a valid Python identifier;
``import mau_b200`` loads it under this name."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "metadata-augmented-unet-for-lst-ndvi_b200")
_spec = _u.spec_from_file_location("mau_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["mau_b200"] = _mod
_spec.loader.exec_module(_mod)
