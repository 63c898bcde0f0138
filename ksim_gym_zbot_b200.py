"""Import shim: the package directory is named ``ksim-gym-zbot_b200`` (not a Python identifier).

``import ksim_gym_zbot_b200`` executes this file, which loads that directory as the package of
the same dotted name, so ``ksim_gym_zbot_b200.engine`` etc. resolve inside it.
"""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ksim-gym-zbot_b200")
_spec = _u.spec_from_file_location(
    "ksim_gym_zbot_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = _u.module_from_spec(_spec)
_sys.modules["ksim_gym_zbot_b200"] = _mod
_spec.loader.exec_module(_mod)
