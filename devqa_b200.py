"""Import alias: ``import devqa_b200`` loads the package that lives in ``de-vqa_b200/``.

The task names the package directory ``de-vqa_b200`` (a hyphen is not importable), so this
module loads that directory under the importable name ``devqa_b200`` and replaces itself in
``sys.modules`` with it.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "de-vqa_b200")
_spec = _ilu.spec_from_file_location("devqa_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["devqa_b200"] = _mod
_spec.loader.exec_module(_mod)
