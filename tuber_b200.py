"""Import shim: ``import tuber_b200`` loads the package directory ``tubelet-transformer_b200/``.

The package directory carries the upstream project's name, which is not a valid Python
identifier; this module replaces itself in ``sys.modules`` with that directory loaded as the
package ``tuber_b200`` so that ``import tuber_b200.models.tuber_ava`` etc. work.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "tubelet-transformer_b200")
_spec = _ilu.spec_from_file_location("tuber_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["tuber_b200"] = _mod
_spec.loader.exec_module(_mod)
