"""Import alias: `import tneq_b200` loads the package that lives in
`quantum_circuits_symmetry_breaking_based_on_tneq-qc_b200/` (the directory name
mandated for this repository contains a hyphen, so it is not importable by name).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "quantum_circuits_symmetry_breaking_based_on_tneq-qc_b200")
_spec = _ilu.spec_from_file_location("tneq_b200", _os.path.join(_PKG_DIR, "__init__.py"),
                                     submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["tneq_b200"] = _mod
_spec.loader.exec_module(_mod)
