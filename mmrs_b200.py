"""Importable alias for the package directory
`multi-modal-retrieval-system-image-search-and-data-governance_b200/` (its name is not a valid
Python identifier).  `import mmrs_b200` loads that directory as the package `mmrs_b200`."""
import importlib.util as _ilu
import sys as _sys
from pathlib import Path as _Path

_PKG_DIR = _Path(__file__).resolve().parent / "multi-modal-retrieval-system-image-search-and-data-governance_b200"
_spec = _ilu.spec_from_file_location("mmrs_b200", _PKG_DIR / "__init__.py",
                                     submodule_search_locations=[str(_PKG_DIR)])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["mmrs_b200"] = _mod
_spec.loader.exec_module(_mod)
