"""Import alias: ``import snsde_b200`` loads the package in ``stable-neural-sdes_b200/``
(the directory name required by the repo layout is not a valid Python identifier)."""
import importlib.util
import pathlib
import sys

_dir = pathlib.Path(__file__).resolve().parent / "stable-neural-sdes_b200"
_spec = importlib.util.spec_from_file_location(
    "snsde_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["snsde_b200"] = _mod
_spec.loader.exec_module(_mod)
