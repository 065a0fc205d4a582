"""Import alias: the package directory is named ``modern-search-engines-project_b200`` (not a valid
identifier), so ``import mse_b200`` loads it through importlib and registers its submodules under
the ``mse_b200.*`` names as well."""
import importlib
import sys

_REAL = "modern-search-engines-project_b200"
_pkg = importlib.import_module(_REAL)
for _k, _m in list(sys.modules.items()):
    if _k.startswith(_REAL + "."):
        sys.modules["mse_b200" + _k[len(_REAL):]] = _m
sys.modules["mse_b200"] = _pkg
