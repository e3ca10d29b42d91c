"""Import shim: the package directory is named `approximate-spmv-topk_b200` (hyphens), which Python
cannot import by name.  pkg() registers it as the module `approximate_spmv_topk_b200`."""
import importlib.util
import sys
from pathlib import Path

_NAME = "approximate_spmv_topk_b200"
_ROOT = Path(__file__).resolve().parent
_DIR = _ROOT / "approximate-spmv-topk_b200"


def pkg():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, _DIR / "__init__.py", submodule_search_locations=[str(_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
