"""Register this package under the reference's module names.

After `install()`, the reference's own `from models.aggregation import
build_volume_generator`, `from utils import volumetric, multiview` resolve to
the B200 implementations; nothing else of the reference changes.  Only the
three hot-path modules are replaced — other `models.*` / `utils.*` modules
keep resolving to the reference tree if it is on sys.path.
"""
import importlib
import sys
import types

_MAP = {
    "models.aggregation": "multiviewhmr_b200.aggregation",
    "utils.volumetric": "multiviewhmr_b200.volumetric",
    "utils.multiview": "multiviewhmr_b200.multiview",
}


def install():
    for ref_name, ours in _MAP.items():
        mod = importlib.import_module(ours)
        pkg_name, leaf = ref_name.split(".")
        try:
            pkg = importlib.import_module(pkg_name)
        except ImportError:
            pkg = types.ModuleType(pkg_name)
            pkg.__path__ = []
            sys.modules[pkg_name] = pkg
        sys.modules[ref_name] = mod
        setattr(pkg, leaf, mod)
    return sorted(_MAP)


def uninstall():
    for ref_name in _MAP:
        sys.modules.pop(ref_name, None)
