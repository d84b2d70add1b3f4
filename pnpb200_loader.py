"""Import helper: the package directory is named `modular-pnp_b200` (not a valid Python
identifier), so it is loaded by path under the module name `modular_pnp_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    name = "modular_pnp_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(_ROOT, "modular-pnp_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
