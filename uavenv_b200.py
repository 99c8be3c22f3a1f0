"""Import shim: the package directory `target-allocation-ppo-transformer_b200/` is not a valid Python
identifier, so it is registered here under the module name `target_allocation_ppo_transformer_b200`
and re-exported.  Usage:  import uavenv_b200 as ub;  env = ub.UAVEnvBatched(4096)."""
import importlib.util
import os
import sys

_NAME = "target_allocation_ppo_transformer_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "target-allocation-ppo-transformer_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
package = sys.modules[_NAME]

Config, cfg, HARD_MODE = package.Config, package.cfg, package.HARD_MODE


def __getattr__(name):
    return getattr(package, name)
