"""Import shim: the package directory is named `cleanrl.jl_b200/` (not a legal dotted
module name), so `import cleanrl_jl_b200` loads that directory as a package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cleanrl.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "cleanrl_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cleanrl_jl_b200"] = _mod
_spec.loader.exec_module(_mod)

if __name__ == "__main__":  # `python -m cleanrl_jl_b200 ppo --num_envs 4096 ...` (see cleanrl.jl_b200/cli.py)
    from cleanrl_jl_b200 import cli
    sys.exit(cli.main())
