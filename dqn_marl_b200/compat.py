"""Run the reference's runners UNMODIFIED on the B200 classes.

The reference's runners import their environment and agent by absolute module path
(`from Louvre_Evacuation.envs.evacuation_env import EvacuationEnv`, `...envs.evacuation_env_multi import
EvacuationEnvMulti`, `...agents.dqn_agent import DQNAgent`; runners/train_dqn.py:19-22, train_double_dqn.py:12-13,
evaluate_strategies.py:29-30) after putting their own project root at the front of sys.path.  `install()` registers a
package called `Louvre_Evacuation` in `sys.modules` whose three hot-path modules are this repo's drop-in classes; every
other submodule (`utils.visualization`, `utils.reward_visualizer`, `envs.map`, ...) still resolves to the reference tree
given by `reference_root`, because the shim packages keep the reference directories on their `__path__`.  Modules already
in `sys.modules` win over sys.path, so the runner's own `sys.path.insert(0, project_root)` does not undo it.

    python -m dqn_marl_b200.compat /path/to/DQN-MARL/Louvre_Evacuation/runners/train_double_dqn.py [runner args]

Nothing here computes anything: the classes fail loudly without a CUDA device, as everywhere else in this package.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types
from typing import Optional

PACKAGE = "Louvre_Evacuation"
# reference module -> module of this package that replaces it
HOT_PATH_MODULES = {
    "envs.evacuation_env": "dqn_marl_b200.envs.evacuation_env",            # EvacuationEnv
    "envs.evacuation_env_multi": "dqn_marl_b200.envs.evacuation_env",      # EvacuationEnvMulti
    "agents.dqn_agent": "dqn_marl_b200.agents.dqn_agent",                  # DQNAgent
}


def _package(name: str, paths) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [p for p in paths if p and os.path.isdir(p)]
    m.__package__ = name
    sys.modules[name] = m
    return m


def install(reference_root: Optional[str] = None, package: str = PACKAGE) -> types.ModuleType:
    """Register `package` (default Louvre_Evacuation) with the B200 env / agent modules.  reference_root = directory that
    CONTAINS the reference's `Louvre_Evacuation/` (its project root), or None when only the hot-path modules are needed."""
    ref_pkg = os.path.join(reference_root, package) if reference_root else None
    top = _package(package, [ref_pkg])
    for sub in ("envs", "agents", "utils", "runners"):
        m = _package(f"{package}.{sub}", [os.path.join(ref_pkg, sub) if ref_pkg else None])
        setattr(top, sub, m)
    for ref_name, ours in HOT_PATH_MODULES.items():
        mod = importlib.import_module(ours)
        full = f"{package}.{ref_name}"
        sys.modules[full] = mod
        parent, leaf = full.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, mod)
    return top


def run_runner(path: str, argv=None):
    """Execute an unmodified reference runner file as __main__ with the shim installed."""
    path = os.path.abspath(path)
    project_root = os.path.dirname(os.path.dirname(os.path.dirname(path)))      # .../<root>/Louvre_Evacuation/runners/x.py
    install(project_root if os.path.isdir(os.path.join(project_root, PACKAGE)) else None)
    old = sys.argv
    sys.argv = [path] + list(argv or [])
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run_runner(sys.argv[1], sys.argv[2:])
