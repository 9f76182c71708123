"""Run the reference's runners UNMODIFIED on the B200 classes.

The reference's runners import their environment and agent by absolute module path
(`from Louvre_Evacuation.envs.evacuation_env import EvacuationEnv`, `...envs.evacuation_env_multi import
EvacuationEnvMulti`, `...agents.dqn_agent import DQNAgent`; runners/train_dqn.py:19-22, train_double_dqn.py:12-13,
evaluate_strategies.py:29-30) after putting their own project root at the front of sys.path.  `install()` registers a
package called `Louvre_Evacuation` in `sys.modules` whose three hot-path modules are this repo's drop-in classes; every
other submodule (`utils.visualization`, `utils.reward_visualizer`, `envs.map`, ...) still resolves to the reference tree
given by `reference_root`, because the shim packages keep the reference directories on their `__path__`.  Modules already
in `sys.modules` win over sys.path, so the runner's own `sys.path.insert(0, project_root)` does not undo it.

    python -m dqn_marl_b200.compat [--headless] /path/to/DQN-MARL/Louvre_Evacuation/runners/train_double_dqn.py [runner args]
    python -m dqn_marl_b200.compat [--headless] --root /path/to/DQN-MARL -m Louvre_Evacuation.main --train_dqn

`--headless` registers a do-nothing `matplotlib` when the real one is not installed (the reference's `utils/visualization.py:8`
and `utils/reward_visualizer.py:8` import it at module scope; its plots are reports, not part of the hot path), so that
`runners/train_dqn.py` runs on a box without plotting libraries.  The shim also repairs the reference's entry point:
`Louvre_Evacuation/main.py:14,17` imports `main` from `runners.train_dqn`, which only defines `train_dqn()`
(`train_dqn.py:28`); after that module has executed, `main` is aliased to `train_dqn`.

Nothing here computes anything: the classes fail loudly without a CUDA device, as everywhere else in this package.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
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


class _AliasAfterExec(importlib.abc.Loader):
    """Wraps a module's real loader; once the module body has run, missing names are aliased (ALIASES)."""

    def __init__(self, loader, aliases):
        self._loader, self._aliases = loader, aliases

    def create_module(self, spec):
        return self._loader.create_module(spec)

    def exec_module(self, module):
        self._loader.exec_module(module)
        for new, old in self._aliases.items():
            if not hasattr(module, new) and hasattr(module, old):
                setattr(module, new, getattr(module, old))


class _AliasFinder(importlib.abc.MetaPathFinder):
    """main.py:14,17 does `from Louvre_Evacuation.runners.train_dqn import main` (falling back to `runners.train_dqn`);
    train_dqn.py defines `train_dqn`, not `main`.  The fix lives here, not in the reference: the module is loaded by the
    normal machinery and gets `main = train_dqn` afterwards."""
    ALIASES = {"runners.train_dqn": {"main": "train_dqn"}}

    def __init__(self, package):
        self._names = {f"{package}.{k}": v for k, v in self.ALIASES.items()}
        self._names.update(self.ALIASES)                      # the `from runners.train_dqn import ...` fallback form
        self._busy = False

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in self._names or self._busy:
            return None
        self._busy = True
        try:
            spec = importlib.util.find_spec(fullname)
        finally:
            self._busy = False
        if spec is None or spec.loader is None:
            return None
        spec.loader = _AliasAfterExec(spec.loader, self._names[fullname])
        return spec


def install_headless_matplotlib() -> bool:
    """Register a do-nothing `matplotlib` (pyplot, patches, ...) when the real package is missing.  Every attribute is a
    callable / indexable / iterable dummy, so module-scope statements such as `plt.rcParams['font.sans-serif'] = [...]`
    (visualization.py:14-15) and the plotting calls inside the runners' try blocks succeed without drawing anything.
    Returns True if the stub was installed."""
    try:
        import matplotlib  # noqa: F401
        return False
    except ImportError:
        pass

    class _Dummy(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__") and name.endswith("__"):
                raise AttributeError(name)
            d = _Dummy(f"{self.__name__}.{name}")
            object.__setattr__(self, name, d)
            return d

        def __call__(self, *a, **k):
            return _Dummy(self.__name__ + "()")

        def __getitem__(self, key):
            return _Dummy(f"{self.__name__}[{key!r}]")

        def __setitem__(self, key, value):
            pass

        def __iter__(self):                  # `fig, ((ax1, ax2), (ax3, ax4)) = plt.subplots(2, 2)` (visualization.py:215)
            return iter((_Dummy("a"), _Dummy("b")))

        def __len__(self):
            return 2

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    root = _Dummy("matplotlib")
    root.__path__ = []                       # a package: `import matplotlib.pyplot` consults sys.modules first
    sys.modules["matplotlib"] = root
    for sub in ("pyplot", "patches", "cm", "colors", "animation", "gridspec", "font_manager"):
        m = _Dummy(f"matplotlib.{sub}")
        sys.modules[f"matplotlib.{sub}"] = m
        object.__setattr__(root, sub, m)
    return True


def _prepare(project_root: Optional[str], headless: bool):
    if headless:
        install_headless_matplotlib()
    install(project_root if project_root and os.path.isdir(os.path.join(project_root, PACKAGE)) else None)
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder(PACKAGE))


def run_runner(path: str, argv=None, headless: bool = False):
    """Execute an unmodified reference runner file as __main__ with the shim installed."""
    path = os.path.abspath(path)
    project_root = os.path.dirname(os.path.dirname(os.path.dirname(path)))      # .../<root>/Louvre_Evacuation/runners/x.py
    _prepare(project_root, headless)
    old = sys.argv
    sys.argv = [path] + list(argv or [])
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old


def run_module(module: str, project_root: str, argv=None, headless: bool = False):
    """`python -m <module>` of the reference tree (e.g. Louvre_Evacuation.main --train_dqn, main.py:1) under the shim."""
    project_root = os.path.abspath(project_root)
    _prepare(project_root, headless)
    if project_root not in sys.path:
        sys.path.insert(0, project_root)
    old = sys.argv
    sys.argv = [module] + list(argv or [])
    try:
        return runpy.run_module(module, run_name="__main__", alter_sys=True)
    finally:
        sys.argv = old


def main(argv=None):
    args = list(sys.argv[1:] if argv is None else argv)
    headless, root = False, None
    while args and args[0] in ("--headless", "--root"):
        if args[0] == "--headless":
            headless = True
            args = args[1:]
        else:
            root, args = args[1], args[2:]
    if not args:
        raise SystemExit(__doc__)
    if args[0] == "-m":
        if root is None or len(args) < 2:
            raise SystemExit("usage: python -m dqn_marl_b200.compat [--headless] --root <dir containing Louvre_Evacuation/> -m <module> [args]")
        return run_module(args[1], root, args[2:], headless)
    return run_runner(args[0], args[1:], headless)


if __name__ == "__main__":
    main()
