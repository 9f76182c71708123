"""The Louvre_Evacuation shim (dqn_marl_b200/compat.py): the reference runners' absolute imports resolve to the B200
classes, everything else still comes from the reference tree, and the runner's own sys.path manipulation does not undo it.
CPU only: with the REAL reference tree (when present) the unmodified runners get as far as constructing this package's
env, which refuses a box without CUDA — the GPU half is tests/test_runners_gpu.py."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_reference(tmp_path):
    pkg = tmp_path / "proj" / "Louvre_Evacuation"
    for sub in ("envs", "agents", "utils", "runners"):
        (pkg / sub).mkdir(parents=True)
    (pkg / "__init__.py").write_text("")
    (pkg / "utils" / "__init__.py").write_text("")
    (pkg / "agents" / "__init__.py").write_text("")
    (pkg / "utils" / "reward_visualizer.py").write_text("class RewardTracker:\n    origin = 'reference'\n")
    (pkg / "envs" / "evacuation_env.py").write_text("class EvacuationEnv:\n    origin = 'reference'\n")
    (pkg / "envs" / "evacuation_env_multi.py").write_text("class EvacuationEnvMulti:\n    origin = 'reference'\n")
    (pkg / "agents" / "dqn_agent.py").write_text("class DQNAgent:\n    origin = 'reference'\n")
    runner = pkg / "runners" / "train_x.py"
    runner.write_text(textwrap.dedent("""
        import os, sys
        project_root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..'))
        if project_root not in sys.path:
            sys.path.insert(0, project_root)                     # what the reference runners do (train_dqn.py:11-14)
        from Louvre_Evacuation.agents.dqn_agent import DQNAgent
        from Louvre_Evacuation.envs.evacuation_env import EvacuationEnv
        from Louvre_Evacuation.envs.evacuation_env_multi import EvacuationEnvMulti
        from Louvre_Evacuation.utils.reward_visualizer import RewardTracker
        print(EvacuationEnv.__module__, EvacuationEnvMulti.__module__, DQNAgent.__module__, RewardTracker.origin, sys.argv[1:])
    """))
    return runner


def test_runner_imports_resolve_to_b200_classes(tmp_path):
    runner = _fake_reference(tmp_path)
    out = subprocess.run([sys.executable, "-m", "dqn_marl_b200.compat", str(runner), "--episodes", "3"], cwd=ROOT,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = out.stdout.strip().splitlines()[-1]
    assert line == ("dqn_marl_b200.envs.evacuation_env dqn_marl_b200.envs.evacuation_env dqn_marl_b200.agents.dqn_agent "
                    "reference ['--episodes', '3']"), line


def test_install_without_reference_tree():
    code = ("import sys; sys.path.insert(0, %r); from dqn_marl_b200 import compat; compat.install();"
            "from Louvre_Evacuation.envs.evacuation_env import EvacuationEnv;"
            "from Louvre_Evacuation.agents.dqn_agent import DQNAgent;"
            "import dqn_marl_b200.envs.evacuation_env as e; assert EvacuationEnv is e.EvacuationEnv; print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr


def _staged_reference(tmp_path):
    import shutil
    import pytest
    for cand in (os.environ.get("MARL_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "Louvre_Evacuation", "runners")):
            root = tmp_path / "proj"           # private writable copy: the runners create <root>/dqn_results
            shutil.copytree(os.path.join(cand, "Louvre_Evacuation"), root / "Louvre_Evacuation", ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            shutil.copytree(os.path.join(cand, "configs"), root / "configs")
            return root
    pytest.skip("no reference tree available")


def _run_compat(*args):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, "-m", "dqn_marl_b200.compat", "--headless", *map(str, args)], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_real_runners_reach_the_b200_classes(tmp_path):
    """Unmodified train_double_dqn.py / train_qmix.py / train_dqn.py of the reference: imports resolve (matplotlib stubbed
    when absent), configs/dqn.yaml is read, and the first object they build is OUR env, which fails loudly without CUDA."""
    root = _staged_reference(tmp_path)
    for runner in ("train_double_dqn.py", "train_qmix.py"):
        out = _run_compat(root / "Louvre_Evacuation" / "runners" / runner)
        assert out.returncode != 0 and "dqn_marl_b200 runs on CUDA devices only" in out.stderr, (runner, out.stderr[-800:])
        assert "dqn_marl_b200/envs/evacuation_env.py" in out.stderr
    out = _run_compat(root / "Louvre_Evacuation" / "runners" / "train_dqn.py")       # catches and prints (train_dqn.py:222-227)
    assert "dqn_marl_b200 runs on CUDA devices only" in out.stdout + out.stderr
    assert (root / "dqn_results" / "reward_logs").is_dir()                              # the reference's RewardTracker ran


def test_main_entry_point_symbol_is_repaired(tmp_path):
    """Louvre_Evacuation/main.py:14,17 imports `main` from runners.train_dqn, which defines only train_dqn(): under the shim
    `python -m Louvre_Evacuation.main --train_dqn` reaches the training function instead of dying with ImportError."""
    root = _staged_reference(tmp_path)
    out = _run_compat("--root", root, "-m", "Louvre_Evacuation.main", "--train_dqn")
    assert "cannot import name" not in out.stderr, out.stderr[-800:]
    assert "dqn_marl_b200 runs on CUDA devices only" in out.stdout + out.stderr, out.stdout[-500:] + out.stderr[-800:]
