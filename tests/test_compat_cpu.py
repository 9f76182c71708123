"""The Louvre_Evacuation shim (dqn_marl_b200/compat.py): the reference runners' absolute imports resolve to the B200
classes, everything else still comes from the reference tree, and the runner's own sys.path manipulation does not undo it.
CPU only: nothing is instantiated (the classes need a CUDA device)."""
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
