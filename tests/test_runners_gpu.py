"""The reference's own runner files, UNMODIFIED, executed end to end on the B200 classes (SURVEY.md §8 f1 and the C1 row).

`python -m dqn_marl_b200.compat --headless <root>/Louvre_Evacuation/runners/<runner>.py` runs the reference's file with its
absolute imports (`Louvre_Evacuation.envs.evacuation_env`, `...agents.dqn_agent`) resolved to this package's classes; the
runner reads `<root>/configs/dqn.yaml` and writes `<root>/dqn_results/` exactly as in the reference
(runners/train_dqn.py:37-46).  The tree comes from baseline/_ref (scripts/stage_reference.py; git-ignored, travels to the
GPU box), $MARL_REFERENCE_ROOT or /root/reference; each test works on a private writable copy whose configs/dqn.yaml is the
reference's file with only the episode count / head count reduced, so the test ends in seconds.  Skipped when no
reference tree is available."""
import csv
import json
import os
import shutil
import subprocess
import sys

import pytest
import yaml

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.environ.get("MARL_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "Louvre_Evacuation", "runners")) and os.path.isfile(os.path.join(cand, "configs", "dqn.yaml")):
            return cand
    return None


@pytest.fixture()
def project(tmp_path):
    ref = _reference_root()
    if ref is None:
        pytest.skip("no reference tree (run scripts/stage_reference.py where /root/reference exists)")
    root = tmp_path / "proj"
    shutil.copytree(os.path.join(ref, "Louvre_Evacuation"), root / "Louvre_Evacuation",
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(ref, "configs", "dqn.yaml"), encoding="utf-8") as f:
        cfg = yaml.safe_load(f)
    (root / "configs").mkdir()
    return root, cfg


def _write_cfg(root, cfg):
    with open(root / "configs" / "dqn.yaml", "w", encoding="utf-8") as f:
        yaml.safe_dump(cfg, f)


def _run(root, *args, timeout=900):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "dqn_marl_b200.compat", "--headless", *map(str, args)], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=timeout)
    return out


REF_KEYS = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
            "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"]


def _check_checkpoint(path):
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=True)
    assert set(ck) == {"q_network", "target_network", "optimizer", "epsilon", "steps"}          # dqn_agent.py:176-182
    assert list(ck["q_network"].keys()) == REF_KEYS
    assert tuple(ck["q_network"]["fc1.weight"].shape) == (512, 15488) and tuple(ck["q_network"]["conv2.weight"].shape) == (64, 32, 3, 3)
    return ck


def test_train_dqn_c1_config_then_evaluate_and_main(project):
    """C1 = configs/dqn.yaml verbatim except `episodes` (100 -> 2): runners/train_dqn.py:93-125 loop, RewardTracker and
    PerformanceRecorder of the reference's utils fed by the facades, checkpoints in the reference's format; then
    evaluate_strategies.py loads that checkpoint (no-robot policy parks the robot at [1000, 1000], :83); then the package
    entry point `python -m Louvre_Evacuation.main --train_dqn` (main.py:14,17 symbol repaired by the shim)."""
    root, cfg = project
    cfg["episodes"] = 2
    _write_cfg(root, cfg)
    out = _run(root, root / "Louvre_Evacuation" / "runners" / "train_dqn.py")
    assert out.returncode == 0, out.stderr[-3000:]
    assert "训练成功完成" in out.stdout and "❌" not in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    res = root / "dqn_results"
    ck = _check_checkpoint(res / "dqn_model.pth")
    _check_checkpoint(res / "best_model.pth")
    # learn() ran every step once len(memory) > 32 (warmup_steps: 0).  Episodes are short: the fire never resets between episodes
    # (quirk Q6), the reference's own second episode ends after ~45 steps (tests/golden/traj_room_single.npz).
    with open(res / "reward_logs" / "reward_data.json", encoding="utf-8") as f:
        n_steps = sum(json.load(f)["episode_steps"])
    assert ck["steps"] == n_steps - 32 and ck["epsilon"] < 1.0
    with open(res / "reward_logs" / "reward_data.json", encoding="utf-8") as f:
        data = json.load(f)                                       # reward_visualizer.py:97-122
    assert len(data["episode_rewards"]) == 2 and len(data["step_rewards"]) == sum(data["episode_steps"])
    assert all(0.0 <= v <= 1.0 for v in data["episode_evacuation_rates"] + data["episode_death_rates"])
    with open(res / "training_performance.csv", encoding="utf-8") as f:
        rows = list(csv.DictReader(f))                            # visualization.py:24-39
    assert len(rows) == 2 and {"evacuated", "dead", "remaining", "avg_health", "total_steps"} <= set(rows[0])
    assert int(rows[0]["evacuated"]) + int(rows[0]["dead"]) + int(rows[0]["remaining"]) == cfg["env"]["num_people"]

    out = _run(root, root / "Louvre_Evacuation" / "runners" / "evaluate_strategies.py", "--episodes", "1")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "无机器人" in out.stdout and "DQN" in out.stdout, out.stdout[-2000:]

    shutil.rmtree(res)
    out = _run(root, "--root", root, "-m", "Louvre_Evacuation.main", "--train_dqn")
    assert out.returncode == 0, out.stderr[-3000:]
    assert (res / "dqn_model.pth").exists(), out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("runner,prefix", [("train_double_dqn.py", "double_dqn"), ("train_qmix.py", "qmix")])
def test_two_robot_runners(project, runner, prefix):
    """train_double_dqn.py:30-86 (two independent agents, shared reward) and train_qmix.py:62-118 (mixing network over both
    agents' Q-values, loss.backward() through `agent.q_network(...)`, clip_grad_norm_ on `q_network.parameters()`,
    `agent.optimizer.step()`).  Both hard-code 200 episodes; four people per episode keep them short."""
    root, cfg = project
    cfg["env"]["num_people"] = 4
    _write_cfg(root, cfg)
    out = _run(root, root / "Louvre_Evacuation" / "runners" / runner, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    res = root / "dqn_results"
    for k in (1, 2):
        ck = _check_checkpoint(res / f"{prefix}_agent{k}.pth")
        assert ck["steps"] > 0 or prefix == "qmix"               # the QMIX runner steps the optimizers itself (train_qmix.py:113)
    with open(res / f"{prefix}_training_log.csv", encoding="utf-8") as f:
        rows = list(csv.DictReader(f))
    assert len(rows) == 200
    assert all(abs(float(r["evac_rate"]) + float(r["death_rate"]) - 1.0) < 1e-9 or float(r["evac_rate"]) + float(r["death_rate"]) < 1.0 for r in rows)


def test_evaluate_all_strategies_runner(project):
    """runners/evaluate_all_strategies.py (five policies: no robot = robots parked at [1000, 1000] through the writable
    `env.map.robot_position(s)`, static robot, single DQN, double DQN, QMIX), unmodified: it loads five reference-format
    checkpoints — written here by this package's own `DQNAgent.save()` — into `DQNAgent(state_size, action_size, device, {})`
    (empty config: the defaults of dqn_agent.py:73-80) and plays whole episodes on `EvacuationEnv` / `EvacuationEnvMulti`."""
    import torch
    from dqn_marl_b200.agents.dqn_agent import DQNAgent
    root, cfg = project
    cfg["env"]["num_people"] = 6
    _write_cfg(root, cfg)
    res = root / "dqn_results"
    res.mkdir()
    for k, name in enumerate(("best_model.pth", "double_dqn_agent1.pth", "double_dqn_agent2.pth", "qmix_agent1.pth", "qmix_agent2.pth")):
        torch.manual_seed(100 + k)
        agent = DQNAgent((11, 11, 6), 5, torch.device("cuda:0"), {"memory_size": 64, "seed": k})
        agent.save(str(res / name))
        del agent
    out = _run(root, root / "Louvre_Evacuation" / "runners" / "evaluate_all_strategies.py", "--episodes", "1")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "五策略比较" in out.stdout, out.stdout[-2000:]
    rows = [ln.split("\t") for ln in out.stdout.splitlines() if ln.count("\t") == 3 and ln.rstrip().endswith(tuple("0123456789"))]
    assert [r[0] for r in rows] == ["无机器人", "静态机器人", "单机器人DQN", "双机器人DoubleDQN", "双机器人QMIX"], out.stdout[-2000:]
    for r in rows:
        assert 0.0 <= float(r[2].rstrip("%")) <= 100.0 and 0.0 < float(r[3]) <= 600.0          # death rate, simulated seconds


def test_overnight_experiments_runner(project):
    """runners/overnight_experiments.py, unmodified, on a 2 x 2 grid of reward coefficients with one episode each: the runner
    mutates the CLASS attributes `EvacuationEnv.DEATH_PENALTY / ALIVE_BONUS` between runs (:69-70; the facade forwards them to
    the kernel's reward coefficients), builds `EvacuationEnv(**cfg['env'])` and samples with `agent.act(state, training=True)`."""
    root, cfg = project
    cfg["env"]["num_people"] = 6
    _write_cfg(root, cfg)
    out = _run(root, root / "Louvre_Evacuation" / "runners" / "overnight_experiments.py", "--episodes", "1", "--death_min", "100",
               "--death_max", "150", "--death_step", "50", "--alive_min", "0.2", "--alive_max", "0.3", "--alive_step", "0.1")
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    res = root / "dqn_results"
    with open(res / "experiment_summary.csv", encoding="utf-8") as f:
        rows = list(csv.DictReader(f))
    assert len(rows) == 4 and {(int(float(r["death_penalty"])), round(float(r["alive_bonus"]), 1)) for r in rows} == {(100, 0.2), (100, 0.3), (150, 0.2), (150, 0.3)}
    assert all(0.0 <= float(r["death_rate"]) <= 1.0 and 0.0 <= float(r["evac_rate"]) <= 1.0 for r in rows)
    with open(res / "best_reward_cfg.json", encoding="utf-8") as f:
        best = json.load(f)
    assert set(best) == {"death_penalty", "alive_bonus"}
