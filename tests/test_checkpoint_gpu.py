"""SURVEY.md §8 f2 (second half) and f4 (log emitters): snapshot / restore of the env batch, the replay ring and the whole
batched training loop — a resumed run continues BIT-IDENTICALLY — and the RewardTracker / PerformanceRecorder file formats
written by the batched runner (reference utils/reward_visualizer.py:97-122, utils/visualization.py:24-39)."""
import csv
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _room():
    from dqn_marl_b200.layout import Layout
    return Layout.reference_room()


def test_env_batch_snapshot_resumes_bit_identically():
    from dqn_marl_b200.envs import VecEvacuationEnv
    E, N = 40, 150
    env = VecEvacuationEnv(_room(), E, N, device=DEV, seed=11, env_id_base=7, strict_reference=False, auto_reset=True)
    env.reset()
    g = torch.Generator(device=DEV); g.manual_seed(3)
    acts = torch.randint(0, 5, (30, E, 1), generator=g, device=DEV, dtype=torch.int32)
    for t in range(12):
        env.step(acts[t])
    sd = env.state_dict()

    def run(e):
        out = []
        for t in range(12, 30):
            o, r, d = e.step(acts[t])
            out.append((o.clone(), r.clone(), d.clone()))
        return out, {k: getattr(e, k).clone() for k in e._STATE_TENSORS}

    a_out, a_state = run(env)
    env2 = VecEvacuationEnv(_room(), E, N, device=DEV, seed=11, env_id_base=7, strict_reference=False, auto_reset=True)
    env2.load_state_dict(sd)
    b_out, b_state = run(env2)
    for (o1, r1, d1), (o2, r2, d2) in zip(a_out, b_out):
        assert torch.equal(o1, o2) and torch.equal(r1.view(torch.int64), r2.view(torch.int64)) and torch.equal(d1, d2)
    for k in a_state:
        a, b = a_state[k], b_state[k]
        assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a, b.view(torch.int64) if b.dtype == torch.float64 else b), k
    with pytest.raises(ValueError):
        VecEvacuationEnv(_room(), E, N, device=DEV, seed=12, env_id_base=7).load_state_dict(sd)      # other keyed draws


def test_replay_ring_snapshot_wraps_and_resamples():
    from dqn_marl_b200.replay import ReplayRing
    cap, n = 1000, 384
    ring = ReplayRing(cap, device=DEV, seed=5)
    g = torch.Generator(device=DEV); g.manual_seed(1)
    for k in range(4):                                             # 1536 pushes into 1000 slots: wrapped
        s = torch.rand((n, 726), generator=g, device=DEV); ns = torch.rand((n, 726), generator=g, device=DEV)
        a = torch.randint(0, 5, (n,), generator=g, device=DEV, dtype=torch.int32)
        r = torch.rand((n,), generator=g, device=DEV, dtype=torch.float64); d = (torch.rand((n,), generator=g, device=DEV) < 0.1).to(torch.uint8)
        ring.push(s, a, r, ns, d)
    ring.sample(64)
    sd = ring.state_dict()
    assert sd["size"] == cap and sd["cursor"] == (4 * n) % cap and sd["draws"] == 1
    ring2 = ReplayRing(cap, device=DEV, seed=99)
    ring2.load_state_dict(sd)
    assert len(ring2) == len(ring) and ring2.cursor == ring.cursor
    for _ in range(3):
        x, y = ring.sample(128, want_idx=True), ring2.sample(128, want_idx=True)
        for k in x:
            assert torch.equal(x[k], y[k]), k
    s = torch.rand((n, 726), generator=g, device=DEV)
    z = torch.zeros(n, device=DEV)
    for rg in (ring, ring2):
        rg.push(s, z.to(torch.int32), z.to(torch.float64), s, z.to(torch.uint8))
    assert torch.equal(ring.state, ring2.state) and ring.cursor == ring2.cursor
    with pytest.raises(ValueError):
        ReplayRing(cap + 1, device=DEV).load_state_dict(sd)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_loop_checkpoint_resumes_bit_identically(tmp_path, precision):
    from dqn_marl_b200.runners.train_dqn_vec import VecTrainer
    E, N, B = 64, 40, 64
    cfg = dict(batch_size=B, learning_rate=1e-4, gamma=0.99, epsilon=0.5, epsilon_min=0.02, epsilon_decay=0.99, dropout="train",
               precision=precision)

    def make():
        torch.manual_seed(0)
        return VecTrainer(_room(), E, N, DEV, cfg, env_id_base=0, seed=21, replay_capacity=4096, target_sync_every=3)

    a = make()
    for _ in range(8):
        a.step()
    path = str(tmp_path / "loop.pt")
    a.save(path)

    def cont(tr):
        losses = []
        for _ in range(7):
            l = tr.step()
            losses.append(l.clone())
        torch.cuda.synchronize()
        return torch.cat(losses), tr.agent.net.flat_p.clone(), tr.agent.net.flat_t.clone(), tr.obs[tr.cur].clone(), tr.agent.epsilon

    la, pa, ta, oa, ea = cont(a)
    torch.manual_seed(12345)                                        # a fresh process would have other initial weights
    b = VecTrainer(_room(), E, N, DEV, cfg, env_id_base=0, seed=21, replay_capacity=4096, target_sync_every=3)
    b.load(path)
    lb, pb, tb, ob, eb = cont(b)
    assert torch.equal(la.view(torch.int32), lb.view(torch.int32)), (la, lb)
    assert torch.equal(pa.view(torch.int32), pb.view(torch.int32)) and torch.equal(ta.view(torch.int32), tb.view(torch.int32))
    assert torch.equal(oa, ob) and ea == eb
    assert a.stats.learn_steps == b.stats.learn_steps and a.agent.steps == b.agent.steps


def test_batched_runner_writes_reward_tracker_and_performance_recorder_files(tmp_path):
    from dqn_marl_b200.runners.train_dqn_vec import VecTrainer
    E, N = 32, 6                                                    # six people: episodes end by evacuation in ~50-80 steps
    torch.manual_seed(0)
    tr = VecTrainer(_room(), E, N, DEV, dict(batch_size=32, epsilon=1.0, dropout="eval"), seed=4, replay_capacity=8192,
                    strict_reference=False, log_dir=str(tmp_path))
    tr.recorder.flush_every = 16
    returns = torch.zeros(E, dtype=torch.float64, device=DEV)
    expect = []
    for t in range(150):
        tr.step(learn=(t % 10 == 0))
        returns += tr.reward
        done = tr.done.bool()
        expect += [(t, int(e), float(returns[e])) for e in torch.nonzero(done).flatten().tolist()]
        returns[done] = 0
    tr.recorder.save_data()
    assert len(expect) >= E                                          # every env finished at least one episode
    with open(tmp_path / "reward_logs" / "reward_data.json", encoding="utf-8") as f:
        data = json.load(f)
    # RewardTracker.save_data keys (reward_visualizer.py:99-106) and get_statistics keys (:65-76)
    assert set(data) == {"episode_rewards", "episode_steps", "episode_evacuation_rates", "episode_death_rates", "step_rewards", "statistics"}
    assert set(data["statistics"]) == {"total_episodes", "total_steps", "avg_reward", "max_reward", "min_reward", "std_reward",
                                       "recent_avg_reward", "avg_evacuation_rate", "avg_death_rate", "avg_steps_per_episode"}
    assert len(data["episode_rewards"]) == len(expect) and len(data["step_rewards"]) == 150
    assert np.array_equal(np.array(data["episode_rewards"]), np.array([r for _, _, r in expect]))     # same (step, env) order, same sums
    assert all(abs(a + b - 1.0) < 1e-12 for a, b in zip(data["episode_evacuation_rates"], data["episode_death_rates"]))
    with open(tmp_path / "reward_logs" / "episode_data.csv", encoding="utf-8") as f:
        rows = list(csv.DictReader(f))
    assert list(rows[0]) == ["episode", "reward", "steps", "evacuation_rate", "death_rate"] and len(rows) == len(expect)
    with open(tmp_path / "training_performance.csv", encoding="utf-8") as f:
        rows = list(csv.DictReader(f))
    assert list(rows[0]) == ["episode", "total_reward", "evacuated", "dead", "remaining", "evacuation_rate", "death_rate", "avg_health",
                             "min_health", "total_steps"]                                      # visualization.py:28-39
    assert all(int(r["evacuated"]) + int(r["dead"]) == N and int(r["remaining"]) == 0 for r in rows)
    assert all(0.0 <= float(r["min_health"]) <= 100.0 for r in rows)
    # the terminal observation, not the next episode's first one, is what was stored with done = 1 (train_dqn.py:104-107)
    assert int(tr.agent.memory.done[:len(tr.agent.memory)].sum().item()) == len(expect)
