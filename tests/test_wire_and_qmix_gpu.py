"""Compact observation wire format (mq_env_set_obs_wire / mq_obs_wire_expand) against the dense f32 observations, and the
batched QMIX-lite loop (runners/train_qmix_vec.py, reference runners/train_qmix.py:62-118)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n_robots,N,E", [(1, 150, 50), (2, 150, 33), (1, 400, 6)])
def test_wire_observations_expand_to_the_dense_ones(n_robots, N, E):
    """Two env batches with the same seed step in lockstep, one through the dense host path, one through the wire path: the
    host observations must be bit-identical every step — including windows that hang over the grid edge and a robot parked far
    off the map (evaluate_strategies.py:83)."""
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import Layout
    lay = Layout.reference_room(n_robots=n_robots)
    a = VecEvacuationEnv(lay, E, N, device=DEV, seed=4, strict_reference=False, auto_reset=True)
    b = VecEvacuationEnv(lay, E, N, device=DEV, seed=4, strict_reference=False, auto_reset=True)
    a.reset(); b.reset()
    a.robots[0, 0, 0], a.robots[0, 0, 1] = 1000, 1000                 # env 0: robot 0 far off the map
    b.robots[0, 0, 0], b.robots[0, 0, 1] = 1000, 1000
    a.scalars[0, 8], a.scalars[0, 9] = 1000, 1000
    b.scalars[0, 8], b.scalars[0, 9] = 1000, 1000
    rng = np.random.default_rng(1)
    for t in range(40):
        act = torch.tensor(rng.integers(0, 5, size=(E, n_robots)).astype(np.int32)).pin_memory()
        a.step_async(act)
        b.step_async(act, wire=True)
        oa, ra, da = a.step_wait()
        ob, rb, db = b.step_wait()
        assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)), f"step {t}"
        assert torch.equal(ra.view(torch.int64), rb.view(torch.int64)) and torch.equal(da, db)
    assert b.h_wire.numel() * 4 == E * n_robots * 544
    # hybrid transfer: part of the envs dense, the rest in wire form, same host buffers; then back to the dense form
    for frac in (0.4, 0.02, 0.97, False):
        act = torch.tensor(rng.integers(0, 5, size=(E, n_robots)).astype(np.int32)).pin_memory()
        a.h_obs.fill_(-3.0); b.h_obs.fill_(-5.0)
        a.step_async(act); b.step_async(act, wire=frac)
        oa, ra, da = a.step_wait()
        ob, rb, db = b.step_wait()
        assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)), frac
        assert torch.equal(ra.view(torch.int64), rb.view(torch.int64)) and torch.equal(da, db)


def test_wire_from_reset_and_explicit_threads():
    from dqn_marl_b200 import _lib
    from dqn_marl_b200.envs import VecEvacuationEnv
    from dqn_marl_b200.layout import Layout
    lay = Layout.synthetic(96, 80, n_exits=2, seed=7)
    E = 9
    env = VecEvacuationEnv(lay, E, 300, device=DEV, seed=2, strict_reference=True, auto_reset=False)
    wire = torch.zeros((E, 1, _lib.MQ_OBS_WIRE_WORDS), dtype=torch.int32, device=DEV)
    _lib.check(env.lib.mq_env_set_obs_wire(env._h, _lib.ptr(wire)))
    dense = env.reset().clone()
    h = wire.cpu()
    out = torch.empty((E, 1, 11, 11, 6), dtype=torch.float32)
    for threads in (1, 3):
        out.fill_(-1)
        _lib.check(env.lib.mq_obs_wire_expand(_lib.ptr(h), E, _lib.ptr(out), threads))
        assert torch.equal(out.view(torch.int32), dense.cpu().view(torch.int32))
    _lib.check(env.lib.mq_env_set_obs_wire(env._h, None))


def test_batched_qmix_loop_trains_both_agents_through_the_mixer():
    from dqn_marl_b200.layout import Layout
    from dqn_marl_b200.runners.train_qmix_vec import VecQmixTrainer
    torch.manual_seed(0)
    tr = VecQmixTrainer(Layout.reference_room(n_robots=2), 24, 60, DEV, dict(batch_size=48, learning_rate=1e-4, epsilon=0.5, dropout="eval"),
                        seed=3, replay_capacity=4096, target_sync_prob=0.5)
    p0 = [a.net.flat_p.clone() for a in tr.agents]
    m0 = [p.detach().clone() for p in tr.mixing.parameters()]
    losses = []
    for t in range(12):
        l = tr.step()
        if l is not None:
            losses.append(float(l))
    assert len(losses) == 11 and all(np.isfinite(losses))                # len(memory) >= batch_size from the second step on
    assert tr.learn_steps == 11 and 0 < tr.target_syncs <= 11
    for a, p in zip(tr.agents, p0):
        assert not torch.equal(a.net.flat_p, p)                           # one loss.backward() reached BOTH agents' kernels
        assert a._adam_t == 11 and a.epsilon == 0.5                       # the reference's QMIX loop never decays epsilon
    assert any(not torch.equal(p.detach(), q) for p, q in zip(tr.mixing.parameters(), m0))
    # the two rings are one joint replay: same rows, same order
    x, y = tr.agents[0].memory.sample(32, want_idx=True), tr.agents[1].memory.sample(32, want_idx=True)
    assert torch.equal(x["idx"], y["idx"]) and torch.equal(x["rewards"], y["rewards"]) and torch.equal(x["dones"], y["dones"])
    # one learn step on a frozen joint batch lowers that batch's loss (lr 1e-4 on the nets, 1e-3 on the mixer)
    a1, a2 = tr.agents
    with torch.no_grad():
        def joint_loss():
            q1 = a1.q_network(x["states"]).gather(1, x["actions"].unsqueeze(1)).squeeze(1)
            q2 = a2.q_network(y["states"]).gather(1, y["actions"].unsqueeze(1)).squeeze(1)
            nq = torch.stack([a1.target_network(x["next_states"]).max(1)[0], a2.target_network(y["next_states"]).max(1)[0]], dim=1)
            tgt = x["rewards"] + 0.99 * tr.target_mixing(nq) * (x["dones"] == 0)
            return float(torch.nn.functional.mse_loss(tr.mixing(torch.stack([q1, q2], dim=1)), tgt))
        assert np.isfinite(joint_loss())
