"""GPU parity of the Q-network / learner kernels (csrc/qnet.cu) through the C-ABI:
Q-values and losses within 1e-5 relative of the reference's PyTorch fp32 agent (north_star), gradients and
post-Adam parameters against the pinned torch restatement (tests/torch_ref.py)."""
import numpy as np
import pytest
import torch

import torch_ref
from util import load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-5          # north_star: "within 1e-5 relative in fp32"


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _qnet(q, t, max_batch=64):
    from dqn_marl_b200.agents.qnet import QNet
    net = QNet("cuda:0", max_batch=max_batch)
    net.load_state_dict(q.state_dict(), "online")
    net.load_state_dict(t.state_dict(), "target")
    return net


def _hp(step, lr=1e-4, gamma=0.99, clip=1.0, huber=0):
    from dqn_marl_b200 import _lib
    hp = _lib.MqHparams()
    hp.gamma, hp.lr, hp.beta1, hp.beta2, hp.adam_eps, hp.clip_norm, hp.huber, hp.adam_step = gamma, lr, 0.9, 0.999, 1e-8, clip, huber, step
    return hp


def _dev_batch(g, idx):
    d = "cuda:0"
    return dict(states=torch.tensor(g["states"][idx], dtype=torch.float32, device=d),
                actions=torch.tensor(g["actions"][idx], dtype=torch.int64, device=d),
                rewards=torch.tensor(g["rewards"][idx], dtype=torch.float32, device=d),
                next_states=torch.tensor(g["next_states"][idx], dtype=torch.float32, device=d),
                dones=torch.tensor(g["dones"][idx], dtype=torch.uint8, device=d))


def _cpu_batch(g, idx):
    f = lambda k, dt: torch.tensor(g[k][idx], dtype=dt)
    return (f("states", torch.float32), f("actions", torch.int64), f("rewards", torch.float32), f("next_states", torch.float32),
            torch.tensor(g["dones"][idx].astype(bool)))


def test_q_values_match_reference_golden():
    g = load_golden("agent_ref.npz")
    m = g["meta"]
    q, t = torch_ref.build_nets(m["seed"], m["target_perturb_seed"])
    net = _qnet(q, t)
    B = m["cfg"]["batch_size"]
    x = torch.tensor(g["states"][:B], dtype=torch.float32, device="cuda:0")
    qo = net.forward(x, "online").cpu().numpy()
    qt = net.forward(x, "target").cpu().numpy()
    assert _rel(qo, g["q_online"]) <= RTOL and _rel(qt, g["q_target"]) <= RTOL
    np.testing.assert_allclose(qo, g["q_online"], rtol=RTOL, atol=2e-7)
    # greedy act == reference's act(training=False) choices
    a = net.act(x.reshape(B, 726), 0.0, 1, 0, 0).cpu().numpy()
    assert a.tolist() == m["greedy"]
    # single-sample batches (the reference's act() path, B = 1)
    for k in range(3):
        q1 = net.forward(x[k:k + 1], "online").cpu().numpy()
        np.testing.assert_allclose(q1[0], g["q_online"][k], rtol=RTOL, atol=2e-7)


@pytest.mark.parametrize("scenario", ["A", "B"])
def test_learn_matches_reference_golden(scenario):
    from dqn_marl_b200.agents import qnet_params as qp
    g = load_golden("agent_ref.npz")
    m = g["meta"]
    q, t = torch_ref.build_nets(m["seed"], m["target_perturb_seed"])
    net = _qnet(q, t)
    opt = torch.optim.Adam(q.parameters(), lr=m["cfg"]["learning_rate"])
    steps = 2 if scenario == "A" else 1
    for step in range(steps):
        idx = g["idx"][step]
        masks = (None, None)
        if scenario == "B":
            masks = (torch.tensor(g["mask_online"], device="cuda:0"), torch.tensor(g["mask_target"], device="cuda:0"))
        loss = net.td_backward(_dev_batch(g, idx), _hp(step + 1), masks[0], masks[1]).item()
        ref_loss = m[scenario][f"loss{step}"]
        assert abs(loss - ref_loss) <= RTOL * abs(ref_loss), (loss, ref_loss)
        # torch restatement of the same step (pinned to the golden on CPU): gradients and parameters
        cm = (None, None) if scenario == "A" else (torch.tensor(g["mask_online"]), torch.tensor(g["mask_target"]))
        p_before = {k: v.clone() for k, v in q.state_dict().items()}
        q.zero_grad()
        # unclipped gradients of the torch restatement
        batch = _cpu_batch(g, idx)
        cur = torch_ref.forward(q, batch[0], cm[0]).gather(1, batch[1].unsqueeze(1))
        with torch.no_grad():
            target = batch[2] + (0.99 * torch_ref.forward(t, batch[3], cm[1]).max(1)[0] * ~batch[4])
        torch.nn.functional.mse_loss(cur.squeeze(), target).backward()
        grads = qp.unpack(net.flat_g)
        for name, p in q.named_parameters():
            gr = grads[name].cpu().numpy()
            assert _rel(gr, p.grad.numpy()) <= 2e-4, (name, _rel(gr, p.grad.numpy()))
        gnorm = net.clip_adam(_hp(step + 1)).item()
        ref_norm = float(torch.nn.utils.clip_grad_norm_(q.parameters(), 1.0))
        assert abs(gnorm - ref_norm) <= 1e-4 * ref_norm
        opt.step()
        new = net.state_dict("online")
        for name, p in q.state_dict().items():
            upd_ref = (p - p_before[name]).numpy()
            upd = (new[name].cpu() - p_before[name]).numpy()
            # Adam's first steps move every weight by ~lr * g/(|g| + eps): elements whose gradient is of the order of
            # eps = 1e-8 amplify fp32 rounding noise of g, so allow a handful of them up to 0.2 * lr
            err = np.abs(new[name].cpu().numpy() - p.numpy())
            assert err.max() <= 2e-5 and (err > 2e-6).mean() <= 1e-3, (name, err.max(), (err > 2e-6).mean())
            assert np.abs(upd - upd_ref).max() <= 0.2 * max(np.abs(upd_ref).max(), 1e-12), name
        # golden parameter samples of the real reference
        for name, cs in m[scenario][f"params{step}"].items():
            v = new[name].reshape(-1).cpu().numpy()[cs["idx"]]
            np.testing.assert_allclose(v, np.array(cs["sample"], dtype=np.float32), rtol=0, atol=2e-5, err_msg=name)


def test_large_batch_vs_torch_fp32():
    """B = 640 (not a multiple of the 128-row tile) on random data: forward, loss, gradients vs torch fp32 (CPU)."""
    from dqn_marl_b200.agents import qnet_params as qp
    q, t = torch_ref.build_nets(7, 8)
    net = _qnet(q, t, max_batch=640)
    B = 640
    gen = torch.Generator().manual_seed(3)
    states = (torch.rand((B, 11, 11, 6), generator=gen) < 0.3).float() * torch.rand((B, 11, 11, 6), generator=gen)
    nstates = (torch.rand((B, 11, 11, 6), generator=gen) < 0.3).float()
    actions = torch.randint(0, 5, (B,), generator=gen)
    rewards = torch.randn(B, generator=gen) * 10
    dones = torch.rand(B, generator=gen) < 0.1
    d = "cuda:0"
    batch = dict(states=states.to(d), actions=actions.to(d), rewards=rewards.to(d), next_states=nstates.to(d), dones=dones.to(torch.uint8).to(d))
    with torch.no_grad():
        ref_q = torch_ref.forward(q, states).numpy()
    np.testing.assert_allclose(net.forward(batch["states"]).cpu().numpy(), ref_q, rtol=RTOL, atol=5e-7)
    for huber in (0, 1):
        loss = net.td_backward(batch, _hp(1, huber=huber)).item()
        q.zero_grad()
        cur = torch_ref.forward(q, states).gather(1, actions.unsqueeze(1)).squeeze()
        with torch.no_grad():
            target = rewards + (0.99 * torch_ref.forward(t, nstates).max(1)[0] * ~dones)
        ref = torch.nn.functional.smooth_l1_loss(cur, target) if huber else torch.nn.functional.mse_loss(cur, target)
        ref.backward()
        assert abs(loss - float(ref)) <= RTOL * abs(float(ref))
        grads = qp.unpack(net.flat_g)
        for name, p in q.named_parameters():
            assert _rel(grads[name].cpu().numpy(), p.grad.numpy()) <= 3e-4, name


def test_epsilon_greedy_keyed_draws_and_target_sync():
    from keyed_draws import Draws
    q, t = torch_ref.build_nets(11, 12)
    net = _qnet(q, t, max_batch=256)
    E, R, seed, base, tick = 100, 2, 99, 40, 17
    x = torch.rand((E * R, 726), device="cuda:0")
    qv = net.forward(x.reshape(-1, 11, 11, 6)).cpu().numpy()
    a = net.act(x, 0.3, seed, base, tick, R).cpu().numpy()
    n_rand = 0
    for b in range(E * R):
        d = Draws(seed, base + b // R); d.tick = tick
        u, ra = d.agent_u_action(b % R)
        exp = ra if u <= np.float32(0.3) else int(qv[b].argmax())
        n_rand += u <= np.float32(0.3)
        assert a[b] == exp, b
    assert 30 < n_rand < 95
    # hard copy and Polyak update (dqn_agent.py:170-172; tau < 1 is the north_star option)
    before_t = net.flat_t.clone()
    net.sync_target(0.25)
    assert torch.allclose(net.flat_t, 0.25 * net.flat_p + 0.75 * before_t, rtol=0, atol=1e-7)
    net.sync_target(1.0)
    assert torch.equal(net.flat_t, net.flat_p)
    m = net.dropout_mask(4096, 5, 1)
    assert abs(m.float().mean().item() - 0.8) < 0.01 and not torch.equal(m, net.dropout_mask(4096, 5, 2))


def test_dqn_agent_facade_reference_surface(tmp_path):
    """Constructor/config defaults, remember/learn gating (warmup quirk), checkpoint dict keys and torch-layout
    state_dict shapes (loadable by the reference), act() return type."""
    from dqn_marl_b200.agents.dqn_agent import DQNAgent
    torch.manual_seed(5)
    cfg = dict(gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995, learning_rate=1e-4, batch_size=32,
               target_update_freq=200, warmup_steps=0, memory_size=500, seed=3, dropout="eval")
    agent = DQNAgent((11, 11, 6), 5, torch.device("cuda:0"), cfg)
    torch.manual_seed(5)
    ref_q, _ = torch_ref.build_nets(5)
    for k, v in agent.q_network.state_dict().items():
        assert torch.equal(v.cpu(), ref_q.state_dict()[k]), k          # same default init as the reference under the same seed
    assert agent.learn() is None                                        # len(memory) < batch_size
    rng = np.random.default_rng(0)
    for i in range(40):
        s = rng.random((11, 11, 6)); ns = rng.random((11, 11, 6))
        a = agent.act(s, training=True)
        assert isinstance(a, int) and 0 <= a < 5
        agent.remember(s, a, float(rng.normal()), ns, bool(i % 7 == 0))
    assert len(agent.memory) == 40
    eps0 = agent.epsilon
    loss = agent.learn()
    assert isinstance(loss, float) and agent.steps == 1 and agent.epsilon == eps0 * 0.9995
    with torch.no_grad():
        x = torch.tensor(rng.random((4, 11, 11, 6)), dtype=torch.float32, device="cuda:0")
        assert agent.q_network(x).shape == (4, 5)
    path = tmp_path / "ck.pth"
    agent.save(str(path))
    ck = torch.load(str(path), map_location="cpu", weights_only=True)
    assert set(ck) == {"q_network", "target_network", "optimizer", "epsilon", "steps"}
    assert ck["q_network"]["fc1.weight"].shape == (512, 15488) and ck["q_network"]["conv1.weight"].shape == (32, 6, 3, 3)
    ref_q.load_state_dict(ck["q_network"])                              # a reference-side module accepts it
    torch.optim.Adam(ref_q.parameters(), lr=1e-4).load_state_dict(ck["optimizer"])
    agent2 = DQNAgent((11, 11, 6), 5, torch.device("cuda:0"), dict(cfg, seed=4))
    agent2.load(str(path))
    assert torch.equal(agent2.net.flat_p, agent.net.flat_p) and torch.equal(agent2.net.flat_m, agent.net.flat_m)
    assert agent2.epsilon == agent.epsilon and agent2.steps == 1
    # default warmup_steps = 1000 when the key is absent: learn() never runs (reference quirk, dqn_agent.py:80,128)
    agent3 = DQNAgent((11, 11, 6), 5, torch.device("cuda:0"), {"memory_size": 100, "seed": 1})
    for i in range(40):
        agent3.remember(rng.random((11, 11, 6)), 1, 0.0, rng.random((11, 11, 6)), False)
    assert agent3.learn() is None


def test_single_state_fast_paths_equal_the_device_paths():
    """The numpy-state paths of the drop-in agent (pinned staging for act / remember, exploration decided on the host by
    mq_qnet_explore_draw so that an exploring act() launches nothing — dqn_agent.py:103-104 returns before the forward too) make
    the same decisions and store the same transitions as the tensor paths that run everything on the device."""
    from dqn_marl_b200.agents.dqn_agent import DQNAgent
    cfg = dict(epsilon=0.5, batch_size=32, warmup_steps=0, memory_size=300, seed=77)
    torch.manual_seed(9); fast = DQNAgent((11, 11, 6), 5, torch.device("cuda:0"), cfg)
    torch.manual_seed(9); slow = DQNAgent((11, 11, 6), 5, torch.device("cuda:0"), cfg)
    rng = np.random.default_rng(3)
    n_explore, launches0 = 0, fast.net.launch_count
    for i in range(120):
        s, ns = rng.random((11, 11, 6)), rng.random((11, 11, 6))
        before = fast.net.launch_count
        a = fast.act(s, training=True)
        n_explore += fast.net.launch_count == before
        b = slow.act(torch.from_numpy(s.astype(np.float32)), training=True)          # tensor input: forward + in-kernel draw
        assert isinstance(a, int) and a == b, i
        r, d = float(rng.normal()), bool(i % 11 == 0)
        fast.remember(s, a, r, ns, d)
        slow.remember(torch.from_numpy(s.astype(np.float32)), a, r, torch.from_numpy(ns.astype(np.float32)), d)
    assert 35 < n_explore < 85 and fast._act_calls == slow._act_calls == 120 and fast._mask_calls == slow._mask_calls
    for name in ("state", "next_state", "action", "reward", "done"):
        assert torch.equal(getattr(fast.memory, name)[:120], getattr(slow.memory, name)[:120]), name
    s = rng.random((11, 11, 6))                                                       # greedy call: always the forward
    assert fast.act(s, training=False) == slow.act(torch.from_numpy(s.astype(np.float32)), training=False)
    la, lb = fast.learn(), slow.learn()
    assert la == lb and torch.equal(fast.net.flat_p, slow.net.flat_p)
    la, lb = fast.learn(), slow.learn()                                              # second call reuses the resident batch tensors
    assert la == lb and torch.equal(fast.net.flat_p, slow.net.flat_p)


@pytest.mark.parametrize("B", [256, 200])
def test_bf16_tensor_core_path_tracks_fp32(B):
    """The tcgen05 bf16 path (every layer but the 5-output head on tensor cores, fp32 accumulate / master weights) against the fp32
    parity path on the same weights and batch: Q within 2e-2 of the Q scale, loss within 1e-2, gradients aligned.  B = 200 ends
    in a partial 128-row tile (the im2col rows the conv1 weight gradient reads are stored by clipped bulk tensor stores)."""
    from dqn_marl_b200.agents import qnet_params as qp
    q, t = torch_ref.build_nets(21, 22)
    gen = torch.Generator().manual_seed(4)
    states = (torch.rand((B, 11, 11, 6), generator=gen) < 0.3).float() * torch.rand((B, 11, 11, 6), generator=gen)
    nstates = (torch.rand((B, 11, 11, 6), generator=gen) < 0.3).float()
    d = "cuda:0"
    batch = dict(states=states.to(d), actions=torch.randint(0, 5, (B,), generator=gen).to(d), rewards=(torch.randn(B, generator=gen) * 0.1).to(d),
                 next_states=nstates.to(d), dones=(torch.rand(B, generator=gen) < 0.1).to(torch.uint8).to(d))
    mask = (torch.rand((B, 512), generator=gen) >= 0.2).to(torch.uint8).to(d)
    res = {}
    for prec in ("fp32", "bf16"):
        net = _qnet(q, t, max_batch=B)
        net.set_precision(prec)
        qv = net.forward(batch["states"]).cpu()
        loss = net.td_backward(batch, _hp(1), mask, mask).item()
        grads = {k: v.cpu() for k, v in qp.unpack(net.flat_g).items()}
        gnorm = net.clip_adam(_hp(1)).item()
        res[prec] = (qv, loss, grads, gnorm, net.flat_p.clone().cpu())
    q32, l32, g32, n32, p32 = res["fp32"]
    q16, l16, g16, n16, p16 = res["bf16"]
    assert (q32 - q16).abs().max() <= 2e-2 * q32.abs().max()
    assert abs(l32 - l16) <= 1e-2 * abs(l32)
    assert abs(n32 - n16) <= 3e-2 * n32
    for k in g32:
        a, b = g32[k].flatten().double(), g16[k].flatten().double()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.995, (k, cos)
        assert abs(float(b.norm() / a.norm()) - 1.0) < 5e-2, (k, float(b.norm() / a.norm()))
    assert (p32 - p16).abs().max() <= 2.1e-4          # one Adam step moves a weight by at most lr = 1e-4 in either path


def _bench_batch(B, gen, d="cuda:0"):
    """Observations shaped like the env's: channels 1, 3, 4, 5 in {0, 1}, channel 2 (danger) in [0, 1], channel 0 == 0."""
    def obs():
        x = (torch.rand((B, 11, 11, 6), generator=gen) < 0.25).float()
        x[..., 2] = torch.rand((B, 11, 11), generator=gen) * (torch.rand((B, 11, 11), generator=gen) < 0.4)
        x[..., 0] = 0
        return x
    return dict(states=obs().to(d), actions=torch.randint(0, 5, (B,), generator=gen).to(d), rewards=(torch.randn(B, generator=gen) * 0.5).to(d),
                next_states=obs().to(d), dones=(torch.rand(B, generator=gen) < 0.05).to(torch.uint8).to(d))


def test_bf16_path_tracks_fp32_at_the_bench_batch_sizes():
    """The sizes bench.py runs: learn B = 4096, act B = 16384 (C3: one observation per env).  Same bars as the small-batch test;
    greedy actions of the two paths agree except where the top-two Q gap is inside the bf16 error."""
    from dqn_marl_b200.agents import qnet_params as qp
    q, t = torch_ref.build_nets(31, 32)
    gen = torch.Generator().manual_seed(6)
    B, BA = 4096, 16384
    batch = _bench_batch(B, gen)
    act_obs = _bench_batch(BA, gen)["states"]
    mask = (torch.rand((B, 512), generator=gen) >= 0.2).to(torch.uint8).to("cuda:0")
    res = {}
    for prec in ("fp32", "bf16"):
        net = _qnet(q, t, max_batch=BA)
        net.set_precision(prec)
        qa = torch.empty((BA, 5), dtype=torch.float32, device="cuda:0")
        acts = net.act(act_obs, 0.0, 1, 0, 0, 1, None, q_out=qa).cpu()
        loss = net.td_backward(batch, _hp(1), mask, mask).item()
        grads = {k: v.cpu() for k, v in qp.unpack(net.flat_g).items()}
        gnorm = net.clip_adam(_hp(1)).item()
        res[prec] = (qa.cpu(), acts, loss, grads, gnorm)
        net.close()
    q32, a32, l32, g32, n32 = res["fp32"]
    q16, a16, l16, g16, n16 = res["bf16"]
    scale = q32.abs().max()
    assert (q32 - q16).abs().max() <= 2e-2 * scale
    top2 = q32.topk(2, dim=1).values
    gap = top2[:, 0] - top2[:, 1]
    differ = a32 != a16
    assert float(differ.float().mean()) < 0.05
    assert bool((gap[differ] <= 4e-2 * scale).all())          # only near-ties flip
    assert abs(l32 - l16) <= 1e-2 * abs(l32) and abs(n32 - n16) <= 3e-2 * n32
    for k in g32:
        a, b = g32[k].flatten().double(), g16[k].flatten().double()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.995, (k, cos)
        assert abs(float(b.norm() / a.norm()) - 1.0) < 5e-2, (k, float(b.norm() / a.norm()))


def test_bf16_path_does_not_drift_over_50_learn_steps():
    """50 consecutive learn steps (target sync every 10) on the same stream of batches and dropout masks, fp32 parity path vs bf16
    tensor-core path from the same initial weights.  Band: every loss within 3 % (+ 1e-4 absolute); the parameter displacements
    of equal length within 5 % and aligned (cosine > 0.8, measured 0.88: Adam divides by sqrt(v), so the many weights whose
    gradients are near zero move by ~lr per step in a direction set by rounding noise in EITHER path); Q-values of the two
    trained networks on a held-out batch within 15 % of the Q scale (measured 8 %: the freshly initialised network's Q-values
    are ~0.1 and 50 Adam steps of 3e-4 move every weight by up to 0.015, so that noise is visible at this scale; the loss
    band is the meaningful bar)."""
    q, t = torch_ref.build_nets(41, 42)
    B, STEPS = 512, 50
    curves, disp, held = {}, {}, {}
    for prec in ("fp32", "bf16"):
        gen = torch.Generator().manual_seed(9)
        net = _qnet(q, t, max_batch=B)
        net.set_precision(prec)
        p0 = net.flat_p.clone()
        losses = []
        for s in range(STEPS):
            batch = _bench_batch(B, gen)
            m1 = (torch.rand((B, 512), generator=gen) >= 0.2).to(torch.uint8).to("cuda:0")
            m2 = (torch.rand((B, 512), generator=gen) >= 0.2).to(torch.uint8).to("cuda:0")
            losses.append(net.td_backward(batch, _hp(s + 1, lr=3e-4), m1, m2).item())
            net.clip_adam(_hp(s + 1, lr=3e-4))
            if s % 10 == 9:
                net.sync_target(1.0)
        curves[prec] = np.array(losses)
        disp[prec] = (net.flat_p - p0).double().cpu()
        held[prec] = net.forward(_bench_batch(B, torch.Generator().manual_seed(77))["states"]).cpu()
        net.close()
    a, b = curves["fp32"], curves["bf16"]
    assert np.all(np.abs(a - b) <= 3e-2 * np.abs(a) + 1e-4), np.abs(a - b) / np.abs(a)
    assert (held["fp32"] - held["bf16"]).abs().max() <= 0.15 * held["fp32"].abs().max()      # measured 0.08: see the docstring
    da, db = disp["fp32"], disp["bf16"]
    cos = float((da @ db) / (da.norm() * db.norm()))
    assert cos > 0.8 and abs(float(db.norm() / da.norm()) - 1.0) < 5e-2, (cos, float(db.norm() / da.norm()))


def test_q_network_autograd_qmix_step_matches_torch():
    """train_qmix.py:88-113 on the drop-in agents: Q-values of two agents go through a monotonic mixing network, ONE
    loss.backward() fills both agents' gradient buffers through mq_qnet_backward, clip_grad_norm_ works on
    q_network.parameters(), optimizer.step() applies Adam.  Compared with the same step on torch modules (fp32)."""
    from dqn_marl_b200.agents import qnet_params as qp
    from dqn_marl_b200.agents.dqn_agent import DQNAgent
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False            # the torch side of this comparison must be real fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(gamma=0.99, learning_rate=1e-4, batch_size=16, warmup_steps=0, memory_size=100, dropout="eval")
    agents = [DQNAgent((11, 11, 6), 5, dev, dict(cfg, seed=11 + k)) for k in range(2)]
    refs = []
    for k, ag in enumerate(agents):
        torch.manual_seed(100 + k)
        q, t = torch_ref.build_nets(100 + k)
        ag.q_network.load_state_dict(q.state_dict()); ag.target_network.load_state_dict(t.state_dict())
        refs.append((q.to(dev).eval(), t.to(dev).eval()))

    class Mixer(torch.nn.Module):                      # MixingNetwork of train_qmix.py:39-54
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(9)
            self.w1 = torch.nn.Parameter(torch.randn(2, 32, generator=g)); self.b1 = torch.nn.Parameter(torch.zeros(32))
            self.w2 = torch.nn.Parameter(torch.randn(32, 1, generator=g)); self.b2 = torch.nn.Parameter(torch.zeros(1))

        def forward(self, q):
            return (torch.relu(q @ self.w1.abs() + self.b1) @ self.w2.abs() + self.b2).squeeze(-1)

    B = 16
    g = torch.Generator().manual_seed(4)
    s = [(torch.rand((B, 11, 11, 6), generator=g) < 0.3).float().to(dev) for _ in range(2)]
    ns = [(torch.rand((B, 11, 11, 6), generator=g) < 0.3).float().to(dev) for _ in range(2)]
    a = [torch.randint(0, 5, (B,), generator=g).to(dev) for _ in range(2)]
    r = torch.randn(B, generator=g).to(dev)
    d = (torch.rand(B, generator=g) < 0.2).to(dev)

    def qmix_loss(qnets, tnets, mixer, tmixer):
        q = torch.stack([qnets[k](s[k]).gather(1, a[k].unsqueeze(1)).squeeze(1) for k in range(2)], dim=1)
        with torch.no_grad():
            nq = torch.stack([tnets[k](ns[k]).max(1)[0] for k in range(2)], dim=1)
            y = r + 0.99 * tmixer(nq) * (~d)
        return torch.nn.functional.mse_loss(mixer(q), y)

    mix_a, mix_b = Mixer().to(dev), Mixer().to(dev)
    tmix = Mixer().to(dev)
    # ours
    for ag in agents:
        ag.optimizer.zero_grad()
    loss_a = qmix_loss([ag.q_network for ag in agents], [ag.target_network for ag in agents], mix_a, tmix)
    loss_a.backward()
    # torch reference
    opts = [torch.optim.Adam(q.parameters(), lr=1e-4) for q, _ in refs]
    loss_b = qmix_loss([lambda x, q=q: torch_ref.forward(q, x) for q, _ in refs], [lambda x, t=t: torch_ref.forward(t, x) for _, t in refs],
                       mix_b, tmix)
    loss_b.backward()
    assert abs(loss_a.item() - loss_b.item()) <= 1e-5 * abs(loss_b.item())
    for p_a, p_b in zip(mix_a.parameters(), mix_b.parameters()):
        assert torch.allclose(p_a.grad, p_b.grad, rtol=1e-4, atol=1e-6)
    for k, ag in enumerate(agents):
        ours = qp.unpack(ag.net.flat_g)
        gmax = max(p.grad.abs().max().item() for p in refs[k][0].parameters())
        for name, p in refs[k][0].named_parameters():
            assert (ours[name] - p.grad).abs().max().item() <= 3e-4 * gmax, (k, name)
        n_a = torch.nn.utils.clip_grad_norm_(ag.q_network.parameters(), 1.0)
        n_b = torch.nn.utils.clip_grad_norm_(refs[k][0].parameters(), 1.0)
        assert abs(n_a.item() - n_b.item()) <= 1e-4 * n_b.item()
        before = ag.net.flat_p.clone()
        ag.optimizer.step(); opts[k].step()
        assert not torch.equal(before, ag.net.flat_p)
        after = qp.unpack(ag.net.flat_p)
        for name, p in refs[k][0].named_parameters():
            assert (after[name] - p.data).abs().max().item() <= 2e-5, (k, name)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_td_backward_in_two_parts_equals_one_call(precision):
    """mq_qnet_td_backward_part: part 1 (forwards, loss, fc backward) + part 2 (conv backward) == mq_qnet_td_backward, and the
    fc gradients are already final after part 1 (what the overlapped all-reduce of VecDQNAgent relies on)."""
    from dqn_marl_b200.agents import qnet_params as qp
    q, t = torch_ref.build_nets(21, target_perturb_seed=5)
    net = _qnet(q, t, max_batch=64)
    net.set_precision(precision)
    g = torch.Generator().manual_seed(8)
    B = 64
    batch = dict(states=(torch.rand((B, 11, 11, 6), generator=g) < 0.3).float().cuda(), actions=torch.randint(0, 5, (B,), generator=g).cuda(),
                 rewards=torch.randn(B, generator=g).cuda(), next_states=(torch.rand((B, 11, 11, 6), generator=g) < 0.3).float().cuda(),
                 dones=(torch.rand(B, generator=g) < 0.2).to(torch.uint8).cuda())
    hp = _hp(1)
    mask = net.dropout_mask(B, 3, 1)
    loss0 = net.td_backward(batch, hp, mask, mask).clone()
    g0 = net.flat_g.clone()
    head = qp.OFFSETS[6]
    net.flat_g.fill_(float("nan"))
    loss1 = net.td_backward(batch, hp, mask, mask, part=1).clone()
    torch.cuda.synchronize()
    assert torch.equal(net.flat_g[head:], g0[head:]) and torch.isnan(net.flat_g[:head]).all()
    net.td_backward(batch, hp, mask, mask, part=2)
    torch.cuda.synchronize()
    assert torch.equal(net.flat_g, g0) and torch.equal(loss0, loss1)


@pytest.mark.gpu
def test_conv1_built_from_observation_equals_im2col_path():
    """bf16 path: forward() builds the conv1 operand tiles in shared memory straight from the observation, the online forward
    inside td_backward() goes through the im2col buffer (its weight gradient needs it).  Same bf16 operands, same MMAs: the TD
    loss recomputed from forward()'s Q-values must equal the loss td_backward() reports, up to the fp32 rounding of the mean.
    Ragged batch sizes exercise the partial last 128-row tile."""
    q, t = torch_ref.build_nets(31, 32)
    for B in (8, 136, 1000):
        gen = torch.Generator().manual_seed(B)
        d = "cuda:0"
        states = ((torch.rand((B, 11, 11, 6), generator=gen) < 0.3).float() * torch.rand((B, 11, 11, 6), generator=gen)).to(d)
        nstates = (torch.rand((B, 11, 11, 6), generator=gen) < 0.3).float().to(d)
        batch = dict(states=states, actions=torch.randint(0, 5, (B,), generator=gen).to(d), rewards=(torch.randn(B, generator=gen) * 0.1).to(d),
                     next_states=nstates, dones=(torch.rand(B, generator=gen) < 0.1).to(torch.uint8).to(d))
        mask = (torch.rand((B, 512), generator=gen) >= 0.2).to(torch.uint8).to(d)
        net = _qnet(q, t, max_batch=B)
        net.set_precision("bf16")
        hp = _hp(1)
        q_on = net.forward(states, "online", mask)
        q_tg = net.forward(nstates, "target", mask)
        y = batch["rewards"] + hp.gamma * q_tg.max(dim=1).values * (1.0 - batch["dones"].float())
        ref = ((q_on.gather(1, batch["actions"].view(-1, 1)).squeeze(1) - y) ** 2).mean().item()
        loss = net.td_backward(batch, hp, mask, mask).item()
        assert abs(loss - ref) <= 1e-5 * abs(ref) + 1e-9, (B, loss, ref)
