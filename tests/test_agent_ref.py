"""CPU: the torch fp32 restatement (tests/torch_ref.py) reproduces what the UNMODIFIED reference DQNAgent
produced (tests/golden/agent_ref.npz, made by oracle/make_golden_agent.py) — pins the floating-point oracle."""
import numpy as np
import torch

import torch_ref
from util import load_golden


def _batch(g, idx):
    f = lambda k, dt: torch.tensor(g[k][idx], dtype=dt)
    return (f("states", torch.float32), f("actions", torch.int64), f("rewards", torch.float32), f("next_states", torch.float32),
            torch.tensor(g["dones"][idx].astype(bool)))


def test_torch_ref_matches_reference_golden():
    g = load_golden("agent_ref.npz")
    m = g["meta"]
    q, t = torch_ref.build_nets(m["seed"], m["target_perturb_seed"])
    B = m["cfg"]["batch_size"]
    x = torch.tensor(g["states"][:B], dtype=torch.float32)
    with torch.no_grad():
        # bit-equal on the machine that made the golden; 1e-6 slack for another CPU's oneDNN/MKL code path
        np.testing.assert_allclose(torch_ref.forward(q, x).numpy(), g["q_online"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(torch_ref.forward(t, x).numpy(), g["q_target"], rtol=1e-6, atol=1e-7)
    opt = torch.optim.Adam(q.parameters(), lr=m["cfg"]["learning_rate"])
    for step in range(2):
        loss, _ = torch_ref.learn_step(q, t, opt, _batch(g, g["idx"][step]))
        assert abs(loss - m["A"][f"loss{step}"]) <= 1e-6 * abs(loss)
        for name, cs in m["A"][f"params{step}"].items():
            v = q.state_dict()[name].reshape(-1)
            np.testing.assert_allclose(v[torch.tensor(cs["idx"])].numpy(), np.array(cs["sample"], dtype=np.float32), rtol=1e-5, atol=1e-7,
                                       err_msg=name)
    # scenario B: injected dropout masks
    q, t = torch_ref.build_nets(m["seed"], m["target_perturb_seed"])
    opt = torch.optim.Adam(q.parameters(), lr=m["cfg"]["learning_rate"])
    loss, _ = torch_ref.learn_step(q, t, opt, _batch(g, g["idx"][0]), drop_online=torch.tensor(g["mask_online"]),
                                   drop_target=torch.tensor(g["mask_target"]))
    assert abs(loss - m["B"]["loss0"]) <= 1e-6 * abs(loss)


def test_param_layout_roundtrip():
    from dqn_marl_b200.agents import qnet_params as qp
    torch.manual_seed(0)
    net = qp.TorchDQN()
    flat = torch.zeros(qp.TOTAL)
    qp.pack(net.state_dict(), flat)
    back = qp.unpack(flat)
    assert list(back) == qp.NAMES == list(net.state_dict())
    for k, v in net.state_dict().items():
        assert torch.equal(back[k], v), k
    assert qp.TOTAL == 8157093
    # conv layout: [(kh*3+kw)*Cin + c][Cout]; fc1 layout: [n][p*128 + c]
    w = net.state_dict()["conv2.weight"]
    o = qp.OFFSETS[2]
    assert flat[o + ((1 * 3 + 2) * 32 + 7) * 64 + 11] == w[11, 7, 1, 2]
    w = net.state_dict()["fc1.weight"]
    o = qp.OFFSETS[6]
    assert flat[o + 5 * 15488 + 37 * 128 + 9] == w[5, 9 * 121 + 37]
