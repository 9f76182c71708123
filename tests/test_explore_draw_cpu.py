"""mq_qnet_explore_draw (include/marl_b200.h) is the HOST evaluation of the keyed exploration draw of mq_qnet_act (the
reference's `np.random.random() <= self.epsilon` -> `random.randrange(5)`, dqn_agent.py:103-104): it must agree with the oracle's
keyed draws word for word.  Host-only: no GPU needed."""
import ctypes

import numpy as np


def test_host_exploration_draw_matches_the_oracle_draws():
    from dqn_marl_b200 import _lib
    from keyed_draws import Draws
    lib = _lib.load()
    out = ctypes.pointer(ctypes.c_int32(-1))
    n_explore = 0
    for seed in (0, 99, (1 << 63) - 5):
        for env in (0, 7, 4095):
            d = Draws(seed, env)
            for tick in (0, 1, 17, 123456, 0xFFFFFFFF):
                d.tick = tick
                for robot in (0, 1):
                    u, ra = d.agent_u_action(robot)
                    for eps in (0.0, 0.02, 0.3, 1.0):
                        out.contents.value = -1
                        rc = lib.mq_qnet_explore_draw(eps, seed, env, tick, robot, out)
                        want = eps > 0 and u <= np.float32(eps)
                        assert rc == int(want), (seed, env, tick, robot, eps)
                        if want:
                            assert out.contents.value == ra
                            n_explore += 1
                        else:
                            assert out.contents.value == -1          # untouched when the agent exploits
    assert n_explore > 90
