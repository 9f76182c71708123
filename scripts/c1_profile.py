"""C1 (configs/dqn.yaml: one env, 150 people, B = 32) through the drop-in facades: wall clock per call, and where the host time
goes (cProfile).  `python scripts/c1_profile.py [iters] [--cprofile]`; under `ncu --metrics gpu__time_duration.sum` the same
script yields the launch list of one C1 iteration."""
import cProfile, io, os, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200.envs.evacuation_env import EvacuationEnv
from dqn_marl_b200.agents.dqn_agent import DQNAgent

iters = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 400
prof = "--cprofile" in sys.argv
env = EvacuationEnv(width=36, height=30, num_people=150, seed=1)
agent = DQNAgent(env.state_size, env.action_size, torch.device("cuda:0"),
                 dict(gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995, learning_rate=1e-4, batch_size=32,
                      target_update_freq=200, warmup_steps=0, memory_size=50000, seed=2))
state = env.reset()
T = dict(act=0.0, step=0.0, remember=0.0, learn=0.0)
n = 0
warm = min(100, iters // 4)
pr = cProfile.Profile()
for it in range(iters):
    if prof and it == warm:
        pr.enable()
    t0 = time.perf_counter(); a = agent.act(state, training=True)
    t1 = time.perf_counter(); nstate, r, done, info = env.step(a)
    t2 = time.perf_counter(); agent.remember(state, a, r, nstate, done)
    t3 = time.perf_counter(); loss = agent.learn() if len(agent.memory) > agent.batch_size else None
    t4 = time.perf_counter()
    state = env.reset() if done else nstate
    if it >= warm and loss is not None:
        T["act"] += t1 - t0; T["step"] += t2 - t1; T["remember"] += t3 - t2; T["learn"] += t4 - t3; n += 1
if prof:
    pr.disable()
print("C1 facade, one env, ms per call: " + ", ".join(f"{k} {v / max(n, 1) * 1e3:.3f}" for k, v in T.items())
      + f"; iteration {sum(T.values()) / max(n, 1) * 1e3:.3f} ms over {n} iterations")
if prof:
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())
