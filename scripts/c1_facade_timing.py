"""BASELINE.json configs[0] (configs/dqn.yaml: one 36x30 env, 150 people, B = 32, replay 50 000) through the drop-in facades:
wall-clock per env.step / agent.act / agent.learn, i.e. what an unmodified reference runner sees (every call synchronises)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqn_marl_b200.envs.evacuation_env import EvacuationEnv
from dqn_marl_b200.agents.dqn_agent import DQNAgent

env = EvacuationEnv(width=36, height=30, num_people=150, seed=1)
agent = DQNAgent(env.state_size, env.action_size, torch.device("cuda:0"),
                 dict(gamma=0.99, epsilon=1.0, epsilon_min=0.02, epsilon_decay=0.9995, learning_rate=1e-4, batch_size=32,
                      target_update_freq=200, warmup_steps=0, memory_size=50000, seed=2))
state = env.reset()
t_step = t_act = t_learn = 0.0
n_step = n_learn = 0
for it in range(400):
    t0 = time.perf_counter(); a = agent.act(state, training=True); t1 = time.perf_counter()
    nstate, r, done, info = env.step(a); t2 = time.perf_counter()
    agent.remember(state, a, r, nstate, done)
    t3 = time.perf_counter()
    loss = agent.learn() if len(agent.memory) > agent.batch_size else None
    t4 = time.perf_counter()
    state = env.reset() if done else nstate
    if it >= 100:
        t_act += t1 - t0; t_step += t2 - t1; n_step += 1
        if loss is not None:
            t_learn += t4 - t3; n_learn += 1
print(f"facade, one env: env.step {t_step / n_step * 1e3:.3f} ms, agent.act {t_act / n_step * 1e3:.3f} ms, agent.learn(B=32) {t_learn / max(n_learn, 1) * 1e3:.3f} ms"
      f"  (reference on one CPU core, BASELINE.md: 7.7 / 0.95 / 55 ms)")
