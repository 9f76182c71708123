"""ORACLE / TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's replay
(Louvre_Evacuation/agents/dqn_agent.py:88-89 deque(maxlen), :97-99 remember, :132-140 sample + stack).
The reference samples with the global MT19937 ``random.sample``; parity tests either inject the picked
indices or use the keyed permutation of oracle/keyed_draws.py (both must give rows of the deque)."""
from collections import deque

import numpy as np

from keyed_draws import sample_indices


class ReplayOracle:
    def __init__(self, capacity, seed=0):
        self.memory = deque(maxlen=capacity)     # dqn_agent.py:88-89
        self.seed = seed

    def remember(self, state, action, reward, next_state, done):     # dqn_agent.py:97-99
        self.memory.append((state, action, reward, next_state, done))

    def __len__(self):
        return len(self.memory)

    def sample(self, B, indices=None, draw_id=0):
        if indices is None:
            indices = sample_indices(self.seed, draw_id, len(self.memory), B)
        batch = [self.memory[int(i)] for i in indices]
        states, actions, rewards, next_states, dones = zip(*batch)     # dqn_agent.py:133-140
        return dict(states=np.array(states).astype(np.float32), actions=np.array(actions, dtype=np.int64),
                    rewards=np.array(rewards).astype(np.float32), next_states=np.array(next_states).astype(np.float32),
                    dones=np.array(dones, dtype=np.uint8), idx=np.asarray(indices, dtype=np.int64))
