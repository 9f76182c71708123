"""dqn_marl_b200 — B200-native hot path of LX-530/DQN-MARL (Louvre_Evacuation).

Python/PyTorch host code over hand-written sm_100a CUDA behind a C-ABI (include/marl_b200.h).
"""
__version__ = "0.1.0"
