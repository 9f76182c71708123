"""The torch fp32 restatement of the reference's Q-network / learn() lives with the other oracles: oracle/qnet_oracle.py."""
from qnet_oracle import build_nets, forward, learn_step  # noqa: F401
