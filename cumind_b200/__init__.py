"""cumind_b200 — B200-native implementation of CuMind's MuZero search hot path.

Drop-in for the reference's `Agent / Config / MCTS / Node / CuMindNetwork` surface
(carletonai/CuMind v0.1.7: src/cumind/agent/agent.py, config.py, core/mcts.py, core/network.py);
the arithmetic runs in hand-written sm_100a CUDA behind the C ABI of include/cumind_b200.h.
"""

from .config import Config
from .prng import Rngs, key

__version__ = "0.1.0"
__all__ = ["Agent", "Trainer", "Config", "MCTS", "Node", "CuMindNetwork", "Rngs", "key", "__version__"]


def __getattr__(name):  # torch / the CUDA library are only imported when the compute classes are used
    if name in ("Agent",):
        from .agent import Agent
        return Agent
    if name in ("MCTS", "Node"):
        from . import mcts
        return getattr(mcts, name)
    if name == "Trainer":
        from .trainer import Trainer
        return Trainer
    if name == "CuMindNetwork":
        from .network import CuMindNetwork
        return CuMindNetwork
    raise AttributeError(name)
