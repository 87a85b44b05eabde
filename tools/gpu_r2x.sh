#!/bin/bash
PYTHONPATH=. timeout 300 python tools/bench_selfplay.py 2>&1 | tail -1
PYTHONPATH=. timeout 300 python - <<'PY' 2>&1 | head -45
import cProfile, pstats, torch, numpy as np
from cumind_b200 import Agent, Config
from cumind_b200.envs import CartPoleVec
from cumind_b200.memory import MemoryBuffer
from cumind_b200.self_play import SelfPlay
cfg = Config(hidden_dim=64, num_blocks=2, num_simulations=50, action_space_size=2, observation_shape=(4,), c_puct=1.25)
agent = Agent(cfg)
sp = SelfPlay(agent, MemoryBuffer(100000))
sp.run_batch(CartPoleVec(4096, seed=1), 5, training=True, philox_seed=1)
pr = cProfile.Profile(); pr.enable()
sp.run_batch(CartPoleVec(4096, seed=2), 60, training=True, philox_seed=7)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
PY
