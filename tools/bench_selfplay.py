"""Batched self-play throughput: B CartPole environments in lockstep, one fused search per step.
PYTHONPATH=. python tools/bench_selfplay.py [--envs 4096] [--steps 60]"""
import argparse
import time

import numpy as np
import torch

from cumind_b200 import Agent, Config
from cumind_b200.envs import CartPoleVec
from cumind_b200.memory import MemoryBuffer
from cumind_b200.self_play import SelfPlay

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--sims", type=int, default=50)
args = ap.parse_args()
cfg = Config(hidden_dim=64, num_blocks=2, num_simulations=args.sims, action_space_size=2, observation_shape=(4,), c_puct=1.25)
agent = Agent(cfg)
sp = SelfPlay(agent, MemoryBuffer(100000))
sp.run_batch(CartPoleVec(args.envs, seed=1), 5, training=True, philox_seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
eps = sp.run_batch(CartPoleVec(args.envs, seed=2), args.steps, training=True, philox_seed=7)
dt = time.perf_counter() - t0
n_ep = sum(len(e) for e in eps)
print(f"{args.envs} environments x {args.steps} steps: {1e3 * dt / args.steps:.3f} ms per lock-step = "
      f"{args.envs * args.steps / dt:.3e} environment steps/s = {args.envs * args.steps * args.sims / dt:.3e} simulations/s; "
      f"{n_ep} episodes, {len(sp.memory)} in the replay buffer")
