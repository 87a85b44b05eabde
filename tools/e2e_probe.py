"""Where does the host-buffer path (Agent.select_actions) spend its time?  Wall-clock split of one call."""
import ctypes as C
import time

import numpy as np
import torch

from cumind_b200 import Config, _native
from cumind_b200.agent import Agent
from cumind_b200.network import _stream_ptr

cfg = Config(hidden_dim=64, num_blocks=2, num_simulations=50, action_space_size=2, observation_shape=(4,), c_puct=1.25)
agent = Agent(cfg)
B = 4096
rng = np.random.default_rng(0)
obs = rng.standard_normal((B, 4)).astype(np.float32)
noise = rng.dirichlet([0.3] * 2, size=B)
for _ in range(20):
    agent.select_actions(obs, training=False, noise=noise)
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    agent.select_actions(obs, training=False, noise=noise)
t1 = time.perf_counter()
print("select_actions total      %.1f us" % ((t1 - t0) / N * 1e6))
net, lib = agent.network, agent.mcts._lib
tree = agent.mcts._tree(B, 50, 2)
scfg = agent.mcts._cfg(1, 0, 0)
need = int(lib.cumind_select_actions_work_bytes(C.byref(net._desc), B, 2))
work = torch.empty(need + 256, dtype=torch.uint8, device=net.device)
wbase = work.data_ptr() + ((-work.data_ptr()) % 256)
pin = next(iter(agent._host_plans.values()))["pin"]
args = (tree.handle, net._handle, C.c_void_p(pin["obs"].data_ptr()), C.byref(scfg), C.c_void_p(pin["noise"].data_ptr()),
        C.c_void_p(pin["actions"].data_ptr()), C.c_void_p(pin["probs"].data_ptr()), C.c_void_p(wbase), need, _stream_ptr())
for _ in range(20):
    lib.cumind_select_actions_host(*args)
t0 = time.perf_counter()
for _ in range(N):
    lib.cumind_select_actions_host(*args)
t1 = time.perf_counter()
print("C call only (incl. sync)  %.1f us" % ((t1 - t0) / N * 1e6))
t0 = time.perf_counter()
for _ in range(N):
    pin["obs"].numpy()[...] = obs
    pin["noise"].numpy()[...] = noise
    p = pin["probs"].numpy().copy()
    a = pin["actions"].numpy().copy()
t1 = time.perf_counter()
print("numpy staging copies      %.1f us" % ((t1 - t0) / N * 1e6))
t0 = time.perf_counter()
for _ in range(N):
    agent.mcts._tree(B, 50, 2)
    agent.mcts._cfg(1, 0, 0)
    lib.cumind_select_actions_work_bytes(C.byref(net._desc), B, 2)
t1 = time.perf_counter()
print("tree/cfg/work_bytes       %.1f us" % ((t1 - t0) / N * 1e6))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(N):
    lib.cumind_select_actions_host(*args)
e.record()
torch.cuda.synchronize()
print("C call, device events     %.1f us" % (s.elapsed_time(e) / N * 1e3))
