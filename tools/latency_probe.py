import time, numpy as np, torch
from cumind_b200 import Agent, Config
cfg = Config(action_space_size=2, observation_shape=(4,), hidden_dim=64, num_simulations=25)
agent = Agent(cfg)
obs = np.ones(4, np.float32)
for _ in range(20): agent.select_action(obs)
torch.cuda.synchronize()
t0=time.perf_counter()
N=500
for _ in range(N): a,p = agent.select_action(obs)
t1=time.perf_counter()
print("configs[0] quick start: select_action (1 tree, 25 simulations) %.1f us per call = %.0f simulations/s; plan %s" % ((t1-t0)/N*1e6, 25*N/(t1-t0), agent.mcts.plan(1)))
for _ in range(20): agent.select_action(obs, training=True)
t0=time.perf_counter()
for _ in range(N): a,p = agent.select_action(obs, training=True)
t1=time.perf_counter()
print("training=True %.1f us per call" % ((t1-t0)/N*1e6))
