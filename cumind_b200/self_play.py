"""Self-play on top of the batched search.

`SelfPlay.run_episode(env)` keeps the reference's single-environment loop and step record
(src/cumind/data/self_play.py:27-63: one `select_action(training=True)` per step, dicts with
observation / action / reward / policy / done).  `SelfPlay.run_batch(vec_env, num_steps)` is the
B200-shaped version: B environments advance in lockstep, one `Agent.select_actions` (one fused search
over B trees) per step, so the host loop no longer starves the GPU.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any, Dict, List, Optional

import numpy as np


class Trajectory(Sequence):
    """One episode as array views (observation [T,*obs], action [T], reward [T], policy [T,A], done [T]) that reads like
    the reference's list of step dicts (src/cumind/data/self_play.py:50-57): `len(t)`, `t[i]` -> dict, `t[:k]` -> list of
    dicts, iteration.  run_batch stores the steps of all environments step-major and hands out episodes as strided
    views, so the per-step host work does not grow with the number of environments."""

    __slots__ = ("_bufs", "_s0", "_s1", "_env")

    def __init__(self, bufs, s0: int, s1: int, env: int):
        # bufs = the step-major arrays (observation, action, reward, policy, done) of the whole run; the views are cut on
        # access, so ending an episode costs one small object (hundreds of episodes end per lock-step early in training)
        self._bufs, self._s0, self._s1, self._env = bufs, s0, s1, env

    observation = property(lambda self: self._bufs[0][self._s0:self._s1, self._env])
    action = property(lambda self: self._bufs[1][self._s0:self._s1, self._env])
    reward = property(lambda self: self._bufs[2][self._s0:self._s1, self._env])
    policy = property(lambda self: self._bufs[3][self._s0:self._s1, self._env])
    done = property(lambda self: self._bufs[4][self._s0:self._s1, self._env])

    def __len__(self) -> int:
        return self._s1 - self._s0

    def _step(self, i: int) -> Dict[str, Any]:
        b, t, e = self._bufs, self._s0 + i, self._env
        return {"observation": b[0][t, e], "action": int(b[1][t, e]), "reward": float(b[2][t, e]), "policy": b[3][t, e], "done": bool(b[4][t, e])}

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._step(j) for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self._step(i)


class SelfPlay:
    """`SelfPlay(config, agent, memory)` as the reference builds it (self_play.py:14-25); `SelfPlay(agent, memory=None)` is
    accepted too (the config is then the agent's)."""

    def __init__(self, *args, **kwargs):
        names = ("config", "agent", "memory")
        if args and (hasattr(args[0], "select_action") or hasattr(args[0], "select_actions")):   # SelfPlay(agent[, memory])
            names = ("agent", "memory")
        bound = dict(zip(names, args))
        for k in kwargs:
            if k not in ("config", "agent", "memory") or k in bound:
                raise TypeError(f"SelfPlay() got an unexpected or duplicate argument {k!r}")
        bound.update(kwargs)
        if "agent" not in bound:
            raise TypeError("SelfPlay() needs an agent")
        self.agent = bound["agent"]
        self.config = bound.get("config", getattr(self.agent, "config", None))
        self.memory = bound.get("memory")

    def run_episode(self, environment, max_steps: int = 10_000):
        """One episode of a gymnasium-style environment (reset() -> (obs, info); step(a) -> 5-tuple).  Returns the
        reference's tuple (total_reward, episode_steps, episode_data) — self_play.py:27-63."""
        episode: List[Dict[str, Any]] = []
        observation, _ = environment.reset()
        done = False
        total_reward = 0.0
        while not done and len(episode) < max_steps:
            action, policy = self.agent.select_action(np.asarray(observation), training=True)
            next_observation, reward, terminated, truncated, _ = environment.step(action)
            done = bool(terminated or truncated)
            episode.append({"observation": observation, "action": action, "reward": reward, "policy": policy, "done": done})
            observation = next_observation
            total_reward += reward
        if self.memory is not None:
            self.memory.add(episode)
        return total_reward, len(episode), episode

    def collect_samples(self, environment, num_episodes: int) -> None:
        """self_play.py:65-76"""
        for _ in range(num_episodes):
            self.run_episode(environment)

    def get_memory(self):
        """self_play.py:78-84"""
        return self.memory

    def run_batch(self, vec_env, num_steps: int, training: bool = True, philox_seed: Optional[int] = None) -> List[List[Trajectory]]:
        """Step `vec_env.num_envs` environments `num_steps` times.  Returns, per environment, the list of
        its finished episodes followed by the (possibly unfinished) current one, each a Trajectory (reads like the
        reference's list of step dicts).  Finished episodes go to `memory.add` as they end."""
        B = vec_env.num_envs
        observation, _ = vec_env.reset()
        observation = np.asarray(observation)
        A = int(self.agent.config.action_space_size)
        # step-major storage of everything the step record holds; episodes are slices [first step : last step + 1, env]
        obs_buf = np.empty((num_steps,) + observation.shape, dtype=observation.dtype)
        act_buf = np.empty((num_steps, B), dtype=np.int32)
        rew_buf = np.empty((num_steps, B), dtype=np.float32)
        pol_buf = np.empty((num_steps, B, A), dtype=np.float64)
        done_buf = np.zeros((num_steps, B), dtype=bool)
        start = np.zeros(B, dtype=np.int64)
        finished: List[List[Trajectory]] = [[] for _ in range(B)]

        bufs = (obs_buf, act_buf, rew_buf, pol_buf, done_buf)
        add_batch = getattr(self.memory, "add_batch", None) if self.memory is not None else None

        for step in range(num_steps):
            seed = None if philox_seed is None else philox_seed + step
            actions, policies = self.agent.select_actions(observation, training=training, philox_seed=seed)
            next_observation, reward, terminated, truncated, _ = vec_env.step(actions)
            done = np.logical_or(terminated, truncated)
            obs_buf[step], act_buf[step], rew_buf[step], pol_buf[step], done_buf[step] = observation, actions, reward, policies, done
            ended = np.nonzero(done)[0]                        # only the environments whose episode just ended
            if ended.size:
                envs = ended.tolist()
                eps = [Trajectory(bufs, s0, step + 1, i) for i, s0 in zip(envs, start[ended].tolist())]
                if add_batch is not None:
                    add_batch(eps)                             # one call (one device copy / sum-tree launch where those exist)
                elif self.memory is not None:
                    for ep in eps:
                        self.memory.add(ep)
                for i, ep in zip(envs, eps):
                    finished[i].append(ep)
                start[ended] = step + 1
            observation = np.asarray(next_observation)
        return [finished[i] + ([Trajectory(bufs, int(start[i]), num_steps, i)] if start[i] < num_steps else []) for i in range(B)]
