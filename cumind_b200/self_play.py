"""Self-play on top of the batched search.

`SelfPlay.run_episode(env)` keeps the reference's single-environment loop and step record
(src/cumind/data/self_play.py:27-63: one `select_action(training=True)` per step, dicts with
observation / action / reward / policy / done).  `SelfPlay.run_batch(vec_env, num_steps)` is the
B200-shaped version: B environments advance in lockstep, one `Agent.select_actions` (one fused search
over B trees) per step, so the host loop no longer starves the GPU.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np


class SelfPlay:
    def __init__(self, agent, memory: Optional[Any] = None):
        self.agent = agent
        self.memory = memory

    def run_episode(self, environment, max_steps: int = 10_000) -> List[Dict[str, Any]]:
        """One episode of a gymnasium-style environment (reset() -> (obs, info); step(a) -> 5-tuple)."""
        episode: List[Dict[str, Any]] = []
        observation, _ = environment.reset()
        done = False
        while not done and len(episode) < max_steps:
            action, policy = self.agent.select_action(np.asarray(observation), training=True)
            next_observation, reward, terminated, truncated, _ = environment.step(action)
            done = bool(terminated or truncated)
            episode.append({"observation": observation, "action": action, "reward": reward, "policy": policy, "done": done})
            observation = next_observation
        if self.memory is not None:
            self.memory.add(episode)
        return episode

    def run_batch(self, vec_env, num_steps: int, training: bool = True, philox_seed: Optional[int] = None) -> List[List[List[Dict[str, Any]]]]:
        """Step `vec_env.num_envs` environments `num_steps` times.  Returns, per environment, the list of
        its finished episodes followed by the (possibly unfinished) current one, each a list of step dicts."""
        B = vec_env.num_envs
        finished: List[List[List[Dict[str, Any]]]] = [[] for _ in range(B)]
        current: List[List[Dict[str, Any]]] = [[] for _ in range(B)]
        observation, _ = vec_env.reset()
        for step in range(num_steps):
            seed = None if philox_seed is None else philox_seed + step
            actions, policies = self.agent.select_actions(observation, training=training, philox_seed=seed)
            next_observation, reward, terminated, truncated, _ = vec_env.step(actions)
            done = np.logical_or(terminated, truncated)
            for i in range(B):
                current[i].append({"observation": observation[i], "action": int(actions[i]), "reward": float(reward[i]),
                                   "policy": policies[i], "done": bool(done[i])})
                if done[i]:
                    if self.memory is not None:
                        self.memory.add(current[i])
                    finished[i].append(current[i])
                    current[i] = []
            observation = next_observation
        return [finished[i] + ([current[i]] if current[i] else []) for i in range(B)]
