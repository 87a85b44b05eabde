"""Vectorised CartPole-v1 in NumPy (gymnasium is not available in this image).

Same physics constants, Euler integration, termination / truncation rules and reset distribution as
gymnasium's `CartPoleEnv` (the environment the reference trains on, src/cumind/runner.py:18,
configuration.json `env_name`), for `num_envs` independent environments stepped in lockstep — the
shape of work `Agent.select_actions` consumes.  Finished environments reset automatically.
"""

from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


class CartPoleVec:
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    theta_threshold = 12 * 2 * np.pi / 360
    x_threshold = 2.4
    max_episode_steps = 500
    observation_shape = (4,)
    action_space_size = 2

    def __init__(self, num_envs: int, seed: int = 0):
        self.num_envs = int(num_envs)
        self.rng = np.random.default_rng(seed)
        self.state = np.zeros((self.num_envs, 4), dtype=np.float64)
        self.steps = np.zeros(self.num_envs, dtype=np.int64)
        self.reset()

    def _sample(self, n: int) -> np.ndarray:
        return self.rng.uniform(-0.05, 0.05, size=(n, 4))

    def reset(self) -> Tuple[np.ndarray, Dict]:
        self.state[:] = self._sample(self.num_envs)
        self.steps[:] = 0
        return self.state.astype(np.float32), {}

    def step(self, actions: np.ndarray):
        """actions [num_envs] in {0,1} -> (obs [B,4] float32, reward [B], terminated [B], truncated [B], info).
        The returned observation of a finished environment is already the first one of its next episode;
        info["final_observation"] holds the terminal observations."""
        a = np.asarray(actions).reshape(self.num_envs)
        x, x_dot, theta, theta_dot = self.state.T
        force = np.where(a == 1, self.force_mag, -self.force_mag)
        costheta, sintheta = np.cos(theta), np.sin(theta)
        total_mass = self.masspole + self.masscart
        polemass_length = self.masspole * self.length
        temp = (force + polemass_length * theta_dot**2 * sintheta) / total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (self.length * (4.0 / 3.0 - self.masspole * costheta**2 / total_mass))
        xacc = temp - polemass_length * thetaacc * costheta / total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = np.stack([x, x_dot, theta, theta_dot], axis=1)
        self.steps += 1
        terminated = (np.abs(x) > self.x_threshold) | (np.abs(theta) > self.theta_threshold)
        truncated = (self.steps >= self.max_episode_steps) & ~terminated
        reward = np.ones(self.num_envs, dtype=np.float32)
        final_obs = self.state.astype(np.float32)
        done = terminated | truncated
        if done.any():
            idx = np.nonzero(done)[0]
            self.state[idx] = self._sample(len(idx))
            self.steps[idx] = 0
        return self.state.astype(np.float32), reward, terminated, truncated, {"final_observation": final_obs}
