"""Vectorised CartPole + batched self-play driver."""

import numpy as np
import pytest

from cumind_b200.envs import CartPoleVec


def test_cartpole_vec_physics_and_auto_reset():
    env = CartPoleVec(64, seed=1)
    obs, _ = env.reset()
    assert obs.shape == (64, 4) and obs.dtype == np.float32 and np.all(np.abs(obs) <= 0.05)
    # one Euler step by hand for environment 0, action 1 (gymnasium CartPoleEnv.step)
    x, xd, th, thd = env.state[0]
    temp = (10.0 + 0.05 * thd**2 * np.sin(th)) / 1.1
    thacc = (9.8 * np.sin(th) - np.cos(th) * temp) / (0.5 * (4.0 / 3.0 - 0.1 * np.cos(th) ** 2 / 1.1))
    xacc = temp - 0.05 * thacc * np.cos(th) / 1.1
    want = np.array([x + 0.02 * xd, xd + 0.02 * xacc, th + 0.02 * thd, thd + 0.02 * thacc])
    obs, r, term, trunc, info = env.step(np.ones(64, dtype=np.int64))
    assert np.allclose(env.state[0], want) and np.all(r == 1.0) and not term.any()
    # pushing right forever must terminate every environment well before the 500-step truncation
    ended = np.zeros(64, bool)
    for _ in range(200):
        obs, r, term, trunc, info = env.step(np.ones(64, dtype=np.int64))
        ended |= term
        assert np.all(np.abs(obs[term]) <= 0.05)          # finished environments were reset
        assert np.all((np.abs(info["final_observation"][term, 0]) > 2.4) | (np.abs(info["final_observation"][term, 2]) > 12 * np.pi / 180))
    assert ended.all()


class _FakeAgent:
    """select_actions stand-in (no GPU): random policies, actions sampled from them."""

    class config:
        action_space_size = 2

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)

    def select_actions(self, obs, training=True, philox_seed=None):
        p = self.rng.dirichlet([1.0, 1.0], size=obs.shape[0])
        return (self.rng.random(obs.shape[0]) < p[:, 1]).astype(np.int32), p


def test_run_batch_bookkeeping_equals_the_per_environment_loop():
    """run_batch stores the steps step-major and hands out episodes as array views (Trajectory); the records, the
    episode boundaries and what reaches memory.add equal the reference-style loop that appends one dict per
    environment and step (src/cumind/data/self_play.py:44-61)."""
    from cumind_b200.self_play import SelfPlay, Trajectory

    class Mem:
        def __init__(self):
            self.items = []

        def add(self, x):
            self.items.append(x)

    B, T = 37, 120
    mem = Mem()
    got = SelfPlay(_FakeAgent(), mem).run_batch(CartPoleVec(B, seed=2), T)
    agent, env = _FakeAgent(), CartPoleVec(B, seed=2)
    obs, _ = env.reset()
    cur, fin = [[] for _ in range(B)], [[] for _ in range(B)]
    for _ in range(T):
        a, p = agent.select_actions(obs)
        nobs, r, te, tr, _ = env.step(a)
        d = te | tr
        for i in range(B):
            cur[i].append({"observation": obs[i], "action": int(a[i]), "reward": float(r[i]), "policy": p[i], "done": bool(d[i])})
            if d[i]:
                fin[i].append(cur[i])
                cur[i] = []
        obs = nobs
    want = [fin[i] + ([cur[i]] if cur[i] else []) for i in range(B)]
    n = 0
    for i in range(B):
        assert len(got[i]) == len(want[i])
        for e, w in zip(got[i], want[i]):
            assert isinstance(e, Trajectory) and len(e) == len(w)
            for s, t in zip(e, w):
                n += 1
                assert set(s) == {"observation", "action", "reward", "policy", "done"}
                assert (s["action"], s["reward"], s["done"]) == (t["action"], t["reward"], t["done"])
                assert np.array_equal(s["observation"], t["observation"]) and np.array_equal(s["policy"], t["policy"])
            assert [x["action"] for x in e[:5]] == [x["action"] for x in w[:5]] and e[-1]["done"] == w[-1]["done"]
            with pytest.raises(IndexError):
                e[len(e)]
    assert n == B * T and len(mem.items) == sum(len(f) for f in fin)


@pytest.mark.gpu
def test_batched_self_play_runs_the_fused_search_per_step():
    from cumind_b200 import Agent, Config
    from cumind_b200.self_play import SelfPlay

    cfg = Config(hidden_dim=32, num_simulations=8, action_space_size=2, observation_shape=(4,))
    agent = Agent(cfg)
    env = CartPoleVec(48, seed=3)
    episodes = SelfPlay(agent).run_batch(env, num_steps=30, training=True, philox_seed=11)
    assert len(episodes) == 48
    steps = 0
    for per_env in episodes:
        for ep in per_env:
            for s in ep:
                assert set(s) == {"observation", "action", "reward", "policy", "done"}      # self_play.py:50-57
                assert s["action"] in (0, 1) and s["policy"].shape == (2,) and abs(s["policy"].sum() - 1.0) < 1e-12
                steps += 1
            assert all(not s["done"] for s in ep[:-1])
    assert steps == 48 * 30
    total, steps, one = SelfPlay(agent).run_episode(_Single(), max_steps=12)
    assert steps == len(one) and total == sum(s["reward"] for s in one)
    assert 1 <= len(one) <= 12 and one[-1]["done"] in (True, False)


class _Single:
    """gymnasium-style single environment view over CartPoleVec(1)."""

    def __init__(self):
        self.env = CartPoleVec(1, seed=5)

    def reset(self):
        obs, info = self.env.reset()
        return obs[0], info

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(np.array([action]))
        return obs[0], float(r[0]), bool(term[0]), bool(trunc[0]), info
