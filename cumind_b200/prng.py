"""Minimal stand-ins for the reference's PRNG plumbing (src/cumind/utils/prng.py:12-160).

The reference splits a JAX threefry key behind a process-wide singleton `key` and hands
`nnx.Rngs(params=key())` to the network constructor (agent.py:38-39).  JAX is not available here and
parameter initialisation is outside the search hot path, so `key` hands out NumPy SeedSequence
children instead: same call pattern (`key.seed(s)`, `key()`, `key.reset()`), different bit stream.
"""

from __future__ import annotations

import threading
from typing import Optional

import numpy as np


class _KeyManager:
    def __init__(self):
        self._lock = threading.RLock()
        self._seq: Optional[np.random.SeedSequence] = None

    def seed(self, seed: int) -> None:
        with self._lock:
            self._seq = np.random.SeedSequence(int(seed))

    def reset(self) -> None:
        with self._lock:
            self._seq = None

    def __call__(self, num: int = 1):
        with self._lock:
            if self._seq is None:
                raise RuntimeError("PRNG not initialized. Call key.seed(seed) first.")  # prng.py:102
            kids = self._seq.spawn(int(num))
        return kids[0] if num == 1 else kids


key = _KeyManager()


class Rngs:
    """Shape of flax.nnx.Rngs(params=...) as the reference uses it (agent.py:39, tests/test_mcts.py:190)."""

    def __init__(self, params=None, **kw):
        self.params = params if params is not None else kw.get("default", 0)

    def generator(self) -> np.random.Generator:
        return np.random.default_rng(self.params)
