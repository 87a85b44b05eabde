"""TEST INFRASTRUCTURE — run the reference's own mcts.py, unmodified, without JAX.

/root/reference/src/cumind/core/mcts.py is loaded from where it lies (never copied) and executed
with stand-in modules for the imports it needs and this image lacks:
    chex.Array; jax.numpy.{asarray, array, expand_dims, float32, int32}; jax.nn.softmax;
    cumind.utils.logger.log (no-op).
(mcts.py:3-11 are its imports; 87-88, 126-128, 181-182, 189-193 the call sites.)  The network it
searches over is oracle/network_np.NumpyCuMindNetwork, and jax.nn.softmax is the pinned softmax
of that module, so the tree logic (Node.ucb_score / select_child / expand / backup, MCTS.search /
_simulate / _add_exploration_noise) is the reference's own code, bit for bit.

This module only works where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to produce the committed golden vectors, and by the container-only
tests; nothing that runs on the GPU box imports it.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import network_np

REFERENCE_ROOT = os.environ.get("CUMIND_REFERENCE_ROOT", "/root/reference")
_MCTS_PATH = os.path.join(REFERENCE_ROOT, "src", "cumind", "core", "mcts.py")

_loaded = None


def available() -> bool:
    return os.path.isfile(_MCTS_PATH)


class _NoLog:
    def __getattr__(self, name):
        return lambda *a, **k: None


def _install_stubs() -> Dict[str, Any]:
    saved = {}

    def put(name, mod):
        saved[name] = sys.modules.get(name)
        sys.modules[name] = mod

    chex = types.ModuleType("chex")
    chex.Array = np.ndarray
    put("chex", chex)

    jnp = types.ModuleType("jax.numpy")
    jnp.asarray = lambda x, dtype=None: np.asarray(x, dtype=dtype)
    jnp.array = lambda x, dtype=None: np.array(x, dtype=dtype)
    jnp.expand_dims = np.expand_dims
    jnp.float32 = np.float32
    jnp.int32 = np.int32
    jnn = types.ModuleType("jax.nn")
    jnn.softmax = lambda x, axis=-1: network_np.softmax(x)
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.nn = jnn
    put("jax", jax)
    put("jax.numpy", jnp)
    put("jax.nn", jnn)

    pkg = types.ModuleType("cumind")
    pkg.__path__ = []
    core = types.ModuleType("cumind.core")
    core.__path__ = []
    utils = types.ModuleType("cumind.utils")
    utils.__path__ = []
    logger = types.ModuleType("cumind.utils.logger")
    logger.log = _NoLog()
    put("cumind", pkg)
    put("cumind.core", core)
    put("cumind.utils", utils)
    put("cumind.utils.logger", logger)
    return saved


def _restore(saved: Dict[str, Any]) -> None:
    for name, mod in saved.items():
        if mod is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = mod


def load():
    """Return the reference module object (classes Node, MCTS), executing mcts.py verbatim."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError(f"reference not present at {_MCTS_PATH}")
    saved = _install_stubs()
    try:
        spec = importlib.util.spec_from_file_location("cumind.core.mcts", _MCTS_PATH)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "cumind.core"
        sys.modules["cumind.core.mcts"] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop("cumind.core.mcts", None)
        _restore(saved)
    _loaded = mod
    return mod


class SearchConfig:
    """The five fields MCTS reads from Config (mcts.py:132,139,169,218,222)."""

    def __init__(self, num_simulations=10, action_space_size=2, c_puct=1.0, dirichlet_alpha=0.3, exploration_fraction=0.25):
        self.num_simulations = num_simulations
        self.action_space_size = action_space_size
        self.c_puct = c_puct
        self.dirichlet_alpha = dirichlet_alpha
        self.exploration_fraction = exploration_fraction


def search_with_tree(network, config, root_hidden: np.ndarray, noise: Optional[np.ndarray]) -> Tuple[np.ndarray, Any]:
    """MCTS.search (mcts.py:115-155) that also hands back the root Node.

    `noise`, when given, is what np.random.dirichlet returns inside _add_exploration_noise
    (mcts.py:218): the global-RNG draw is the only thing replaced, the mixing code is the
    reference's.  The root is captured by wrapping _simulate (called with it, mcts.py:140).
    """
    ref = load()
    mcts = ref.MCTS(network, config)
    captured: List[Any] = []
    orig_sim = mcts._simulate

    def sim(root):
        if not captured:
            captured.append(root)
        return orig_sim(root)

    mcts._simulate = sim
    orig_dir = np.random.dirichlet
    if noise is not None:
        np.random.dirichlet = lambda alpha, size=None: np.asarray(noise, dtype=np.float64)
    try:
        probs = mcts.search(np.asarray(root_hidden, dtype=np.float32), add_noise=noise is not None)
    finally:
        np.random.dirichlet = orig_dir
    return probs, (captured[0] if captured else None)
