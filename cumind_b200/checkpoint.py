"""Checkpoint ingest: read the reference's pickle without JAX/flax.

The reference saves `pickle.dump({"network_state": nnx.state(net), "optimizer_state": opt_state})`
(src/cumind/agent/agent.py:97-107, src/cumind/utils/checkpoint.py:13-28).  Unpickling that normally
imports flax, jax and optax; here every non-NumPy class is replaced by a stub and the arrays are
rebuilt through `numpy._core.multiarray._reconstruct`, which is how `jax.Array.__reduce__` stores them
(`jax._src.array._reconstruct_array(fun, args, arr_state, aval_state)`).  The result is the plain
nested dict of float32 arrays that `CuMindNetwork.load_state` / `Agent.load_state` take.

Writing: `save_checkpoint` pickles the same nested dict of NumPy arrays (no flax classes, so the
reference cannot read it back without a small adapter; `load_checkpoint` reads both formats).
"""

from __future__ import annotations

import os
import pickle
from typing import Any, Dict

import numpy as np


class _Stub:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def __setstate__(self, state):
        self.state = state


class _ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy") or module == "builtins" or module == "collections":
            return super().find_class(module, name)
        if module == "jax._src.array" and name == "_reconstruct_array":
            def rebuild(fun, args, arr_state, aval_state):
                arr = fun(*args)
                arr.__setstate__(arr_state)
                return arr
            return rebuild
        return type(name, (_Stub,), {"__module__": module})


def _plain(node: Any) -> Any:
    """nnx.State / VariableState stubs -> nested dict of float32 arrays (rng bookkeeping dropped)."""
    if isinstance(node, dict):
        return {k: _plain(v) for k, v in node.items() if k != "rngs"}
    if isinstance(node, _Stub):
        st = getattr(node, "state", None)
        if isinstance(st, dict) and "_mapping" in st:          # nnx.State
            return _plain(st["_mapping"])
        if isinstance(st, tuple) and len(st) == 2 and isinstance(st[1], dict) and "value" in st[1]:  # VariableState
            return np.asarray(st[1]["value"], dtype=np.float32)
        raise TypeError(f"unsupported object in checkpoint: {type(node).__name__}")
    if isinstance(node, np.ndarray):
        return np.asarray(node, dtype=np.float32)
    return node


def _adam_state(opt: Any) -> Dict[str, Any]:
    """optax (ScaleByAdamState(count, mu, nu), EmptyState, EmptyState) -> {"count", "mu", "nu"} trees."""
    try:
        adam = opt[0]
        fields = adam.kwargs or {}
        args = adam.args
        count = fields.get("count", args[0] if args else 0)
        mu = fields.get("mu", args[1] if len(args) > 1 else None)
        nu = fields.get("nu", args[2] if len(args) > 2 else None)
        return {"count": int(np.asarray(count)) if count is not None else 0,
                "mu": _plain(mu) if mu is not None else None, "nu": _plain(nu) if nu is not None else None}
    except Exception:
        return {"count": 0, "mu": None, "nu": None}


def load_reference_checkpoint(path: str) -> Dict[str, Any]:
    """{"network_state": parameter tree, "optimizer_state": {...}} from a reference .pkl file."""
    with open(path, "rb") as f:
        obj = _ReferenceUnpickler(f).load()
    if not isinstance(obj, dict) or "network_state" not in obj:
        raise ValueError(f"{path} is not a CuMind checkpoint")
    return {"network_state": _plain(obj["network_state"]), "optimizer_state": _adam_state(obj.get("optimizer_state"))}


def save_checkpoint(state: Dict[str, Any], path: str) -> None:
    """Pickle {"network_state": tree, "optimizer_state": ...} as plain dicts of NumPy arrays."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({"format": "cumind_b200/1", **state}, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_checkpoint(path: str) -> Dict[str, Any]:
    """Reads either this package's checkpoints or the reference's (checkpoint.py:31-47)."""
    with open(path, "rb") as f:
        head = f.read(4096)
    if b"cumind_b200/1" in head:
        with open(path, "rb") as f:
            obj = pickle.load(f)
        obj.pop("format", None)
        return obj
    return load_reference_checkpoint(path)


def infer_dims(tree: Dict[str, Any]) -> Dict[str, Any]:
    """hidden_dim / num_blocks / action_space_size / observation_shape[-1] implied by a parameter tree."""
    dyn = tree["dynamics_network"]
    A, H = dyn["action_embedding"]["embedding"].shape
    enc = tree["representation_network"]["encoder"]
    out = {"hidden_dim": int(H), "action_space_size": int(A), "num_blocks": len(dyn["layers"])}
    if "layers" in enc:
        first = enc["layers"][0] if 0 in enc["layers"] else enc["layers"]["0"]
        out["observation_shape"] = (int(first["kernel"].shape[0]),)
    else:
        out["conv_channels"] = int(enc["initial_conv"]["kernel"].shape[3])
        out["input_channels"] = int(enc["initial_conv"]["kernel"].shape[2])
    return out
