"""Checkpoint ingest: read the reference's pickle without JAX/flax.

The reference saves `pickle.dump({"network_state": nnx.state(net), "optimizer_state": opt_state})`
(src/cumind/agent/agent.py:97-107, src/cumind/utils/checkpoint.py:13-28).  Unpickling that normally
imports flax, jax and optax; here every non-NumPy class is replaced by a stub and the arrays are
rebuilt through `numpy._core.multiarray._reconstruct`, which is how `jax.Array.__reduce__` stores them
(`jax._src.array._reconstruct_array(fun, args, arr_state, aval_state)`).  The result is the plain
nested dict of float32 arrays that `CuMindNetwork.load_state` / `Agent.load_state` take.

Writing: `save_checkpoint` pickles the same nested dict of NumPy arrays (no flax classes, so the
reference cannot read it back without a small adapter; `load_checkpoint` reads both formats).
Both formats are read through a RESTRICTED unpickler (exact allow-list of containers, numbers and the NumPy array
reconstruction; the reference's flax / jax / optax classes become inert stubs; anything else is refused).
"""

from __future__ import annotations

import os
import pickle
from typing import Any, Dict

import numpy as np


class _Stub:
    """Stand-in for every flax / jax / optax class in a reference checkpoint.  Records how it was built: NamedTuples such
    as optax.ScaleByAdamState are pickled with NEWOBJ (`cls.__new__(cls, *fields)`, no __init__ call), dataclass-like
    objects with REDUCE + BUILD."""

    def __new__(cls, *args, **kwargs):
        obj = object.__new__(cls)
        obj.args, obj.kwargs, obj.state = args, kwargs, None
        return obj

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        self.state = state


# exactly what a checkpoint needs: containers, numbers and the NumPy array reconstruction; nothing callable beyond that
_ALLOWED = {
    ("builtins", n) for n in ("dict", "list", "tuple", "set", "frozenset", "int", "float", "complex", "bool", "bytes", "bytearray", "str", "slice")
} | {
    ("collections", "OrderedDict"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "scalar"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.numeric", "_frombuffer"), ("numpy.core.numeric", "_frombuffer"),
}
_STUBBED_PREFIXES = ("flax", "jax", "jaxlib", "optax", "chex", "cumind")


class _ReferenceUnpickler(pickle.Unpickler):
    """Restricted unpickler: names on the allow-list resolve to the real object, classes of the reference's own
    dependencies (flax / jax / optax / chex / cumind) become inert stubs, and anything else is refused — a checkpoint
    is data, and a pickle that asks for eval / os.system / ... must not get it."""

    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        if module == "jax._src.array" and name == "_reconstruct_array":
            def rebuild(fun, args, arr_state, aval_state):
                arr = fun(*args)
                arr.__setstate__(arr_state)
                return arr
            return rebuild
        if module.split(".")[0] in _STUBBED_PREFIXES:
            return type(name, (_Stub,), {"__module__": module})
        raise pickle.UnpicklingError(f"checkpoint refers to {module}.{name}, which is not allowed in a checkpoint")


def _plain(node: Any) -> Any:
    """nnx.State / VariableState stubs -> nested dict of float32 arrays (rng bookkeeping dropped)."""
    if isinstance(node, dict):
        return {k: _plain(v) for k, v in node.items() if k != "rngs"}
    if isinstance(node, _Stub):
        st = node.state
        if isinstance(st, dict) and "_mapping" in st:          # nnx.State
            return _plain(st["_mapping"])
        if isinstance(st, tuple) and len(st) == 2 and isinstance(st[1], dict) and "value" in st[1]:  # VariableState
            return np.asarray(st[1]["value"], dtype=np.float32)
        raise TypeError(f"unsupported object in checkpoint: {type(node).__name__}")
    if isinstance(node, np.ndarray):
        return np.asarray(node, dtype=np.float32)
    return node


def _flat_moment(tree: Any, dims) -> np.ndarray:
    """An AdamW moment tree (same structure as the parameters; BatchNorm statistics have no entry) -> blob layout."""
    from . import params as P

    out = np.zeros(P.param_count(*dims), dtype=np.float32)
    o = 0
    for path, shape in P._schema(*dims):
        n = int(np.prod(shape))
        node = tree
        try:
            for key in path:
                node = node[key] if key in node else node[str(key)]
            out[o:o + n] = np.asarray(node, dtype=np.float32).reshape(-1)
        except (KeyError, TypeError):
            if path[-1] not in ("mean", "var"):
                raise ValueError(f"optimizer state has no entry for {'/'.join(map(str, path))}")
        o += n
    return out


def _adam_state(opt: Any, network_tree: Dict[str, Any]) -> Dict[str, Any]:
    """optax state of adamw = (ScaleByAdamState(count, mu, nu), EmptyState, EmptyState) -> {"count", "mu", "nu"} with the
    moments as flat fp32 arrays in the blob layout (what Trainer and cumind_adamw_step use)."""
    if opt is None:
        return {"count": 0, "mu": None, "nu": None}
    if isinstance(opt, dict) and "count" in opt:           # this package's own format
        return opt
    adam = opt[0] if isinstance(opt, (tuple, list)) and len(opt) > 0 else opt
    if not isinstance(adam, _Stub):
        raise ValueError(f"unsupported optimizer state of type {type(adam).__name__}")
    fields = dict(adam.kwargs)
    for name, value in zip(("count", "mu", "nu"), adam.args):
        fields.setdefault(name, value)
    if isinstance(adam.state, dict):
        fields.update({k: v for k, v in adam.state.items() if k in ("count", "mu", "nu")})
    if not {"count", "mu", "nu"} <= set(fields):
        raise ValueError("optimizer state does not look like optax.ScaleByAdamState(count, mu, nu)")
    from . import params as P

    dims = P.dims_of(network_tree)
    return {"count": int(np.asarray(fields["count"])), "mu": _flat_moment(_plain(fields["mu"]), dims),
            "nu": _flat_moment(_plain(fields["nu"]), dims)}


def load_reference_checkpoint(path: str) -> Dict[str, Any]:
    """{"network_state": parameter tree, "optimizer_state": {...}} from a reference .pkl file."""
    with open(path, "rb") as f:
        obj = _ReferenceUnpickler(f).load()
    if not isinstance(obj, dict) or "network_state" not in obj:
        raise ValueError(f"{path} is not a CuMind checkpoint")
    net = _plain(obj["network_state"])
    return {"network_state": net, "optimizer_state": _adam_state(obj.get("optimizer_state"), net)}


def save_checkpoint(state: Dict[str, Any], path: str) -> None:
    """Pickle {"network_state": tree, "optimizer_state": ...} as plain dicts of NumPy arrays."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({"format": "cumind_b200/1", **state}, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_checkpoint(path: str) -> Dict[str, Any]:
    """Reads either this package's checkpoints or the reference's (checkpoint.py:31-47)."""
    with open(path, "rb") as f:
        obj = _ReferenceUnpickler(f).load()      # both formats go through the restricted unpickler
    if isinstance(obj, dict) and obj.get("format") == "cumind_b200/1":
        obj.pop("format", None)
        return obj
    if not isinstance(obj, dict) or "network_state" not in obj:
        raise ValueError(f"{path} is not a CuMind checkpoint")
    net = _plain(obj["network_state"])
    return {"network_state": net, "optimizer_state": _adam_state(obj.get("optimizer_state"), net)}


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
