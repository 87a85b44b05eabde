"""Parameter tree <-> flat blob, and the initialisers.

Tree layout = the reference's nnx state (read out of the shipped checkpoint):
  representation_network/encoder/layers/{i}/{kernel,bias}                      (vector observations)
  representation_network/encoder/{initial_conv, residual_blocks/{i}/{conv1,bn1,conv2,bn2}, final_dense}
  dynamics_network/{action_embedding/embedding, layers/{i}/{kernel,bias}, reward_head/{kernel,bias}}
  prediction_network/{policy_head,value_head}/{kernel,bias}
Blob order = include/cumind_b200.h.  Initialisers follow flax defaults in distribution (lecun-normal
kernels, zero biases, N(0, 1/H) embedding, BatchNorm scale 1 / bias 0 / mean 0 / var 1) but not in
bit stream (JAX threefry is unavailable).
"""

from __future__ import annotations

from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

f32 = np.float32
_BN_KEYS = ("scale", "bias", "mean", "var")


def _schema(observation_shape: Sequence[int], A: int, H: int, L: int, Cc: int) -> List[Tuple[Tuple[Any, ...], Tuple[int, ...]]]:
    """Ordered list of (path, shape) — the single source of truth for the blob order."""
    shape = tuple(int(s) for s in observation_shape)
    out: List[Tuple[Tuple[Any, ...], Tuple[int, ...]]] = []
    enc = ("representation_network", "encoder")
    if len(shape) == 1:
        cur = shape[0]
        for i in range(L):
            out += [(enc + ("layers", i, "kernel"), (cur, H)), (enc + ("layers", i, "bias"), (H,))]
            cur = H
    elif len(shape) == 3:
        C = shape[2]
        out += [(enc + ("initial_conv", "kernel"), (3, 3, C, Cc)), (enc + ("initial_conv", "bias"), (Cc,))]
        for i in range(L):
            blk = enc + ("residual_blocks", i)
            for conv, bn in (("conv1", "bn1"), ("conv2", "bn2")):
                out += [(blk + (conv, "kernel"), (3, 3, Cc, Cc)), (blk + (conv, "bias"), (Cc,))]
                out += [(blk + (bn, k), (Cc,)) for k in _BN_KEYS]
        out += [(enc + ("final_dense", "kernel"), (Cc, H)), (enc + ("final_dense", "bias"), (H,))]
    else:
        raise ValueError(f"Unsupported observation shape: {shape}")
    dyn = ("dynamics_network",)
    out.append((dyn + ("action_embedding", "embedding"), (A, H)))
    for i in range(L):
        out += [(dyn + ("layers", i, "kernel"), (H, H)), (dyn + ("layers", i, "bias"), (H,))]
    out += [(dyn + ("reward_head", "kernel"), (H, 1)), (dyn + ("reward_head", "bias"), (1,))]
    pred = ("prediction_network",)
    out += [(pred + ("policy_head", "kernel"), (H, A)), (pred + ("policy_head", "bias"), (A,)),
            (pred + ("value_head", "kernel"), (H, 1)), (pred + ("value_head", "bias"), (1,))]
    return out


def _get(tree: Dict[Any, Any], path: Tuple[Any, ...]):
    node = tree
    for k in path:
        if k not in node and isinstance(k, int) and str(k) in node:
            k = str(k)
        node = node[k]
    return node


def _set(tree: Dict[Any, Any], path: Tuple[Any, ...], value) -> None:
    node = tree
    for k in path[:-1]:
        node = node.setdefault(k, {})
    node[path[-1]] = value


def dims_of(tree: Dict[str, Any]):
    """(observation_shape, A, H, L, Cc) implied by a parameter tree; for image encoders only the channel count of the
    observation is recoverable, which is all the blob layout depends on (height and width are reported as 0)."""
    dyn = tree["dynamics_network"]
    A, H = (int(v) for v in np.shape(_get(dyn, ("action_embedding", "embedding"))))
    L = len(dyn["layers"])
    enc = tree["representation_network"]["encoder"]
    if "layers" in enc:
        return (int(np.shape(_get(enc, ("layers", 0, "kernel")))[0]),), A, H, L, 32
    k = np.shape(_get(enc, ("initial_conv", "kernel")))
    return (0, 0, int(k[2])), A, H, L, int(k[3])


def param_count(observation_shape, A, H, L, Cc) -> int:
    return int(sum(int(np.prod(s)) for _, s in _schema(observation_shape, A, H, L, Cc)))


def flatten(tree: Dict[str, Any], observation_shape, A, H, L, Cc) -> np.ndarray:
    parts = []
    for path, shape in _schema(observation_shape, A, H, L, Cc):
        a = np.asarray(_get(tree, path), dtype=f32)
        if a.shape != shape:
            raise ValueError(f"parameter {'/'.join(map(str, path))} has shape {a.shape}, expected {shape}")
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def unflatten(blob: np.ndarray, observation_shape, A, H, L, Cc) -> Dict[str, Any]:
    blob = np.asarray(blob, dtype=f32).reshape(-1)
    tree: Dict[str, Any] = {}
    o = 0
    for path, shape in _schema(observation_shape, A, H, L, Cc):
        n = int(np.prod(shape))
        _set(tree, path, blob[o:o + n].reshape(shape).copy())
        o += n
    if o != blob.size:
        raise ValueError(f"parameter blob has {blob.size} values, expected {o}")
    return tree


def init(observation_shape, A, H, L, Cc, rng: np.random.Generator) -> Dict[str, Any]:
    tree: Dict[str, Any] = {}
    for path, shape in _schema(observation_shape, A, H, L, Cc):
        leaf = path[-1]
        if leaf == "kernel":
            fan_in = int(np.prod(shape[:-1]))
            # lecun_normal: truncated normal (+-2 sigma), variance 1/fan_in
            z = rng.standard_normal(shape)
            bad = np.abs(z) > 2.0
            while bad.any():
                z[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(z) > 2.0
            val = (z / 0.87962566103423978 / np.sqrt(fan_in)).astype(f32)
        elif leaf == "embedding":
            val = (rng.standard_normal(shape) / np.sqrt(shape[1])).astype(f32)
        elif leaf in ("scale", "var"):
            val = np.ones(shape, f32)
        else:  # bias, mean
            val = np.zeros(shape, f32)
        _set(tree, path, val)
    return tree
