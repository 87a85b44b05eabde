"""Shared loaders for the committed golden vectors (tests/golden/*.npz)."""

from __future__ import annotations

import json
import os
from functools import lru_cache
from typing import Any, Dict, List

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CANON_KEYS = ("path_len", "paths", "visit", "value_sum", "prior", "expanded")


@lru_cache(maxsize=None)
def _npz(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def search_cases() -> List[Dict[str, Any]]:
    z = _npz("search_cases.npz")
    return json.loads(str(z["manifest"]))


def search_case_data(case: Dict[str, Any]) -> Dict[str, Any]:
    z = _npz("search_cases.npz")
    pre = case["id"]
    d = {"canon": {k: z[f"{pre}_{k}"] for k in _CANON_KEYS}, "probs": z[f"{pre}_probs"], "root_h": z[f"{pre}_root_h"],
         "obs": z[f"{pre}_obs"], "noise": z[f"{pre}_noise"] if case["noise"] else None}
    d["params"] = ckpt_params() if case["net"] == "ckpt" else z[case["net"]]
    return d


def ckpt_params() -> np.ndarray:
    return _npz("cartpole_ckpt.npz")["params"]


def ckpt_config() -> Dict[str, Any]:
    return json.loads(str(_npz("cartpole_ckpt.npz")["config"]))


def network_cases() -> List[Dict[str, Any]]:
    z = _npz("network_cases.npz")
    return json.loads(str(z["manifest"]))


def network_case_data(case: Dict[str, Any]) -> Dict[str, np.ndarray]:
    z = _npz("network_cases.npz")
    pre = case["id"]
    keys = ("params", "obs", "h", "logits", "value", "action", "next", "reward", "logits2", "value2", "probs2")
    return {k: z[f"{pre}_{k}"] for k in keys}


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-exact equality of float arrays, except that any NaN equals any NaN (the CPU and the GPU
    emit different canonical NaN payloads)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind != "f":
        return bool(np.array_equal(a, b))
    u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    same = a.view(u) == b.view(u)
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(same | both_nan))
