"""TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libcumind_oracle.so (cumind_oracle.c)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
def _host_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma" in line
    except OSError:
        pass
    return False


_SO = os.path.join(_HERE, "_build", "libcumind_oracle_fma.so" if _host_has_fma() else "libcumind_oracle.so")
_lib = None


class NetDesc(C.Structure):
    _fields_ = [("obs_ndim", C.c_int32), ("obs_shape", C.c_int32 * 3), ("hidden_dim", C.c_int32),
                ("num_blocks", C.c_int32), ("action_space_size", C.c_int32), ("conv_channels", C.c_int32)]


def make_desc(observation_shape: Sequence[int], action_space_size: int, hidden_dim: int, num_blocks: int, conv_channels: int = 32) -> NetDesc:
    d = NetDesc()
    d.obs_ndim = len(observation_shape)
    for i in range(3):
        d.obs_shape[i] = int(observation_shape[i]) if i < len(observation_shape) else 0
    d.hidden_dim, d.num_blocks, d.action_space_size, d.conv_channels = hidden_dim, num_blocks, action_space_size, conv_channels
    return d


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cumind_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.cmo_param_count.restype = C.c_size_t
        _lib.cmo_pexp.restype = C.c_double
        _lib.cmo_pexp.argtypes = [C.c_double]
        _lib.cmo_bn_gain.restype = C.c_float
        _lib.cmo_bn_gain.argtypes = [C.c_float, C.c_float]
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def param_count(desc: NetDesc) -> int:
    return int(lib().cmo_param_count(C.byref(desc)))


def pexp(d: float) -> float:
    return float(lib().cmo_pexp(float(d)))


def fmaf(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    a, b, c = (np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float32), np.broadcast(a, b, c).shape)) for v in (a, b, c))
    out = np.empty_like(a)
    lib().cmo_fmaf_array(_p(a), _p(b), _p(c), _p(out), C.c_long(a.size))
    return out


def softmax(logits: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(logits, dtype=np.float32)
    out = np.empty_like(x)
    flat_in, flat_out = x.reshape(-1, x.shape[-1]), out.reshape(-1, x.shape[-1])
    for i in range(flat_in.shape[0]):
        lib().cmo_softmax(_p(flat_in[i]), C.c_int(x.shape[-1]), _p(flat_out[i]))
    return out


def initial_inference(desc: NetDesc, params: np.ndarray, obs: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    B, H, A = obs.shape[0], desc.hidden_dim, desc.action_space_size
    h, lg, v = np.empty((B, H), np.float32), np.empty((B, A), np.float32), np.empty((B,), np.float32)
    rc = lib().cmo_initial_inference(C.byref(desc), _p(params), _p(obs), C.c_int(B), _p(h), _p(lg), _p(v))
    assert rc == 0, rc
    return h, lg, v


def recurrent_inference(desc: NetDesc, params: np.ndarray, hidden: np.ndarray, action: np.ndarray):
    hidden = np.ascontiguousarray(hidden, dtype=np.float32)
    action = np.ascontiguousarray(action, dtype=np.int32)
    B, H, A = hidden.shape[0], desc.hidden_dim, desc.action_space_size
    nx, rw = np.empty((B, H), np.float32), np.empty((B,), np.float32)
    lg, v = np.empty((B, A), np.float32), np.empty((B,), np.float32)
    rc = lib().cmo_recurrent_inference(C.byref(desc), _p(params), _p(hidden), _p(action), C.c_int(B), _p(nx), _p(rw), _p(lg), _p(v))
    assert rc == 0, rc
    return nx, rw, lg, v


def prediction(desc: NetDesc, params: np.ndarray, hidden: np.ndarray):
    hidden = np.ascontiguousarray(hidden, dtype=np.float32)
    B, A = hidden.shape[0], desc.action_space_size
    lg, v = np.empty((B, A), np.float32), np.empty((B,), np.float32)
    rc = lib().cmo_prediction(C.byref(desc), _p(params), _p(hidden), C.c_int(B), _p(lg), _p(v))
    assert rc == 0, rc
    return lg, v


def search(desc: NetDesc, params: np.ndarray, root_hidden: np.ndarray, num_simulations: int, action_space_size: int,
           c_puct: float, exploration_fraction: float = 0.25, noise: Optional[np.ndarray] = None,
           s_max: Optional[int] = None) -> Dict[str, np.ndarray]:
    """cmo_search over B trees; returns the SoA arrays (same layout as include/cumind_b200.h) and outputs."""
    root_hidden = np.ascontiguousarray(root_hidden, dtype=np.float32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    B, H = root_hidden.shape
    S = int(num_simulations)
    S_max = S if s_max is None else int(s_max)
    A_cfg = int(action_space_size)
    A = min(A_cfg, desc.action_space_size)
    N = 1 + A * (S_max + 1)
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        assert noise.shape == (B, A)
    out = {
        "visit": np.empty((B, N), np.int32), "value_sum": np.empty((B, N), np.float64),
        "prior": np.empty((B, N), np.float32), "root_prior": np.empty((B, A), np.float64),
        "child_eidx": np.empty((B, N), np.int32), "exp_node": np.empty((B, S_max + 1), np.int32),
        "hidden": np.zeros((B, S_max + 1, H), np.float32),
        "out_visits": np.empty((B, A_cfg), np.int32), "out_probs": np.empty((B, A_cfg), np.float64),
        "out_root_value": np.empty((B,), np.float64), "depth_sum": np.empty((B,), np.int64),
    }
    rc = lib().cmo_search(C.byref(desc), _p(params), _p(root_hidden), C.c_int(B), C.c_int(S), C.c_int(S_max), C.c_int(A_cfg),
                          C.c_double(c_puct), C.c_double(exploration_fraction), _p(noise),
                          _p(out["visit"]), _p(out["value_sum"]), _p(out["prior"]), _p(out["root_prior"]),
                          _p(out["child_eidx"]), _p(out["exp_node"]), _p(out["hidden"]),
                          _p(out["out_visits"]), _p(out["out_probs"]), _p(out["out_root_value"]), _p(out["depth_sum"]))
    assert rc == 0, rc
    out["A"] = A
    return out


def search_minmax(desc: NetDesc, params: np.ndarray, root_hidden: np.ndarray, num_simulations: int, action_space_size: int,
                  c_puct: float, discount: float, exploration_fraction: float = 0.25, noise: Optional[np.ndarray] = None,
                  s_max: Optional[int] = None) -> Dict[str, np.ndarray]:
    """cmo_search_mm: the OPTIONAL min-max-normalised / discounted tree mode (not the reference's semantics; defined in
    cumind_oracle.c).  Same arrays as search() plus reward [B,N] float32 and minmax [B,2] float64."""
    root_hidden = np.ascontiguousarray(root_hidden, dtype=np.float32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    B, H = root_hidden.shape
    S = int(num_simulations)
    S_max = S if s_max is None else int(s_max)
    A_cfg = int(action_space_size)
    A = min(A_cfg, desc.action_space_size)
    N = 1 + A * (S_max + 1)
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        assert noise.shape == (B, A)
    out = {
        "visit": np.empty((B, N), np.int32), "value_sum": np.empty((B, N), np.float64),
        "prior": np.empty((B, N), np.float32), "root_prior": np.empty((B, A), np.float64),
        "reward": np.empty((B, N), np.float32), "minmax": np.empty((B, 2), np.float64),
        "child_eidx": np.empty((B, N), np.int32), "exp_node": np.empty((B, S_max + 1), np.int32),
        "hidden": np.zeros((B, S_max + 1, H), np.float32),
        "out_visits": np.empty((B, A_cfg), np.int32), "out_probs": np.empty((B, A_cfg), np.float64),
        "out_root_value": np.empty((B,), np.float64), "depth_sum": np.empty((B,), np.int64),
    }
    rc = lib().cmo_search_mm(C.byref(desc), _p(params), _p(root_hidden), C.c_int(B), C.c_int(S), C.c_int(S_max), C.c_int(A_cfg),
                             C.c_double(c_puct), C.c_double(discount), C.c_double(exploration_fraction), _p(noise),
                             _p(out["visit"]), _p(out["value_sum"]), _p(out["prior"]), _p(out["root_prior"]),
                             _p(out["reward"]), _p(out["minmax"]), _p(out["child_eidx"]), _p(out["exp_node"]), _p(out["hidden"]),
                             _p(out["out_visits"]), _p(out["out_probs"]), _p(out["out_root_value"]), _p(out["depth_sum"]))
    assert rc == 0, rc
    out["A"] = A
    return out
