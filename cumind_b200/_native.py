"""ctypes binding of the C ABI in include/cumind_b200.h (cumind_b200/_lib/libcumind_b200.so).

There is no CPU fallback: if the CUDA library is missing this module raises, and every entry point
that computes needs a CUDA device.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _build

_lib = None

EXPORTS = [
    "cumind_last_error", "cumind_abi_version", "cumind_param_count", "cumind_net_create", "cumind_net_update_params",
    "cumind_net_destroy", "cumind_initial_inference", "cumind_recurrent_inference", "cumind_prediction", "cumind_softmax",
    "cumind_tree_bytes", "cumind_tree_create", "cumind_tree_destroy", "cumind_tree_get_view", "cumind_search",
    "cumind_search_plan", "cumind_search_phase_clocks", "cumind_search_stepwise", "cumind_select_actions_work_bytes", "cumind_select_actions_host",
    "cumind_select_actions_host_u8", "cumind_replay_create", "cumind_replay_destroy", "cumind_replay_record_bytes", "cumind_replay_write",
    "cumind_replay_gather", "cumind_replay_values",
    "cumind_tree_init_root", "cumind_select", "cumind_expand", "cumind_backup", "cumind_finalize", "cumind_dirichlet",
    "cumind_launch_count", "cumind_selftest_division", "cumind_adamw_step", "cumind_adamw_step_dev", "cumind_net_update_params_device",
    "cumind_sumtree_create", "cumind_sumtree_destroy", "cumind_sumtree_tree_size", "cumind_sumtree_clear", "cumind_sumtree_update",
    "cumind_sumtree_sample", "cumind_sumtree_read", "cumind_search_from_obs", "cumind_net_reserve",
    "cumind_train_supported", "cumind_train_work_bytes", "cumind_train_loss_grad",
    "cumind_p2p_create", "cumind_p2p_open", "cumind_p2p_grad_ptr", "cumind_p2p_allreduce_adamw", "cumind_p2p_error", "cumind_p2p_destroy", "cumind_p2p_reset",
]

NET_FP32_EXACT = 0
NET_BF16_TC = 1
TREE_REFERENCE, TREE_MINMAX_DISCOUNT = 0, 1
SCHED_AUTO, SCHED_WARP, SCHED_TEAM, SCHED_TEAM_CLASSIC, SCHED_TENSOR = 0, 1, 2, 3, 4


class NetDesc(C.Structure):
    _fields_ = [("obs_ndim", C.c_int32), ("obs_shape", C.c_int32 * 3), ("hidden_dim", C.c_int32), ("num_blocks", C.c_int32),
                ("action_space_size", C.c_int32), ("conv_channels", C.c_int32)]


class SearchCfg(C.Structure):
    _fields_ = [("num_simulations", C.c_int32), ("action_space_size", C.c_int32), ("c_puct", C.c_double),
                ("exploration_fraction", C.c_double), ("dirichlet_alpha", C.c_double), ("noise_mode", C.c_int32),
                ("net_mode", C.c_int32), ("tree_mode", C.c_int32), ("lanes_per_tree", C.c_int32), ("seed", C.c_uint64),
                ("tree_index_base", C.c_uint64), ("schedule", C.c_int32), ("teams_per_cta", C.c_int32),
                ("net_warps_per_team", C.c_int32), ("trees_per_team", C.c_int32), ("discount", C.c_double)]


class TreeView(C.Structure):
    _fields_ = [("B", C.c_int32), ("S_max", C.c_int32), ("A", C.c_int32), ("H", C.c_int32), ("N", C.c_int32),
                ("nodes", C.c_void_p), ("child_eidx", C.c_void_p), ("root_prior", C.c_void_p), ("exp_node", C.c_void_p),
                ("hidden", C.c_void_p), ("path", C.c_void_p), ("path_len", C.c_void_p), ("leaf", C.c_void_p),
                ("n_expanded", C.c_void_p), ("depth_sum", C.c_void_p), ("reward", C.c_void_p), ("minmax", C.c_void_p)]


class CuMindNativeError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if os.environ.get("CUMIND_B200_AUTOBUILD") == "1":
            _build.build()
        else:
            raise CuMindNativeError(
                f"{path} is missing: build it with `python -m cumind_b200._build` (needs nvcc) — "
                "cumind_b200 has no CPU fallback")
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise CuMindNativeError(f"{path} does not export {name}")
    lib.cumind_last_error.restype = C.c_char_p
    lib.cumind_param_count.restype = C.c_size_t
    lib.cumind_param_count.argtypes = [C.POINTER(NetDesc)]
    lib.cumind_tree_bytes.restype = C.c_size_t
    lib.cumind_tree_bytes.argtypes = [C.c_int32] * 4
    lib.cumind_select_actions_work_bytes.restype = C.c_size_t
    lib.cumind_select_actions_work_bytes.argtypes = [C.POINTER(NetDesc), C.c_int32, C.c_int32]
    lib.cumind_launch_count.restype = C.c_uint64
    vp, i32, f64, u64, sz = C.c_void_p, C.c_int32, C.c_double, C.c_uint64, C.c_size_t
    lib.cumind_net_create.argtypes = [C.POINTER(NetDesc), vp, sz, C.POINTER(vp)]
    lib.cumind_net_update_params.argtypes = [vp, vp, sz, vp]
    lib.cumind_net_destroy.argtypes = [vp]
    lib.cumind_initial_inference.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp]
    lib.cumind_recurrent_inference.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, i32, vp]
    lib.cumind_prediction.argtypes = [vp, vp, i32, vp, vp, i32, vp]
    lib.cumind_softmax.argtypes = [vp, i32, i32, vp, vp]
    lib.cumind_tree_create.argtypes = [i32, i32, i32, i32, vp, sz, C.POINTER(vp)]
    lib.cumind_tree_destroy.argtypes = [vp]
    lib.cumind_tree_get_view.argtypes = [vp, C.POINTER(TreeView)]
    search_args = [vp, vp, vp, C.POINTER(SearchCfg), vp, vp, vp, vp, vp]
    lib.cumind_search.argtypes = search_args
    lib.cumind_search_stepwise.argtypes = search_args
    lib.cumind_search_from_obs.argtypes = [vp, vp, vp, C.POINTER(SearchCfg), vp, vp, vp, vp, vp, vp]
    lib.cumind_net_reserve.argtypes = [vp, i32]
    lib.cumind_search_phase_clocks.argtypes = search_args[:-1] + [vp, vp]
    lib.cumind_search_plan.argtypes = [vp, vp, C.POINTER(SearchCfg), C.POINTER(i32 * 8)]
    lib.cumind_select_actions_host.argtypes = [vp, vp, vp, C.POINTER(SearchCfg), vp, vp, vp, vp, sz, vp]
    lib.cumind_select_actions_host_u8.argtypes = [vp, vp, vp, C.POINTER(SearchCfg), vp, vp, vp, vp, sz, vp]
    lib.cumind_tree_init_root.argtypes = [vp, vp, vp, C.POINTER(SearchCfg), vp, vp]
    lib.cumind_select.argtypes = [vp, C.POINTER(SearchCfg), vp]
    lib.cumind_expand.argtypes = [vp, vp, C.POINTER(SearchCfg), vp, vp]
    lib.cumind_backup.argtypes = [vp, vp, vp]
    lib.cumind_finalize.argtypes = [vp, C.POINTER(SearchCfg), vp, vp, vp, vp]
    lib.cumind_dirichlet.argtypes = [vp, i32, i32, f64, u64, u64, vp]
    lib.cumind_selftest_division.argtypes = [u64, u64, i32, vp, vp]
    f32 = C.c_float
    lib.cumind_adamw_step.argtypes = [vp, vp, vp, vp, vp, sz, f32, f32, f32, f32, f32, C.c_int64, f32, vp]
    lib.cumind_adamw_step_dev.argtypes = [vp, vp, vp, vp, vp, sz, f32, f32, f32, f32, f32, vp, f32, vp]
    lib.cumind_net_update_params_device.argtypes = [vp, vp, sz, vp]
    lib.cumind_p2p_create.argtypes = [sz, i32, i32, C.POINTER(vp), vp]
    lib.cumind_p2p_open.argtypes = [vp, vp]
    lib.cumind_p2p_grad_ptr.argtypes = [vp]
    lib.cumind_p2p_grad_ptr.restype = vp
    lib.cumind_p2p_allreduce_adamw.argtypes = [vp, vp, vp, vp, vp, sz, f32, f32, f32, f32, f32, vp, vp, vp]
    lib.cumind_p2p_error.argtypes = [vp]
    lib.cumind_p2p_reset.argtypes = [vp]
    lib.cumind_p2p_destroy.argtypes = [vp]
    lib.cumind_train_supported.argtypes = [C.POINTER(NetDesc), i32]
    lib.cumind_train_work_bytes.argtypes = [C.POINTER(NetDesc), i32]
    lib.cumind_train_work_bytes.restype = C.c_size_t
    lib.cumind_train_loss_grad.argtypes = [C.POINTER(NetDesc), vp, vp, vp, vp, vp, vp, i32, i32, vp, sz, vp, vp, vp]
    i64 = C.c_int64
    lib.cumind_sumtree_create.argtypes = [i64, C.POINTER(vp)]
    lib.cumind_sumtree_destroy.argtypes = [vp]
    lib.cumind_sumtree_destroy.restype = None
    lib.cumind_sumtree_tree_size.argtypes = [vp]
    lib.cumind_sumtree_tree_size.restype = i64
    lib.cumind_sumtree_clear.argtypes = [vp, vp]
    lib.cumind_sumtree_update.argtypes = [vp, vp, vp, i64, vp]
    lib.cumind_sumtree_sample.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.cumind_sumtree_read.argtypes = [vp, vp, vp]
    lib.cumind_replay_create.argtypes = [i64, i64, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.cumind_replay_destroy.argtypes = [vp]
    lib.cumind_replay_destroy.restype = None
    lib.cumind_replay_record_bytes.argtypes = [vp]
    lib.cumind_replay_record_bytes.restype = i64
    lib.cumind_replay_write.argtypes = [vp, i64, vp, i64, vp]
    lib.cumind_replay_gather.argtypes = [vp, vp, i64, i64, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.cumind_replay_values.argtypes = [vp, vp, vp, C.c_double, C.c_int32, vp, vp]
    _lib = lib
    return lib


_pending_destroy = []


def destroy_handle(fn, handle) -> None:
    """Call a `*_destroy` entry point (they cudaFree) — but never while the current stream is capturing a CUDA graph: a
    finaliser that runs during a capture (garbage collection) would invalidate it.  Deferred handles are destroyed by
    the next call made outside a capture."""
    import torch

    capturing = False
    try:
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:
        capturing = False
    if capturing:
        _pending_destroy.append((fn, handle))
        return
    while _pending_destroy:
        f, h = _pending_destroy.pop()
        f(h)
    fn(handle)


def check(rc: int) -> None:
    if rc != 0:
        msg = load().cumind_last_error().decode("utf-8", "replace")
        if msg.startswith("Unsupported observation shape") or "must be positive" in msg:
            raise ValueError(msg)
        raise CuMindNativeError(msg)


def make_desc(observation_shape, action_space_size: int, hidden_dim: int, num_blocks: int, conv_channels: int) -> NetDesc:
    d = NetDesc()
    shape = tuple(int(s) for s in observation_shape)
    d.obs_ndim = len(shape)
    for i in range(3):
        d.obs_shape[i] = shape[i] if i < len(shape) else 0
    d.hidden_dim, d.num_blocks, d.action_space_size, d.conv_channels = int(hidden_dim), int(num_blocks), int(action_space_size), int(conv_channels)
    return d


def launch_count() -> int:
    return int(load().cumind_launch_count())
