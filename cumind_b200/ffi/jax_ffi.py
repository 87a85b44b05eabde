"""jax.ffi front end of the C ABI (north_star: "thin jax.ffi C-ABI custom calls").

Import-guarded: jax / jaxlib are not part of this image, so importing this module without them raises ImportError with the
reason, and nothing else in cumind_b200 imports it.  With JAX present:

    from cumind_b200.ffi import jax_ffi
    jax_ffi.register()                                   # builds libcumind_ffi.so against jax.ffi.include_dir() once, registers the targets
    next_h, reward, logits, value = jax_ffi.recurrent_inference(net_handle, hidden, action)
    visits, probs, root_value = jax_ffi.search(tree_handle, net_handle, root_hidden, noise, num_simulations=50, ...)

`net_handle` / `tree_handle` are the integer values of the cumind_net* / cumind_tree* created through ctypes
(cumind_b200/_native.py); arrays are jax.Arrays on the CUDA device, the calls are traced like any other JAX primitive.
Not exercised by the test-suite of this repository (no JAX here): the handlers are thin forwards to entry points that are.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

try:
    import jax
    import jax.numpy as jnp
    import numpy as np
except ImportError as exc:  # pragma: no cover - depends on the image
    raise ImportError("cumind_b200.ffi.jax_ffi needs jax and jaxlib (not installed in this image); use the ctypes binding "
                      "in cumind_b200/_native.py instead") from exc

from .. import _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcumind_ffi.so")
_TARGETS = {"cumind_initial_inference": "CumindInitialInference", "cumind_recurrent_inference": "CumindRecurrentInference",
            "cumind_search": "CumindSearch"}
_registered = False


def build() -> str:
    src = os.path.join(_HERE, "cumind_ffi.cc")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        lib_dir = os.path.dirname(_build.LIB_PATH)
        subprocess.run(["g++", "-std=c++17", "-shared", "-fPIC", "-O2", f"-I{jax.ffi.include_dir()}", f"-I{os.path.join(_HERE, '..', '..', 'include')}",
                        "-I/usr/local/cuda/include", src, f"-L{lib_dir}", "-lcumind_b200", f"-Wl,-rpath,{lib_dir}", "-o", _SO], check=True)
    return _SO


def register() -> None:
    global _registered
    if _registered:
        return
    lib = ctypes.CDLL(build())
    for target, symbol in _TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    _registered = True


def initial_inference(net_handle: int, obs, hidden_dim: int, action_space_size: int, net_mode: int = 0):
    register()
    B = obs.shape[0]
    out = (jax.ShapeDtypeStruct((B, hidden_dim), jnp.float32), jax.ShapeDtypeStruct((B, action_space_size), jnp.float32),
           jax.ShapeDtypeStruct((B,), jnp.float32))
    return jax.ffi.ffi_call("cumind_initial_inference", out)(obs, net=np.int64(net_handle), net_mode=np.int32(net_mode))


def recurrent_inference(net_handle: int, hidden, action, action_space_size: int, net_mode: int = 0):
    register()
    B, H = hidden.shape
    out = (jax.ShapeDtypeStruct((B, H), jnp.float32), jax.ShapeDtypeStruct((B,), jnp.float32),
           jax.ShapeDtypeStruct((B, action_space_size), jnp.float32), jax.ShapeDtypeStruct((B,), jnp.float32))
    return jax.ffi.ffi_call("cumind_recurrent_inference", out)(hidden, action.astype(jnp.int32), net=np.int64(net_handle), net_mode=np.int32(net_mode))


def search(tree_handle: int, net_handle: int, root_hidden, noise=None, *, num_simulations: int, action_space_size: int, c_puct: float = 1.0,
           exploration_fraction: float = 0.25, dirichlet_alpha: float = 0.3, net_mode: int = 0, tree_mode: int = 0, discount: float = 0.997,
           philox_seed=None, tree_index_base: int = 0):
    """MCTS.search for B trees: (visits [B,A] int32, probs [B,A] float64, root_value [B] float64); needs jax_enable_x64."""
    register()
    B = root_hidden.shape[0]
    mode = 1 if noise is not None else (2 if philox_seed is not None else 0)
    if noise is None:
        noise = jnp.zeros((B, action_space_size), jnp.float64)
    out = (jax.ShapeDtypeStruct((B, action_space_size), jnp.int32), jax.ShapeDtypeStruct((B, action_space_size), jnp.float64),
           jax.ShapeDtypeStruct((B,), jnp.float64))
    return jax.ffi.ffi_call("cumind_search", out)(
        root_hidden, noise, tree=np.int64(tree_handle), net=np.int64(net_handle), num_simulations=np.int32(num_simulations),
        action_space_size=np.int32(action_space_size), c_puct=np.float64(c_puct), exploration_fraction=np.float64(exploration_fraction),
        dirichlet_alpha=np.float64(dirichlet_alpha), noise_mode=np.int32(mode), net_mode=np.int32(net_mode), tree_mode=np.int32(tree_mode),
        discount=np.float64(discount), seed=np.int64(philox_seed or 0), tree_index_base=np.int64(tree_index_base))
