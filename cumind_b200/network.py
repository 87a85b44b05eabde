"""CuMindNetwork on the B200: same constructor, attributes and call signatures as the reference
(src/cumind/core/network.py:269-315), computing through the C ABI (include/cumind_b200.h).

  CuMindNetwork.initial_inference(obs)        network.py:288-300  -> cumind_initial_inference
  CuMindNetwork.recurrent_inference(h, a)     network.py:302-315  -> cumind_recurrent_inference
  .representation_network(obs)                network.py:182-191  -> cumind_initial_inference (hidden only)
  .dynamics_network(h, a)                     network.py:217-236  -> cumind_recurrent_inference (next, reward)
  .prediction_network(h)                      network.py:254-266  -> cumind_prediction
Inputs may be NumPy arrays or torch tensors; outputs are torch CUDA tensors (value / reward with the
reference's trailing [B,1] axis).  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native, params as P
from .prng import Rngs


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr() -> C.c_void_p:
    """The current CUDA stream of the current device as a pointer (the raw accessor skips building a Stream object)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class _SubNetwork:
    def __init__(self, owner: "CuMindNetwork"):
        self._owner = owner


class RepresentationNetwork(_SubNetwork):
    """r_theta (network.py:150-191): observation -> hidden state."""

    @property
    def encoder(self):
        return self

    @property
    def layers(self):
        enc = self._owner.params["representation_network"]["encoder"]
        return [enc["layers"][i] for i in range(len(enc["layers"]))] if "layers" in enc else []

    def __call__(self, observation) -> torch.Tensor:
        return self._owner._initial(observation, want_heads=False)[0]


class DynamicsNetwork(_SubNetwork):
    """g_theta (network.py:194-236): (hidden, action) -> (next hidden, reward[B,1])."""

    def __call__(self, state, action) -> Tuple[torch.Tensor, torch.Tensor]:
        nxt, reward, _, _ = self._owner._recurrent(state, action, want_heads=False)
        return nxt, reward


class PredictionNetwork(_SubNetwork):
    """f_theta (network.py:239-266): hidden -> (policy logits [B,A], value [B,1])."""

    def __call__(self, hidden_state) -> Tuple[torch.Tensor, torch.Tensor]:
        o = self._owner
        h = o._as_device(hidden_state, (o.hidden_dim,))
        B = h.shape[0]
        logits = torch.empty((B, o.action_space_size), dtype=torch.float32, device=o.device)
        value = torch.empty((B,), dtype=torch.float32, device=o.device)
        _native.check(o._lib.cumind_prediction(o._handle, _ptr(h), B, _ptr(logits), _ptr(value), o.net_mode, _stream_ptr()))
        return logits, value.unsqueeze(1)


class CuMindNetwork:
    def __init__(self, observation_shape: Sequence[int], action_space_size: int, hidden_dim: int, num_blocks: int,
                 conv_channels: int, rngs: Any = None, *, params: Optional[Dict[str, Any]] = None,
                 device: Optional[torch.device] = None, net_mode: int = _native.NET_FP32_EXACT):
        self.observation_shape = tuple(int(s) for s in observation_shape)
        if len(self.observation_shape) not in (1, 3):
            raise ValueError(f"Unsupported observation shape: {self.observation_shape}")  # network.py:179-180
        self.action_space_size = int(action_space_size)
        self.hidden_dim = int(hidden_dim)
        self.num_blocks = int(num_blocks)
        self.conv_channels = int(conv_channels)
        self.net_mode = int(net_mode)
        self._lib = _native.load()
        if not torch.cuda.is_available():
            raise _native.CuMindNativeError("cumind_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._dims = (self.observation_shape, self.action_space_size, self.hidden_dim, self.num_blocks, self.conv_channels)
        if params is None:
            if isinstance(rngs, Rngs):
                gen = rngs.generator()
            elif isinstance(rngs, np.random.Generator):
                gen = rngs
            else:
                gen = np.random.default_rng(rngs)
            params = P.init(*self._dims, gen)
        self._desc = _native.make_desc(*self._dims)
        self._handle = C.c_void_p()
        self.params: Dict[str, Any] = {}
        self._create(params)
        self.representation_network = RepresentationNetwork(self)
        self.dynamics_network = DynamicsNetwork(self)
        self.prediction_network = PredictionNetwork(self)

    # ---- parameters ------------------------------------------------------------------------------
    def _create(self, params: Dict[str, Any]) -> None:
        blob = P.flatten(params, *self._dims)
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_net_create(C.byref(self._desc), blob.ctypes.data_as(C.c_void_p), blob.size,
                                                      C.byref(self._handle)))
        self.params = P.unflatten(blob, *self._dims)
        self._blob = blob
        self._reserved = 0

    def reserve(self, max_batch: int) -> None:
        """Size the image encoder's activation workspace for batches of up to `max_batch` observations
        (cumind_net_reserve: allocates and synchronises — the inference entry points never do).  Called on demand
        by the wrappers below when a larger batch shows up; call it up front before capturing a CUDA graph."""
        if len(self.observation_shape) == 3 and int(max_batch) > self._reserved:
            with torch.cuda.device(self.device):
                _native.check(self._lib.cumind_net_reserve(self._handle, int(max_batch)))
            self._reserved = int(max_batch)

    def _refresh_host(self) -> None:
        """After a device-side optimiser step the host copy is stale: pull it back on demand."""
        src = getattr(self, "_device_source", None)
        if src is not None and getattr(self, "_host_stale", False):
            self._blob = src.detach().cpu().numpy().copy()
            self.params = P.unflatten(self._blob, *self._dims)
            self._host_stale = False

    def state(self) -> Dict[str, Any]:
        """The parameter tree (NumPy arrays, reference naming) — what nnx.state(network) holds."""
        self._refresh_host()
        return P.unflatten(self._blob, *self._dims)

    def load_state(self, params: Dict[str, Any]) -> None:
        blob = P.flatten(params, *self._dims)
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_net_update_params(self._handle, blob.ctypes.data_as(C.c_void_p), blob.size, _stream_ptr()))
        self._blob = blob
        self.params = P.unflatten(blob, *self._dims)
        self._host_stale = False

    def clone(self) -> "CuMindNetwork":
        return CuMindNetwork(*self._dims, params=self.state(), device=self.device, net_mode=self.net_mode)

    @property
    def blob(self) -> np.ndarray:
        self._refresh_host()
        return self._blob

    def update_from_device(self, flat: "torch.Tensor") -> None:
        """nnx.update(network, params) from a flat device tensor (blob layout): asynchronous, stays on the GPU."""
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_net_update_params_device(self._handle, C.c_void_p(flat.data_ptr()), flat.numel(), _stream_ptr()))
        self._device_source, self._host_stale = flat, True

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                _native.destroy_handle(self._lib.cumind_net_destroy, self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ---- helpers ---------------------------------------------------------------------------------
    def _as_device(self, x, tail: Tuple[int, ...], dtype=torch.float32) -> torch.Tensor:
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        t = t.to(device=self.device, dtype=dtype)
        if tuple(t.shape[1:]) != tuple(tail) or t.dim() != len(tail) + 1:
            raise ValueError(f"expected shape (batch, {', '.join(map(str, tail))}), got {tuple(t.shape)}")
        return t.contiguous()

    def _initial(self, observation, want_heads: bool):
        obs = self._as_device(observation, self.observation_shape)
        B = obs.shape[0]
        self.reserve(B)
        hidden = torch.empty((B, self.hidden_dim), dtype=torch.float32, device=self.device)
        logits = torch.empty((B, self.action_space_size), dtype=torch.float32, device=self.device) if want_heads else None
        value = torch.empty((B,), dtype=torch.float32, device=self.device) if want_heads else None
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_initial_inference(self._handle, _ptr(obs), B, _ptr(hidden), _ptr(logits), _ptr(value),
                                                             self.net_mode, _stream_ptr()))
        return hidden, logits, value

    def _recurrent(self, hidden_state, action, want_heads: bool):
        h = self._as_device(hidden_state, (self.hidden_dim,))
        B = h.shape[0]
        a = action if isinstance(action, torch.Tensor) else torch.as_tensor(np.asarray(action))
        a = a.to(device=self.device, dtype=torch.int32).reshape(-1).contiguous()
        if a.shape[0] != B:
            raise ValueError(f"expected {B} actions, got {a.shape[0]}")
        nxt = torch.empty((B, self.hidden_dim), dtype=torch.float32, device=self.device)
        reward = torch.empty((B,), dtype=torch.float32, device=self.device)
        logits = torch.empty((B, self.action_space_size), dtype=torch.float32, device=self.device) if want_heads else None
        value = torch.empty((B,), dtype=torch.float32, device=self.device) if want_heads else None
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_recurrent_inference(self._handle, _ptr(h), _ptr(a), B, _ptr(nxt), _ptr(reward), _ptr(logits),
                                                               _ptr(value), self.net_mode, _stream_ptr()))
        return nxt, reward.unsqueeze(1), logits, (value.unsqueeze(1) if value is not None else None)

    # ---- reference API ---------------------------------------------------------------------------
    def initial_inference(self, observation) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(hidden_state [B,H], policy_logits [B,A], value [B,1])  — network.py:288-300."""
        hidden, logits, value = self._initial(observation, want_heads=True)
        return hidden, logits, value.unsqueeze(1)

    def recurrent_inference(self, hidden_state, action) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """(next_hidden [B,H], reward [B,1], policy_logits [B,A], value [B,1])  — network.py:302-315."""
        return self._recurrent(hidden_state, action, want_heads=True)

    def softmax(self, logits) -> torch.Tensor:
        """jax.nn.softmax(logits, axis=-1) with the pinned exp (mcts.py:128,191)."""
        lg = logits if isinstance(logits, torch.Tensor) else torch.as_tensor(np.asarray(logits))
        lg = lg.to(device=self.device, dtype=torch.float32).contiguous()
        flat = lg.reshape(-1, lg.shape[-1])
        out = torch.empty_like(flat)
        _native.check(self._lib.cumind_softmax(_ptr(flat), flat.shape[0], flat.shape[1], _ptr(out), _stream_ptr()))
        return out.reshape(lg.shape)
