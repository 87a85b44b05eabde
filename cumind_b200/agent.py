"""Agent with the reference's surface (src/cumind/agent/agent.py:20-121) driving the B200 search.

  Agent(config, existing_state=None)              agent.py:23-59
  select_action(observation, training=False)      agent.py:61-89   (one environment)
  select_actions(observations[B,...], training)   batched extension: B environments, one fused search
  update_target_network / save_state / load_state agent.py:91-121
The optimiser is carried as hyper-parameters + AdamW moment arrays so checkpoints round-trip; the
training step itself is outside this round's hot path.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _native, params as P
from .config import Config
from .mcts import MCTS
from .network import CuMindNetwork, _stream_ptr
from .prng import Rngs, key


class AdamW:
    """Hyper-parameters of optax.adamw(learning_rate, weight_decay) (agent.py:48)."""

    def __init__(self, learning_rate: float, weight_decay: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.learning_rate, self.weight_decay, self.b1, self.b2, self.eps = learning_rate, weight_decay, b1, b2, eps

    def init(self, blob: np.ndarray) -> Dict[str, Any]:
        return {"count": 0, "mu": np.zeros_like(blob), "nu": np.zeros_like(blob)}


_current_device = getattr(torch._C, "_cuda_getDevice", torch.cuda.current_device)


class Agent:
    def __init__(self, config: Config, existing_state: Optional[Dict[str, Any]] = None):
        self.config = config
        if not torch.cuda.is_available():
            raise _native.CuMindNativeError("cumind_b200 needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        key.seed(config.seed)
        rngs = Rngs(params=key())
        # config.model_dtype (validated but unused by the reference, config.py:164-167) selects the arithmetic:
        # "float32" = fp32-exact parity mode, "bfloat16" = tcgen05 tensor-core mode
        net_mode = _native.NET_BF16_TC if getattr(config, "model_dtype", "float32") == "bfloat16" else _native.NET_FP32_EXACT
        self.network = CuMindNetwork(observation_shape=config.observation_shape, action_space_size=config.action_space_size,
                                     hidden_dim=config.hidden_dim, num_blocks=config.num_blocks,
                                     conv_channels=config.conv_channels, rngs=rngs, device=self.device, net_mode=net_mode)
        self.target_network = self.network.clone()
        self.optimizer = AdamW(config.learning_rate, config.weight_decay)
        self._state_version = 0      # bumped by load_state: a Trainer bound to this agent reloads its device copies
        self._trainer = None         # weakref to the Trainer that owns the live optimizer state, if any
        if existing_state:
            self.load_state(existing_state)
        else:
            self.optimizer_state = self.optimizer.init(self.network.blob)
        self.mcts = MCTS(self.network, config)
        self._rng = np.random.default_rng(key())
        self._host_plans: Dict[Any, Dict[str, Any]] = {}   # select_actions: buffers + C arguments per batch shape
        self._device_index = int(self.device.index)

    # ---- action selection ------------------------------------------------------------------------
    def select_action(self, observation: np.ndarray, training: bool = False) -> Tuple[int, np.ndarray]:
        """encode -> search -> argmax (eval) or sample (training)  — agent.py:61-89."""
        obs = np.asarray(observation, dtype=np.float32)[None]
        hidden, _, _ = self.network.initial_inference(obs)
        action_probs = self.mcts.search(root_hidden_state=hidden[0], add_noise=training)
        if training:
            action_idx = int(self._rng.choice(len(action_probs), p=action_probs))
        else:
            action_idx = int(np.argmax(action_probs))
        return action_idx, action_probs

    def _host_plan(self, B: int, A_cfg: int, A: int) -> Dict[str, Any]:
        """Per batch shape: pinned host buffers (and NumPy views of them), the device work buffer, the C arguments."""
        net, lib = self.network, self.mcts._lib
        plan = self._host_plans.get((B, A_cfg, A))
        if plan is None:
            net.reserve(B)
            need = int(lib.cumind_select_actions_work_bytes(C.byref(net._desc), B, A_cfg))
            work = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            pin = {"obs": torch.empty((B,) + tuple(net.observation_shape), dtype=torch.float32).pin_memory(),
                   "obs_u8": torch.empty((B,) + tuple(net.observation_shape) if len(net.observation_shape) == 3 else (1,), dtype=torch.uint8).pin_memory(),
                   "actions": torch.empty((B,), dtype=torch.int32).pin_memory(),
                   "probs": torch.empty((B, A_cfg), dtype=torch.float64).pin_memory(),
                   "noise": torch.empty((B, A), dtype=torch.float64).pin_memory()}
            plan = {"need": need, "work": work, "wbase": C.c_void_p(work.data_ptr() + ((-work.data_ptr()) % 256)), "pin": pin,
                    "np": {k: v.numpy() for k, v in pin.items()}, "ptr": {k: C.c_void_p(v.data_ptr()) for k, v in pin.items()},
                    "null": C.c_void_p(0)}
            self._host_plans.clear()  # one shape at a time: the buffers are sized for it
            self._host_plans[(B, A_cfg, A)] = plan
        return plan

    def host_buffers(self, B: int) -> Dict[str, np.ndarray]:
        """Pinned host arrays `{"obs": [B,*obs] float32, "noise": [B,A] float64}` that select_actions reads without a
        staging copy when they are passed back to it: environments write their observations straight into them."""
        A_cfg = int(self.config.action_space_size)
        views = self._host_plan(int(B), A_cfg, min(A_cfg, self.network.action_space_size))["np"]
        return {"obs": views["obs"], "noise": views["noise"], "obs_u8": views["obs_u8"]}

    def select_actions(self, observations: np.ndarray, training: bool = False, noise: Optional[np.ndarray] = None,
                       philox_seed: Optional[int] = None, tree_index_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """Batched select_action through ONE C-ABI call with host buffers (cumind_select_actions_host):
        H2D of the observations, initial_inference, fused search, first-arg-max, D2H of actions+probs.
        training=True draws root noise on the device (Philox) unless `noise` [B,A] is given, and
        samples the actions on the host from the returned probabilities."""
        net, lib = self.network, self.mcts._lib
        # uint8 image frames go to the device as bytes (a quarter of the PCIe traffic) and are scaled there: float32(u8) / 255
        obs = observations
        as_u8 = False
        if type(obs) is not np.ndarray or obs.dtype != np.float32 or not obs.flags.c_contiguous:   # (the common case skips all of this)
            as_u8 = isinstance(obs, np.ndarray) and obs.dtype == np.uint8 and len(net.observation_shape) == 3
            obs = np.ascontiguousarray(obs) if as_u8 else np.ascontiguousarray(obs, dtype=np.float32)
        B = obs.shape[0]
        if obs.shape[1:] != tuple(net.observation_shape):
            raise ValueError(f"expected observations of shape (B, {net.observation_shape}), got {obs.shape}")
        A_cfg = int(self.config.action_space_size)
        A = min(A_cfg, net.action_space_size)
        if B == 0:   # an empty batch is an empty result, not an error
            return np.zeros(0, dtype=np.int32), np.zeros((0, A_cfg), dtype=np.float64)
        tree = self.mcts._tree(B, int(self.config.num_simulations), A)
        mode = 0
        if noise is not None:
            mode = 1
        elif training or philox_seed is not None:
            mode = 2
            if philox_seed is None:
                philox_seed = int(self._rng.integers(0, 2**63 - 1))
        cfg = self.mcts._cfg(mode, philox_seed or 0, tree_index_base)
        plan = self._host_plan(B, A_cfg, A)
        views, ptr = plan["np"], plan["ptr"]
        # inputs written in place into host_buffers() are already in the pinned memory the GPU reads
        okey = "obs_u8" if as_u8 else "obs"
        dst = views[okey]
        if obs is not dst and obs.ctypes.data != dst.ctypes.data:
            dst[...] = obs
        if noise is not None:
            dst = views["noise"]
            if noise is not dst and not (isinstance(noise, np.ndarray) and noise.ctypes.data == dst.ctypes.data):
                dst[...] = noise
        def call():
            fn = lib.cumind_select_actions_host_u8 if as_u8 else lib.cumind_select_actions_host
            _native.check(fn(tree.handle, net._handle, ptr[okey], C.byref(cfg), ptr["noise"] if noise is not None else plan["null"],
                             ptr["actions"], ptr["probs"], plan["wbase"], plan["need"], _stream_ptr()))

        if _current_device() == self._device_index:
            call()
        else:
            with torch.cuda.device(self.device):
                call()
        probs = views["probs"].copy()
        if training:
            cdf = np.cumsum(probs, axis=1)
            u = self._rng.random(B)[:, None] * cdf[:, -1:]
            actions = np.minimum((u >= cdf).sum(axis=1), A_cfg - 1).astype(np.int32)
        else:
            actions = views["actions"].copy()
        return actions, probs

    # ---- state -----------------------------------------------------------------------------------
    def update_target_network(self) -> None:
        self.target_network.load_state(self.network.state())

    def save_state(self) -> Dict[str, Any]:
        trainer = self._trainer() if self._trainer is not None else None
        if trainer is not None and getattr(trainer, "_state_stale", False):
            trainer.sync_optimizer_state()   # the AdamW moments live on the device while training
        return {"network_state": self.network.state(), "optimizer_state": self.optimizer_state}

    def load_state(self, state: Dict[str, Any]) -> None:
        self.network.load_state(state["network_state"])
        self.optimizer_state = state["optimizer_state"]
        self.update_target_network()
        self._state_version += 1
