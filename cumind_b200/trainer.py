"""Train step on the GPU behind the reference's `Trainer` surface
(src/cumind/agent/trainer.py:23-219).

  Trainer.train_step()            trainer.py:83-107   sample -> prepare -> loss + grad -> AdamW -> update
  Trainer._prepare_batch()        trainer.py:118-153  first observation / policy, K actions / rewards (zero padded)
  n-step value targets            trainer.py:155-170  ONE batched target-network inference instead of one
                                                      batch-1 call per sampled item
  Trainer.compute_losses()        trainer.py:172-205  value MSE, policy CE, reward MSE over a K-step unroll,
                                                      same value / policy target at every step, /(K+1), /K
For vector-observation networks (hidden_dim 32 / 64: every BASELINE configuration) the loss and its gradient are ONE
hand-written kernel, `cumind_train_loss_grad` (csrc/train_fused.cu: the whole K-step unroll forward and backward on
chip, per-CTA partial gradients summed in a fixed order); image encoders and other widths go through torch autograd
over views of the same flat fp32 parameter tensor (blob layout of include/cumind_b200.h).  The optimiser is
`cumind_adamw_step` = optax.adamw fused with the 1/world scaling of the gradient, after one NCCL all-reduce of the
flat gradient (the only collective in the product).  The updated blob is pushed to the search kernels with
`cumind_net_update_params_device`.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import gc
import os
import weakref
from datetime import datetime
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import _native, params as P
from .checkpoint import load_checkpoint, save_checkpoint
from .memory import MemoryBuffer  # noqa: F401  (re-exported: the buffer runner.train uses, memory.py:88-113)
from .network import _stream_ptr


@contextlib.contextmanager
def _no_gc():
    """No cyclic garbage collection while a CUDA graph is being captured: a collected Agent / MCTS / DeviceSumTree frees
    device memory in its finaliser (cudaFree), and any such call during a capture invalidates it."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class _ParamViews:
    """Views of a flat parameter tensor, addressed by the reference's parameter paths."""

    def __init__(self, flat: torch.Tensor, dims):
        self.leaf: Dict[Tuple[Any, ...], torch.Tensor] = {}
        o = 0
        for path, shape in P._schema(*dims):
            n = int(np.prod(shape))
            self.leaf[path] = flat[o:o + n].view(*shape)
            o += n
        assert o == flat.numel()

    def __call__(self, *path):
        return self.leaf[tuple(path)]


def trainable_mask(dims) -> np.ndarray:
    """1 for nnx.Param leaves, 0 for BatchNorm running statistics (nnx.BatchStat; never trained, trainer.py:92)."""
    parts = []
    for path, shape in P._schema(*dims):
        parts.append(np.full(int(np.prod(shape)), 0 if path[-1] in ("mean", "var") else 1, dtype=np.uint8))
    return np.concatenate(parts)


def forward_losses(flat: torch.Tensor, dims, observations: torch.Tensor, actions: torch.Tensor, targets: Dict[str, torch.Tensor],
                   num_unroll_steps: int) -> Dict[str, torch.Tensor]:
    """trainer.py:172-205 as differentiable torch ops over `flat`."""
    obs_shape, A, H, L, Cc = dims
    w = _ParamViews(flat, dims)
    enc = ("representation_network", "encoder")

    def representation(obs):
        x = obs
        if len(obs_shape) == 1:
            for i in range(L):
                x = x @ w(*enc, "layers", i, "kernel") + w(*enc, "layers", i, "bias")
                if i < L - 1:
                    x = F.relu(x)
            return x

        def conv(x, name, stride):
            k = w(*name, "kernel").permute(3, 2, 0, 1)
            return F.conv2d(x, k, w(*name, "bias"), stride=stride, padding=1)

        def bn(x, name):
            g = w(*name, "scale") / torch.sqrt(w(*name, "var") + 1e-5)
            return (x - w(*name, "mean")[None, :, None, None]) * g[None, :, None, None] + w(*name, "bias")[None, :, None, None]

        x = F.relu(conv(obs.permute(0, 3, 1, 2), enc + ("initial_conv",), 2))
        for i in range(L):
            blk = enc + ("residual_blocks", i)
            t = F.relu(bn(conv(x, blk + ("conv1",), 1), blk + ("bn1",)))
            t = bn(conv(t, blk + ("conv2",), 1), blk + ("bn2",))
            x = F.relu(t + x)
        pooled = x.mean(dim=(2, 3))
        return pooled @ w(*enc, "final_dense", "kernel") + w(*enc, "final_dense", "bias")

    def prediction(h):
        logits = h @ w("prediction_network", "policy_head", "kernel") + w("prediction_network", "policy_head", "bias")
        value = h @ w("prediction_network", "value_head", "kernel") + w("prediction_network", "value_head", "bias")
        return logits, value[:, 0]

    def dynamics(h, a):
        x = h + w("dynamics_network", "action_embedding", "embedding")[a]
        for l in range(L):
            x = F.relu(x @ w("dynamics_network", "layers", l, "kernel") + w("dynamics_network", "layers", l, "bias")) + x
        reward = x @ w("dynamics_network", "reward_head", "kernel") + w("dynamics_network", "reward_head", "bias")
        return x, reward[:, 0]

    tv, tp, tr = targets["values"], targets["policies"], targets["rewards"]
    h = representation(observations)
    logits, value = prediction(h)
    value_loss = torch.mean((value - tv) ** 2)
    policy_loss = -torch.mean(torch.sum(tp * F.log_softmax(logits, dim=-1), dim=-1))
    reward_loss = torch.zeros((), dtype=flat.dtype, device=flat.device)
    K = int(num_unroll_steps)
    for step in range(K):
        h, reward = dynamics(h, actions[:, step].long())
        logits, value = prediction(h)
        reward_loss = reward_loss + torch.mean((reward - tr[:, step]) ** 2)
        value_loss = value_loss + torch.mean((value - tv) ** 2)          # same target at every step (trainer.py:193)
        policy_loss = policy_loss + -torch.mean(torch.sum(tp * F.log_softmax(logits, dim=-1), dim=-1))
    if K > 0:
        reward_loss = reward_loss / K
        value_loss = value_loss / (K + 1)
        policy_loss = policy_loss / (K + 1)
    return {"value_loss": value_loss, "policy_loss": policy_loss, "reward_loss": reward_loss}


class Trainer:
    def __init__(self, agent, memory, config):
        self.agent, self.memory, self.config = agent, memory, config
        timestamp = datetime.now().strftime("%Y%m%d-%H%M%S")
        self.checkpoint_dir = f"{config.checkpoint_root_dir}/{config.env_name}/{timestamp}"
        self.train_step_count = 0
        net = agent.network
        self._dims = net._dims
        self._lib = _native.load()
        dev = net.device
        self._flat = torch.as_tensor(net.blob.copy()).to(dev)      # parameters (blob layout); leaves alias it per step
        self.last_grad: Optional[torch.Tensor] = None
        self._m = torch.zeros_like(self._flat, requires_grad=False)
        self._v = torch.zeros_like(self._flat, requires_grad=False)
        self._count = 0
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._load_optimizer_state()
        agent._trainer = weakref.ref(self)
        self._agent_version = getattr(agent, "_state_version", 0)
        self._mask = torch.as_tensor(trainable_mask(self._dims)).to(dev)
        self.use_cuda_graph = True
        self.graph_distributed = True   # torch.distributed: two graphs around an eager all-reduce
        # the hand-written loss + gradient kernel (train_fused.cu) when the network shape is covered; False = torch autograd
        self.use_fused_kernel = bool(self._lib.cumind_train_supported(C.byref(net._desc), int(config.num_unroll_steps)))
        self._fused_buf: Dict[int, Any] = {}
        # torch.distributed on one node: the gradient all-reduce fused into AdamW over NVLink peer memory (csrc/p2p_adamw.cu)
        # instead of an NCCL call; False (or a failed set-up on any rank) = one NCCL all-reduce of the flat gradient
        self.use_p2p_allreduce = True
        self._p2p, self._p2p_grad, self._p2p_tried = None, None, False
        self._graph, self._graph_key, self._graph_out, self._static = None, None, None, None
        self._state_stale = False

    def _setup_p2p(self) -> bool:
        """CUDA IPC exchange of the per-rank gradient buffers (once).  Every rank must call this at the same point."""
        import socket

        import torch.distributed as dist

        self._p2p_tried = True
        world, rank = dist.get_world_size(), dist.get_rank()
        handle = C.c_void_p()
        raw = (C.c_ubyte * 128)()
        ok = world <= 16
        if ok:
            with torch.cuda.device(self._flat.device):
                ok = self._lib.cumind_p2p_create(self._flat.numel(), rank, world, C.byref(handle), C.cast(raw, C.c_void_p)) == 0
        info = [None] * world
        dist.all_gather_object(info, (socket.gethostname(), bool(ok), bytes(raw)))
        same_node = len({h for h, _, _ in info}) == 1
        if ok and same_node and all(o for _, o, _ in info):
            blob = b"".join(b for _, _, b in info)
            with torch.cuda.device(self._flat.device):
                ok = self._lib.cumind_p2p_open(handle, C.c_char_p(blob)) == 0
        else:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            if handle.value:
                self._lib.cumind_p2p_destroy(handle)
            return False
        self._p2p = handle
        n, ptr = self._flat.numel(), int(self._lib.cumind_p2p_grad_ptr(handle))

        class _View:   # the rank's peer-visible gradient buffer as a torch tensor (no copy)
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}

        self._p2p_grad = torch.as_tensor(_View(), device=self._flat.device)
        dist.barrier()
        return True

    def _p2p_active(self) -> bool:
        if not (self.use_p2p_allreduce and torch.distributed.is_available() and torch.distributed.is_initialized()):
            return False
        if torch.distributed.get_world_size() < 2 or torch.distributed.get_backend() != "nccl":
            return False
        if not self._p2p_tried:
            self._setup_p2p()
        return self._p2p is not None

    def _load_optimizer_state(self) -> None:
        """agent.optimizer_state ({"count", "mu", "nu"}, moments flat in blob layout or None) -> the device buffers, IN PLACE
        (captured CUDA graphs keep pointing at them)."""
        st = self.agent.optimizer_state if isinstance(self.agent.optimizer_state, dict) else {}
        n = self._flat.numel()
        with torch.no_grad():
            for name, buf in (("mu", self._m), ("nu", self._v)):
                val = st.get(name)
                if val is None:
                    buf.zero_()
                    continue
                arr = np.asarray(val, dtype=np.float32).reshape(-1)
                if arr.size != n:
                    raise ValueError(f"optimizer_state[{name!r}] has {arr.size} values, the parameter blob has {n}")
                buf.copy_(torch.as_tensor(arr).to(buf.device))
            self._count = int(st.get("count", 0) or 0)
            self._step_dev.fill_(self._count)

    def _reload_from_agent(self) -> None:
        """The agent's state was replaced (Agent.load_state / load_checkpoint): parameters, AdamW moments and step count."""
        with torch.no_grad():
            self._flat.copy_(torch.as_tensor(self.agent.network.blob).to(self._flat.device))
        self._load_optimizer_state()
        self._agent_version = getattr(self.agent, "_state_version", 0)
        self._state_stale = False
        if self._p2p is not None:      # the step counter (= barrier epoch) may have gone back: every rank reloads at the same point
            torch.distributed.barrier()
            _native.check(self._lib.cumind_p2p_reset(self._p2p))
            torch.distributed.barrier()

    # ---- batch preparation (trainer.py:118-170) --------------------------------------------------
    def _n_step_returns(self, batch: Sequence[List[Dict[str, Any]]]) -> np.ndarray:
        n, g = self.config.td_steps, self.config.discount
        out = np.zeros(len(batch), dtype=np.float64)
        boot_idx, boot_obs = [], []
        for b, item in enumerate(batch):
            for i in range(min(len(item), n)):
                out[b] += item[i]["reward"] * (g ** i)
            if len(item) > n:
                boot_idx.append(b)
                boot_obs.append(np.asarray(item[n]["observation"], dtype=np.float32))
        if boot_idx:  # one batched target-network call instead of one per item
            _, _, value = self.agent.target_network.initial_inference(np.stack(boot_obs))
            out[boot_idx] += (g ** n) * value[:, 0].double().cpu().numpy()
        return out

    def _prepare_batch(self, batch: Sequence[List[Dict[str, Any]]]):
        K = self.config.num_unroll_steps
        for item in batch:
            if not item:
                raise RuntimeError("Encountered empty item in batch.")
        values = self._n_step_returns(batch)
        obs = np.stack([np.asarray(item[0]["observation"], dtype=np.float32) for item in batch])
        policies = np.stack([np.asarray(item[0]["policy"], dtype=np.float32) for item in batch])
        actions = np.zeros((len(batch), K), dtype=np.int32)
        rewards = np.zeros((len(batch), K), dtype=np.float32)
        for b, item in enumerate(batch):
            for s, step in enumerate(item[:K]):
                actions[b, s], rewards[b, s] = step["action"], step["reward"]
        return obs, actions, {"values": values.astype(np.float32), "rewards": rewards, "policies": policies}

    def _device_batch(self):
        """_prepare_batch for a buffer with a DeviceTrajectoryStore: indices drawn as `memory.sample` draws them, rows
        gathered from the records in HBM, bootstrap values from ONE target-network call over the batch, n-step targets
        by cumind_replay_values.  The only host read is the two error counters (the reference's IndexError for a deque
        index past the end, memory.py:272-273, and its RuntimeError for an empty trajectory, trainer.py:123-125)."""
        store = self.memory.device_store
        g = self.memory.sample_batch_device(self.config.batch_size)
        oob, empty = g["flags"].tolist()
        if oob:
            raise IndexError("deque index out of range")
        if empty:
            raise RuntimeError("Encountered empty item in batch.")
        _, _, boot_value = self.agent.target_network.initial_inference(g["boot_observations"])
        values = store.n_step_values(g, boot_value)
        return g["observations"], g["actions"], {"values": values, "rewards": g["rewards"], "policies": g["policies"]}

    # ---- the step -----------------------------------------------------------------------------------
    def compute_losses(self, observations, actions, targets) -> Dict[str, torch.Tensor]:
        dev = self._flat.device
        t = {k: torch.as_tensor(np.asarray(v)).to(dev) for k, v in targets.items()}
        with torch.no_grad():
            return forward_losses(self._flat, self._dims, torch.as_tensor(np.asarray(observations, np.float32)).to(dev),
                                  torch.as_tensor(np.asarray(actions)).to(dev), t, self.config.num_unroll_steps)

    def _fwd_bwd_fused(self, observations, actions, targets, grad_out=None):
        """cumind_train_loss_grad: the K-step unroll forward + backward in one kernel, gradient in blob layout."""
        B, K = int(observations.shape[0]), int(self.config.num_unroll_steps)
        buf = self._fused_buf.get(B)
        if buf is None:
            dev = self._flat.device
            need = int(self._lib.cumind_train_work_bytes(C.byref(self.agent.network._desc), B))
            buf = (torch.empty(need, dtype=torch.uint8, device=dev), torch.empty_like(self._flat), torch.empty(4, dtype=torch.float32, device=dev))
            self._fused_buf = {B: buf}
        work, grad, losses = buf
        if grad_out is not None:
            grad = grad_out
        obs = observations.to(torch.float32).contiguous()
        act = actions.to(torch.int32).contiguous()
        tv, tp, tr = (targets[k].to(torch.float32).contiguous() for k in ("values", "policies", "rewards"))
        _native.check(self._lib.cumind_train_loss_grad(
            C.byref(self.agent.network._desc), C.c_void_p(self._flat.data_ptr()), C.c_void_p(obs.data_ptr()), C.c_void_p(act.data_ptr()),
            C.c_void_p(tv.data_ptr()), C.c_void_p(tp.data_ptr()), C.c_void_p(tr.data_ptr()), B, K, C.c_void_p(work.data_ptr()), work.numel(),
            C.c_void_p(grad.data_ptr()), C.c_void_p(losses.data_ptr()), _stream_ptr()))
        self.last_grad = grad
        return {"value_loss": losses[0], "policy_loss": losses[1], "reward_loss": losses[2], "total_loss": losses[3]}, grad

    def _fwd_bwd(self, observations, actions, targets, grad_out=None):
        """loss -> grad of the flat parameter blob (the fused kernel, or torch autograd over views of the blob); written to
        `grad_out` when given (the peer-visible buffer of the fused all-reduce)."""
        if self.use_fused_kernel and tuple(actions.shape) == (observations.shape[0], int(self.config.num_unroll_steps)):
            return self._fwd_bwd_fused(observations, actions, targets, grad_out)
        leaf = self._flat.detach().requires_grad_(True)   # fresh autograd leaf over the same storage
        losses = forward_losses(leaf, self._dims, observations, actions, targets, self.config.num_unroll_steps)
        total = losses["value_loss"] + losses["policy_loss"] + losses["reward_loss"]
        total.backward()
        grad = leaf.grad
        if grad_out is not None:
            with torch.no_grad():
                grad_out.copy_(grad)
            grad = grad_out
        self.last_grad = grad
        return {**{k: v.detach() for k, v in losses.items()}, "total_loss": total.detach()}, grad

    def _apply(self, grad, world: int) -> None:
        """AdamW on the (all-reduced) flat gradient, then push the parameters to the search kernels' layouts."""
        opt = self.agent.optimizer
        with torch.no_grad():
            _native.check(self._lib.cumind_adamw_step_dev(
                C.c_void_p(self._flat.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(self._m.data_ptr()),
                C.c_void_p(self._v.data_ptr()), C.c_void_p(self._mask.data_ptr()), self._flat.numel(), opt.learning_rate, opt.b1,
                opt.b2, opt.eps, opt.weight_decay, C.c_void_p(self._step_dev.data_ptr()), 1.0 / world, _stream_ptr()))
            self.agent.network.update_from_device(self._flat)

    def _apply_p2p(self) -> None:
        """cumind_p2p_allreduce_adamw: wait for every rank's gradient, sum over NVLink peer memory, AdamW, parameter push."""
        opt = self.agent.optimizer
        with torch.no_grad():
            _native.check(self._lib.cumind_p2p_allreduce_adamw(
                self._p2p, C.c_void_p(self._flat.data_ptr()), C.c_void_p(self._m.data_ptr()), C.c_void_p(self._v.data_ptr()),
                C.c_void_p(self._mask.data_ptr()), self._flat.numel(), opt.learning_rate, opt.b1, opt.b2, opt.eps, opt.weight_decay,
                C.c_void_p(self._step_dev.data_ptr()), None, _stream_ptr()))
            self.agent.network.update_from_device(self._flat)

    def _eager_step(self, observations, actions, targets, collective: bool = True) -> Dict[str, torch.Tensor]:
        """loss -> grad -> (all-reduce) -> AdamW -> push parameters; every launch on the current stream.  collective=False
        (graph warm-up) keeps the step local: a rank that re-captures alone must not issue collectives its peers do not."""
        if collective and self._p2p_active():
            out, _ = self._fwd_bwd(observations, actions, targets, grad_out=self._p2p_grad)
            self._apply_p2p()
            return out
        out, grad = self._fwd_bwd(observations, actions, targets)
        if not collective:
            self._apply(grad, 1)
            return out
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            torch.distributed.all_reduce(grad)            # one flat bucket over NCCL / NVLink
        self._apply(grad, world)
        return out

    def _graph_step(self, observations, actions, targets) -> Dict[str, torch.Tensor]:
        """Replay the whole step from one CUDA graph (static shapes; single process)."""
        key = (tuple(observations.shape), tuple(actions.shape))
        if self._graph is None or self._graph_key != key:
            self._static = (observations.clone(), actions.clone(), {k: v.clone() for k, v in targets.items()})
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            state = (self._flat.detach().clone(), self._m.clone(), self._v.clone(), self._step_dev.clone())
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
            with torch.cuda.stream(side):
                for _ in range(3):                       # warm-up outside the capture (allocator, cuBLAS handles); local when
                    self._eager_step(*self._static, collective=not distributed)   # distributed (see _eager_step)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():                        # undo the warm-up updates
                self._flat.copy_(state[0]); self._m.copy_(state[1]); self._v.copy_(state[2]); self._step_dev.copy_(state[3])
            self._graph = torch.cuda.CUDAGraph()
            with _no_gc(), torch.cuda.graph(self._graph):
                self._graph_out = self._eager_step(*self._static)
            with torch.no_grad():                        # capture does not execute: state is unchanged, but be explicit
                self._flat.copy_(state[0]); self._m.copy_(state[1]); self._v.copy_(state[2]); self._step_dev.copy_(state[3])
            self.agent.network.update_from_device(self._flat)
            self._graph_key = key
        so, sa, st = self._static
        so.copy_(observations); sa.copy_(actions)
        for k in st:
            st[k].copy_(targets[k])
        self._graph.replay()
        return self._graph_out

    def _graph_step_distributed(self, observations, actions, targets) -> Dict[str, torch.Tensor]:
        """torch.distributed: the step as two CUDA graphs (loss + gradient | AdamW + parameter push) around ONE eager
        NCCL all-reduce of the flat gradient — the collective itself is not captured."""
        world = torch.distributed.get_world_size()
        key = (tuple(observations.shape), tuple(actions.shape))
        if self._graph is None or self._graph_key != key:
            self._static = (observations.clone(), actions.clone(), {k: v.clone() for k, v in targets.items()})
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            state = (self._flat.detach().clone(), self._m.clone(), self._v.clone(), self._step_dev.clone())
            with torch.cuda.stream(side):
                for _ in range(3):                       # warm-up outside the capture, WITHOUT the collective: ranks may
                    _, g = self._fwd_bwd(*self._static)  # re-capture on different steps (batch shapes), and an extra
                    self._apply(g, world)                # all-reduce here would pair with another rank's real one
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():
                self._flat.copy_(state[0]); self._m.copy_(state[1]); self._v.copy_(state[2]); self._step_dev.copy_(state[3])
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with _no_gc(), torch.cuda.graph(g1):
                self._graph_out, self._graph_grad = self._fwd_bwd(*self._static)
            with _no_gc(), torch.cuda.graph(g2, pool=g1.pool()):
                self._apply(self._graph_grad, world)
            with torch.no_grad():
                self._flat.copy_(state[0]); self._m.copy_(state[1]); self._v.copy_(state[2]); self._step_dev.copy_(state[3])
            self.agent.network.update_from_device(self._flat)
            self._graph, self._graph_key = (g1, g2), key
        so, sa, st = self._static
        so.copy_(observations); sa.copy_(actions)
        for k in st:
            st[k].copy_(targets[k])
        g1, g2 = self._graph
        g1.replay()
        torch.distributed.all_reduce(self._graph_grad)
        g2.replay()
        self.last_grad = self._graph_grad
        return self._graph_out

    def train_step(self, batch: Optional[Tuple[Any, Any, Dict[str, Any]]] = None) -> Dict[str, float]:
        """One optimiser step.  `batch` = (observations, actions [B,K], targets) skips the replay sampling."""
        if getattr(self.agent, "_state_version", 0) != self._agent_version:   # Agent.load_state since the last step
            self._reload_from_agent()
        if batch is None:
            if not self.memory.is_ready(self.config.min_memory_size, self.config.min_memory_pct):
                return {}
            if getattr(self.memory, "device_store", None) is not None:
                batch = self._device_batch()      # sampled, gathered and targeted on the GPU
            else:
                batch = self._prepare_batch(self.memory.sample(self.config.batch_size))
        dev = self._flat.device
        to_dev = lambda x, dt=None: x.to(dev) if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dt)).to(dev)
        observations = to_dev(batch[0], np.float32)
        actions = to_dev(batch[1])
        targets = {k: to_dev(v) for k, v in batch[2].items()}
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        if self.use_cuda_graph and (not distributed or self._p2p_active()):
            out = self._graph_step(observations, actions, targets)      # ONE graph, the fused peer-memory all-reduce included
        elif self.use_cuda_graph and self.graph_distributed:
            out = self._graph_step_distributed(observations, actions, targets)
        else:
            out = self._eager_step(observations, actions, targets)
        self._count += 1
        self.train_step_count += 1
        self._state_stale = True
        # ONE device-to-host read per step (the reference returns Python floats, trainer.py:102-107)
        keys = [k for k in out if k != "total_loss"] + ["total_loss"]
        vals = torch.stack([out[k].detach().reshape(()).float() for k in keys]).tolist()
        res = {f"train/{k}": v for k, v in zip(keys[:-1], vals[:-1])}
        res["total_loss"] = vals[-1]
        return res

    def run_training_loop(self, env: Any) -> None:
        """trainer.py:44-81: num_episodes episodes of self-play, a train step every train_frequency episodes, the target
        network every target_update_frequency train steps, a checkpoint every checkpoint_interval episodes."""
        from .self_play import SelfPlay

        self_play = SelfPlay(self.config, self.agent, self.memory)
        for episode in range(1, int(self.config.num_episodes) + 1):
            self_play.run_episode(env)
            if episode % int(self.config.train_frequency) == 0:
                self.train_step()
                if self.train_step_count > 0 and self.train_step_count % int(self.config.target_update_frequency) == 0:
                    self.agent.update_target_network()
            if episode % int(self.config.checkpoint_interval) == 0:
                self.save_checkpoint(episode)

    def sync_optimizer_state(self) -> None:
        """Copy the AdamW moments back into agent.optimizer_state (for save_state / checkpoints)."""
        self.agent.optimizer_state = {"count": self._count, "mu": self._m.cpu().numpy(), "nu": self._v.cpu().numpy()}
        self._state_stale = False

    # ---- checkpoints (trainer.py:207-219) ------------------------------------------------------------
    def save_checkpoint(self, episode: int) -> str:
        self.sync_optimizer_state()
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        path = f"{self.checkpoint_dir}/episode_{episode:05d}.pkl"
        save_checkpoint(self.agent.save_state(), path)
        return path

    def load_checkpoint(self, path: str) -> None:
        self.agent.load_state(load_checkpoint(path))
        self._reload_from_agent()
