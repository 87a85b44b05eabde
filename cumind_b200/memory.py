"""Replay buffers: the reference's three buffer classes, same names, constructor
arguments, return values, error behaviour and quirks (src/cumind/data/memory.py), plus a device-resident
sum tree for TreeBuffer.

Host side (the reference's own data structures — a deque of trajectories, Python / NumPy global RNGs):
    Memory                    memory.py:13-85    capacity, deque(maxlen), get_size / get_pct / is_ready / clear
    MemoryBuffer              memory.py:88-113   uniform `random.sample`; a batch larger than the buffer returns everything
    PrioritizedMemoryBuffer   memory.py:116-187  p_i**alpha / sum, `np.random.choice(..., replace=False)`
    TreeBuffer                memory.py:190-299  sum tree in a flat array of 2*tree_size-1 float64

Quirks that are reproduced on purpose (the parity tests run the reference file itself next to this one):
  * `_propagate` (memory.py:212-217) recurses once more after reaching the root: `(0 - 1) // 2 == -1`, so every
    update also adds `change` to `sum_tree[-1]`, the LAST leaf.  With a power-of-two capacity that is a real
    slot whose priority drifts; otherwise it is a padding leaf that is never compared during retrieval;
  * `TreeBuffer.sample` indexes the deque with the leaf's data index (memory.py:272-273): once the deque has
    wrapped, slot i of the tree and element i of the deque are different trajectories; an index past the end of
    the deque raises IndexError;
  * `update_priorities` of TreeBuffer takes TREE indices and stores `(priority + epsilon) ** alpha`
    (memory.py:285-288); PrioritizedMemoryBuffer stores the raw priority and ignores epsilon / beta.

Device side: `TreeBuffer(..., device=torch.device)` keeps `sum_tree` on the GPU (cumind_sumtree_* in the C ABI,
csrc/replay.cu): a batch of updates is applied in one launch, in parallel over the tree's nodes, with the float64
additions every node receives in the reference's order; sampling retrieves the whole batch in one launch from
injected or host-drawn uniforms.  The host mirror and the device tree hold the same bits (tests/test_memory.py).

Trajectory payloads on the device: `buffer.attach_device_store(DeviceTrajectoryStore(...))` mirrors every add into a
ring of fixed-size records in HBM (cumind_replay_*: what Trainer._prepare_batch reads of a trajectory), `add_batch`
stores many trajectories with one copy and one sum-tree launch, and `sample_batch_device(batch_size)` draws the
indices exactly as `sample` does (same RNG stream) and gathers the training batch on the device — with a device
TreeBuffer the sampled indices never leave the GPU.  The host deque stays, so `sample()` keeps the reference's API.
"""

from __future__ import annotations

import ctypes as C
import random
from collections import deque
from typing import Any, Deque, Dict, List, Optional, Sequence

import numpy as np


class Memory:
    """memory.py:13-85."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.buffer: Deque[Any] = deque(maxlen=capacity)
        self.device_store: Optional["DeviceTrajectoryStore"] = None
        self._added = 0   # adds since the last clear: slot of the next record = _added % capacity

    def add(self, sample: Any) -> None:
        raise NotImplementedError

    # ---- device payloads ---------------------------------------------------------------------------------
    def attach_device_store(self, store: "DeviceTrajectoryStore") -> None:
        """Mirror every trajectory added from now on into `store` (the buffer must be empty)."""
        if len(self.buffer):
            raise RuntimeError("attach_device_store needs an empty buffer")
        if store.capacity != self.capacity:
            raise ValueError(f"store capacity {store.capacity} != buffer capacity {self.capacity}")
        self.device_store, self._added = store, 0

    def _store_write(self, samples: Sequence[Any]) -> None:
        if self.device_store is not None and len(samples):
            self.device_store.write(self._added % self.capacity, samples)
        self._added += len(samples)

    def add_batch(self, samples: Sequence[Any]) -> None:
        """`add` for every sample in order; subclasses do it with one device copy / one sum-tree launch."""
        for sample in samples:
            self.add(sample)

    def _batch_indices(self, batch_size: int):
        """The deque indices `sample(batch_size)` would read, drawn from the same RNG stream (host array or device tensor)."""
        raise NotImplementedError

    def sample_batch_device(self, batch_size: int) -> Dict[str, Any]:
        """The training batch of `sample(batch_size)` gathered on the device (DeviceTrajectoryStore.gather)."""
        if self.device_store is None:
            raise RuntimeError("no device store attached (attach_device_store)")
        n = len(self.buffer)
        head = self._added % self.capacity if self._added >= self.capacity else 0   # slot of deque[0]
        return self.device_store.gather(self._batch_indices(batch_size), head, n)

    def sample(self, batch_size: int) -> List[Any]:
        raise NotImplementedError

    def __len__(self) -> int:
        return len(self.buffer)

    def get_size(self) -> int:
        return len(self)

    def get_pct(self) -> float:
        if self.capacity == 0:
            raise ValueError("Buffer capacity is zero, cannot calculate percentage.")
        return float((len(self) / self.capacity) * 100)

    def is_ready(self, min_size: int = 0, min_fill_pct: float = 0) -> bool:
        if min_size < 0 or min_fill_pct < 0 or min_fill_pct > 1:
            raise ValueError("min_size must be non-negative and min_fill_pct must be between 0 and 1.")
        if self.capacity == 0:
            raise ValueError("Buffer capacity is zero, cannot check readiness.")
        if min_size == 0 and min_fill_pct == 0:
            raise ValueError("At least one of min_size or min_fill_pct must be greater than zero.")
        return (len(self) >= min_size) and (len(self) >= min_fill_pct * self.capacity)

    def clear(self) -> None:
        self.buffer.clear()
        self._added = 0


class MemoryBuffer(Memory):
    """memory.py:88-113: uniform sampling without replacement (`random.sample`)."""

    def add(self, sample: Any) -> None:
        self.buffer.append(sample)
        self._store_write([sample])

    def add_batch(self, samples: Sequence[Any]) -> None:
        self.buffer.extend(samples)
        self._store_write(list(samples))

    def _batch_indices(self, batch_size: int):
        n = len(self.buffer)
        if n < batch_size:
            return np.arange(n, dtype=np.int64)
        return np.asarray(random.sample(range(n), batch_size), dtype=np.int64)   # the picks of random.sample(list(buffer), k)

    def sample(self, batch_size: int) -> List[Any]:
        if len(self.buffer) < batch_size:
            return list(self.buffer)
        return random.sample(list(self.buffer), batch_size)


class PrioritizedMemoryBuffer(Memory):
    """memory.py:116-187."""

    def __init__(self, capacity: int, alpha: float, epsilon: float, beta: float):
        super().__init__(capacity)
        self.alpha = alpha
        self.beta = beta
        self.epsilon = epsilon
        self.priorities: Deque[float] = deque(maxlen=capacity)
        self.max_priority = 1.0

    def add(self, sample: Any) -> None:
        self.buffer.append(sample)
        self.priorities.append(self.max_priority)
        self._store_write([sample])

    def add_batch(self, samples: Sequence[Any]) -> None:
        self.buffer.extend(samples)
        self.priorities.extend([self.max_priority] * len(samples))
        self._store_write(list(samples))

    def _batch_indices(self, batch_size: int):
        if len(self.buffer) < batch_size:
            return np.arange(len(self.buffer), dtype=np.int64)
        priorities = np.array(self.priorities)
        probabilities = (priorities**self.alpha) / np.sum(priorities**self.alpha)
        return np.random.choice(len(self.buffer), size=batch_size, p=probabilities, replace=False).astype(np.int64)

    def sample(self, batch_size: int) -> List[Any]:
        if len(self.buffer) < batch_size:
            return list(self.buffer)
        return [self.buffer[i] for i in self._batch_indices(batch_size)]

    def update_priorities(self, indices: Sequence[int], priorities: Sequence[float]) -> None:
        for idx, priority in zip(indices, priorities):
            if idx < len(self.priorities):
                self.priorities[idx] = priority
                self.max_priority = max(self.max_priority, priority)

    def clear(self) -> None:
        super().clear()
        self.priorities.clear()
        self.max_priority = 1.0


class DeviceSumTree:
    """The flat sum tree of TreeBuffer (memory.py:204-231, 249-253) in device memory, behind the C ABI."""

    def __init__(self, capacity: int, device=None):
        import torch

        from . import _native

        if not torch.cuda.is_available():
            raise _native.CuMindNativeError("DeviceSumTree needs a CUDA device (there is no CPU fallback)")
        self._native = _native
        self._lib = _native.load()
        self._torch = torch
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.capacity = int(capacity)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_sumtree_create(C.c_int64(self.capacity), C.byref(handle)))
        self._handle = handle
        self.tree_size = int(self._lib.cumind_sumtree_tree_size(handle))
        self.n_nodes = 2 * self.tree_size - 1

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._native.destroy_handle(self._lib.cumind_sumtree_destroy, self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self) -> None:
        with self._torch.cuda.device(self.device):
            self._native.check(self._lib.cumind_sumtree_clear(self._handle, self._stream()))

    def update(self, tree_idx, priority, check: bool = True) -> None:
        """`_update(tree_idx[k], priority[k])` for k = 0..n-1 IN ORDER (memory.py:249-253).  `check=False` skips the range check
        of indices that are already on the device (it costs a device synchronisation)."""
        torch = self._torch
        if torch.is_tensor(tree_idx):
            idx = tree_idx.to(self.device, torch.int64)
            lo, hi = (int(idx.min()), int(idx.max())) if idx.numel() and check else (0, 0)   # device indices: checked on the device (a sync)
        else:
            host_idx = np.asarray(tree_idx, dtype=np.int64)
            lo, hi = (int(host_idx.min()), int(host_idx.max())) if host_idx.size else (0, 0)
            idx = torch.as_tensor(host_idx).to(self.device)
        pri = torch.as_tensor(np.asarray(priority, dtype=np.float64)).to(self.device) if not torch.is_tensor(priority) else priority.to(self.device, torch.float64)
        if idx.numel() != pri.numel():
            raise ValueError("tree_idx and priority must have the same length")
        if idx.numel() == 0:
            return
        if lo < 0 or hi >= self.n_nodes:   # the NumPy tree of the reference raises here (negative indices are not supported)
            raise IndexError(f"tree index out of range [0, {self.n_nodes}): {lo if lo < 0 else hi}")
        idx, pri = idx.contiguous(), pri.contiguous()
        with torch.cuda.device(self.device):
            self._native.check(self._lib.cumind_sumtree_update(self._handle, C.c_void_p(idx.data_ptr()), C.c_void_p(pri.data_ptr()),
                                                               C.c_int64(idx.numel()), self._stream()))

    def sample(self, uniforms):
        """Retrieval of memory.py:265-273 for a whole batch: returns (tree_idx, data_idx) int64 device tensors.
        uniforms[i] in [0, 1) stands for the `random.random()` inside `random.uniform(segment*i, segment*(i+1))`."""
        torch = self._torch
        u = torch.as_tensor(np.asarray(uniforms, dtype=np.float64)).to(self.device) if not torch.is_tensor(uniforms) else uniforms.to(self.device, torch.float64)
        u = u.contiguous()
        n = u.numel()
        tree_idx = torch.empty(n, dtype=torch.int64, device=self.device)
        data_idx = torch.empty(n, dtype=torch.int64, device=self.device)
        if n:
            with torch.cuda.device(self.device):
                self._native.check(self._lib.cumind_sumtree_sample(self._handle, C.c_void_p(u.data_ptr()), C.c_int64(n),
                                                                   C.c_void_p(tree_idx.data_ptr()), C.c_void_p(data_idx.data_ptr()), self._stream()))
        return tree_idx, data_idx

    def numpy(self) -> np.ndarray:
        torch = self._torch
        out = torch.empty(self.n_nodes, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._native.check(self._lib.cumind_sumtree_read(self._handle, C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()


class DeviceTrajectoryStore:
    """A ring of fixed-size trajectory records in HBM (cumind_replay_* in include/cumind_b200.h): of every stored
    trajectory exactly what Trainer._prepare_batch (trainer.py:118-170) and the n-step target (trainer.py:172-205)
    read — observation and policy of step 0, the first num_unroll_steps actions and rewards, the discounted sum of the
    first td_steps rewards (float64, computed here with the reference's expression) and the bootstrap observation."""

    def __init__(self, capacity: int, observation_shape: Sequence[int], n_actions: int, unroll_steps: int, td_steps: int,
                 discount: float, device=None):
        import torch

        from . import _native

        if not torch.cuda.is_available():
            raise _native.CuMindNativeError("DeviceTrajectoryStore needs a CUDA device (there is no CPU fallback)")
        self._native, self._lib, self._torch = _native, _native.load(), torch
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.capacity, self.observation_shape = int(capacity), tuple(int(x) for x in observation_shape)
        self.obs_size = int(np.prod(self.observation_shape))
        self.n_actions, self.unroll_steps, self.td_steps, self.discount = int(n_actions), int(unroll_steps), int(td_steps), discount
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self._lib.cumind_replay_create(C.c_int64(self.capacity), C.c_int64(self.obs_size), self.n_actions,
                                                         self.unroll_steps, C.byref(handle)))
        self._handle = handle
        self.record_bytes = int(self._lib.cumind_replay_record_bytes(handle))
        self._staging = None

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._native.destroy_handle(self._lib.cumind_replay_destroy, self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack(self, trajectory: Sequence[Dict[str, Any]], out: Optional[np.ndarray] = None) -> np.ndarray:
        """One record (uint8 [record_bytes]) from a trajectory: a list of {"observation", "action", "reward", "policy"}."""
        rec = np.zeros(self.record_bytes, dtype=np.uint8) if out is None else out
        rec[...] = 0
        n, g, K, O, A = self.td_steps, self.discount, self.unroll_steps, self.obs_size, self.n_actions
        acc = np.zeros(1, dtype=np.float64)
        for i in range(min(len(trajectory), n)):
            acc[0] += trajectory[i]["reward"] * (g ** i)      # trainer.py:172-205, the same expression
        rec[0:8].view(np.float64)[0] = acc[0]
        rec[8:16].view(np.int32)[:] = (len(trajectory), 1 if len(trajectory) > n else 0)
        if len(trajectory) == 0:
            return rec
        w = rec[16:16 + 4 * (2 * O + A + K)].view(np.float32)
        w[0:O] = np.asarray(trajectory[0]["observation"], dtype=np.float32).reshape(O)
        if len(trajectory) > n:
            w[O:2 * O] = np.asarray(trajectory[n]["observation"], dtype=np.float32).reshape(O)
        w[2 * O:2 * O + A] = np.asarray(trajectory[0]["policy"], dtype=np.float32).reshape(A)
        steps = trajectory[:K]
        w[2 * O + A:2 * O + A + len(steps)] = [s["reward"] for s in steps]
        rec[16 + 4 * (2 * O + A + K):16 + 4 * (2 * O + A + 2 * K)].view(np.int32)[:len(steps)] = [s["action"] for s in steps]
        return rec

    def write(self, slot0: int, trajectories: Sequence[Sequence[Dict[str, Any]]]) -> None:
        """Records of `trajectories` to slots (slot0 + i) % capacity: one packed pinned staging buffer, one copy (two if the ring wraps)."""
        torch = self._torch
        trajectories = list(trajectories)
        drop = max(0, len(trajectories) - self.capacity)   # more than a ring's worth: only the last `capacity` survive
        trajectories, slot0 = trajectories[drop:], (slot0 + drop) % self.capacity
        n = len(trajectories)
        if n == 0:
            return
        if self._staging is None or self._staging.shape[0] < n:
            self._staging = torch.empty((max(n, 16), self.record_bytes), dtype=torch.uint8).pin_memory()
        torch.cuda.current_stream(self.device).synchronize()   # the previous copy out of the staging buffer is over
        stage = self._staging.numpy()
        for i, t in enumerate(trajectories):
            self.pack(t, stage[i])
        with torch.cuda.device(self.device):
            self._native.check(self._lib.cumind_replay_write(self._handle, C.c_int64(slot0), C.c_void_p(self._staging.data_ptr()),
                                                             C.c_int64(n), self._stream()))

    def gather(self, index, head: int, n_valid: int) -> Dict[str, Any]:
        """Batch rows from the records at deque indices `index` (host sequence or device int64 tensor).  Returns device
        tensors: observations [B,*obs], boot_observations, policies [B,A], rewards [B,K], actions [B,K] int32, partial
        [B] float64, has_boot [B] int32, flags int32[2] (out-of-range indices, empty trajectories)."""
        torch = self._torch
        idx = index.to(self.device, torch.int64) if torch.is_tensor(index) else torch.as_tensor(np.asarray(index, dtype=np.int64)).to(self.device)
        idx = idx.contiguous()
        B, dev, K, A = int(idx.numel()), self.device, self.unroll_steps, self.n_actions
        out = {"observations": torch.empty((B,) + self.observation_shape, dtype=torch.float32, device=dev),
               "boot_observations": torch.empty((B,) + self.observation_shape, dtype=torch.float32, device=dev),
               "policies": torch.empty((B, A), dtype=torch.float32, device=dev), "rewards": torch.empty((B, K), dtype=torch.float32, device=dev),
               "actions": torch.empty((B, K), dtype=torch.int32, device=dev), "partial": torch.empty(B, dtype=torch.float64, device=dev),
               "has_boot": torch.empty(B, dtype=torch.int32, device=dev), "flags": torch.zeros(2, dtype=torch.int32, device=dev)}
        if B:
            p = lambda k: C.c_void_p(out[k].data_ptr())
            with torch.cuda.device(dev):
                self._native.check(self._lib.cumind_replay_gather(
                    self._handle, C.c_void_p(idx.data_ptr()), C.c_int64(head), C.c_int64(n_valid), B, p("observations"), p("boot_observations"),
                    p("policies"), p("rewards"), p("actions"), p("partial"), p("has_boot"), p("flags"), self._stream()))
        return out

    def n_step_values(self, batch: Dict[str, Any], boot_value) -> Any:
        """values [B] float32 = float32(partial + discount**td_steps * float64(boot_value)) where the trajectory has a bootstrap."""
        torch = self._torch
        B = int(batch["partial"].numel())
        v = boot_value.to(self.device, torch.float32).reshape(B).contiguous()
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        if B:
            with torch.cuda.device(self.device):
                self._native.check(self._lib.cumind_replay_values(C.c_void_p(batch["partial"].data_ptr()), C.c_void_p(batch["has_boot"].data_ptr()),
                                                                  C.c_void_p(v.data_ptr()), C.c_double(self.discount ** self.td_steps), B,
                                                                  C.c_void_p(out.data_ptr()), self._stream()))
        return out


class TreeBuffer(Memory):
    """memory.py:190-299.  `device=None`: the NumPy sum tree of the reference.  `device=<cuda device>`: the same
    tree in device memory (DeviceSumTree); `sum_tree` then reads it back."""

    def __init__(self, capacity: int, alpha: float, epsilon: float, device=None):
        super().__init__(capacity)
        self.alpha = alpha
        self.epsilon = epsilon
        self.max_priority = 1.0
        self.tree_size = 1
        while self.tree_size < capacity:
            self.tree_size *= 2
        self._dev: Optional[DeviceSumTree] = DeviceSumTree(capacity, device) if device is not None else None
        self._host_tree = np.zeros(2 * self.tree_size - 1) if self._dev is None else None
        self.data_pointer = 0
        self.n_entries = 0

    @property
    def sum_tree(self) -> np.ndarray:
        return self._host_tree if self._dev is None else self._dev.numpy()

    # ---- host tree: the reference's recursion, iteratively (memory.py:212-231) ------------------
    def _propagate(self, idx: int, change: float) -> None:
        t = self._host_tree
        while True:
            parent = (idx - 1) // 2
            t[parent] += change      # idx == 0: parent == -1, the last leaf (reference quirk)
            if idx <= 0:
                break
            idx = parent

    def _retrieve(self, idx: int, s: float) -> int:
        t = self._host_tree
        while True:
            left = 2 * idx + 1
            if left >= len(t):
                return idx
            if s <= t[left]:
                idx = left
            else:
                s = s - t[left]
                idx = left + 1

    def _update(self, tree_idx: int, priority: float) -> None:
        if self._dev is not None:
            self._dev.update([tree_idx], [priority])
            return
        change = priority - self._host_tree[tree_idx]
        self._host_tree[tree_idx] = priority
        self._propagate(tree_idx, change)

    def add(self, sample: Any) -> None:
        self.buffer.append(sample)
        tree_idx = self.data_pointer + self.tree_size - 1
        self._update(tree_idx, self.max_priority)
        self.data_pointer = (self.data_pointer + 1) % self.capacity
        if self.n_entries < self.capacity:
            self.n_entries += 1
        self._store_write([sample])

    def add_batch(self, samples: Sequence[Any]) -> None:
        samples = list(samples)
        idx = [(self.data_pointer + k) % self.capacity + self.tree_size - 1 for k in range(len(samples))]
        self.buffer.extend(samples)
        if self._dev is not None:
            self._dev.update(idx, [self.max_priority] * len(idx))   # one launch, the updates in order
        else:
            for t in idx:
                self._update(t, self.max_priority)
        self.data_pointer = (self.data_pointer + len(samples)) % self.capacity
        self.n_entries = min(self.capacity, self.n_entries + len(samples))
        self._store_write(samples)

    def _batch_indices(self, batch_size: int):
        if len(self.buffer) < batch_size:
            return np.arange(len(self.buffer), dtype=np.int64)
        uniforms = np.array([random.random() for _ in range(batch_size)], dtype=np.float64)
        if self._dev is not None:
            return self._dev.sample(uniforms)[1]   # device tensor: the indices never visit the host
        return self.sample_indices(batch_size, uniforms)

    def sample_indices(self, batch_size: int, uniforms: Optional[np.ndarray] = None) -> np.ndarray:
        """The data indices `sample` would read (memory.py:265-273).  `uniforms` injects the draws of
        `random.random()`; by default they come from the `random` module in the reference's order."""
        if uniforms is None:
            uniforms = np.array([random.random() for _ in range(batch_size)], dtype=np.float64)
        uniforms = np.asarray(uniforms, dtype=np.float64)
        if self._dev is not None:
            _, data_idx = self._dev.sample(uniforms)
            return data_idx.cpu().numpy()
        segment = self._host_tree[0] / batch_size
        out = np.empty(batch_size, dtype=np.int64)
        for i in range(batch_size):
            a, b = segment * i, segment * (i + 1)
            s = a + (b - a) * uniforms[i]          # random.uniform(a, b)
            out[i] = self._retrieve(0, s) - self.tree_size + 1
        return out

    def sample(self, batch_size: int) -> List[Any]:
        if len(self.buffer) < batch_size:
            return list(self.buffer)
        return [self.buffer[int(i)] for i in self.sample_indices(batch_size)]

    def update_priorities(self, indices: Sequence[int], priorities: Sequence[float]) -> None:
        idx, pri = [], []
        for tree_idx, priority in zip(indices, priorities):
            p = (priority + self.epsilon) ** self.alpha
            idx.append(int(tree_idx))
            pri.append(float(p))
            self.max_priority = max(self.max_priority, p)
        if not idx:
            return
        if self._dev is not None:
            self._dev.update(idx, pri)  # one launch, applied in order
        else:
            for t, p in zip(idx, pri):
                self._update(t, p)

    def clear(self) -> None:
        super().clear()
        if self._dev is not None:
            self._dev.clear()
        else:
            self._host_tree.fill(0)
        self.data_pointer = 0
        self.n_entries = 0
        self.max_priority = 1.0
