"""MCTS on the B200 behind the reference's `MCTS` / `Node` surface (src/cumind/core/mcts.py).

  MCTS(network, config).search(root_hidden[H], add_noise)   mcts.py:115-155 -> cumind_search, B = 1
  MCTS.search_batch(root_hidden[B,H], ...)                  the same search for B trees in one launch
  Node                                                      mcts.py:18-98: host-side view used by the
                                                            reference's unit tests and by tree dumps
`config` fields are read at call time (the reference's tests mutate it after construction,
tests/test_mcts.py:274).  The tree lives in a device slab owned by this object (torch allocation).
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _native
from .network import CuMindNetwork, _ptr, _stream_ptr

NODE_DTYPE = np.dtype([("value_sum", np.float64), ("visit", np.int32), ("prior", np.float32)])


class Node:
    """Host-side tree node with the reference's attributes and methods (mcts.py:18-98)."""

    def __init__(self, prior: float):
        self.prior = prior
        self.visit_count = 0
        self.value_sum = 0.0
        self.children: Dict[int, "Node"] = {}
        self.hidden_state = None

    def is_expanded(self) -> bool:
        return len(self.children) > 0

    def value(self) -> float:
        return 0.0 if self.visit_count == 0 else self.value_sum / self.visit_count

    def ucb_score(self, parent_visit_count: int, c_puct: float) -> float:
        if self.visit_count == 0:
            return float("inf")
        return self.value() + c_puct * self.prior * math.sqrt(parent_visit_count) / (1 + self.visit_count)

    def select_child(self, c_puct: float) -> int:
        total = self.visit_count
        return max(self.children.keys(), key=lambda a: self.children[a].ucb_score(total, c_puct))

    def expand(self, actions: List[int], priors, hidden_state) -> None:
        self.hidden_state = hidden_state
        pri = np.asarray(priors.detach().cpu() if isinstance(priors, torch.Tensor) else priors, dtype=np.float32)
        for action, prior in zip(actions, pri):
            self.children[action] = Node(float(prior))

    def backup(self, value: float) -> None:
        self.visit_count += 1
        self.value_sum += value


class _TreeSlab:
    """Device slab + C handle for B trees of at most S_max simulations."""

    def __init__(self, lib, device: torch.device, B: int, S_max: int, A: int, H: int):
        self.lib, self.device = lib, device
        self.B, self.S_max, self.A, self.H = B, S_max, A, H
        nbytes = int(lib.cumind_tree_bytes(B, S_max, A, H))
        if nbytes == 0:
            raise ValueError("invalid tree dimensions")
        self.slab = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        off = (-self.slab.data_ptr()) % 256
        self.base = self.slab.data_ptr() + off
        self.handle = C.c_void_p()
        _native.check(lib.cumind_tree_create(B, S_max, A, H, C.c_void_p(self.base), nbytes, C.byref(self.handle)))
        self.view = _native.TreeView()
        _native.check(lib.cumind_tree_get_view(self.handle, C.byref(self.view)))
        self.offset0 = off

    def _section(self, ptr: int, nbytes: int) -> torch.Tensor:
        start = ptr - self.slab.data_ptr()
        return self.slab[start:start + nbytes]

    def dump(self) -> Dict[str, np.ndarray]:
        """Copy the tree arrays to the host (NumPy)."""
        v, B, N, A, S1, H = self.view, self.B, self.view.N, self.A, self.S_max + 1, self.H
        def get(ptr, nbytes, dtype, shape):
            return self._section(ptr, nbytes).cpu().numpy().view(dtype).reshape(shape).copy()
        nodes = get(v.nodes, B * N * 16, NODE_DTYPE, (B, N))
        return {
            "visit": np.ascontiguousarray(nodes["visit"]), "value_sum": np.ascontiguousarray(nodes["value_sum"]),
            "prior": np.ascontiguousarray(nodes["prior"]),
            "child_eidx": get(v.child_eidx, B * N * 4, np.int32, (B, N)),
            "root_prior": get(v.root_prior, B * A * 8, np.float64, (B, A)),
            "exp_node": get(v.exp_node, B * S1 * 4, np.int32, (B, S1)),
            "hidden": get(v.hidden, B * S1 * H * 4, np.float32, (B, S1, H)),
            "n_expanded": get(v.n_expanded, B * 4, np.int32, (B,)),
            "depth_sum": get(v.depth_sum, B * 8, np.int64, (B,)),
            "reward": get(v.reward, B * N * 4, np.float32, (B, N)),     # CUMIND_TREE_MINMAX_DISCOUNT only (else uninitialised)
            "minmax": get(v.minmax, B * 2 * 8, np.float64, (B, 2)),
        }

    def __del__(self):
        try:
            if self.handle.value:
                _native.destroy_handle(self.lib.cumind_tree_destroy, self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class MCTS:
    """Monte Carlo Tree Search (mcts.py:101-222) for one tree or a batch of independent trees."""

    def __init__(self, network: CuMindNetwork, config: Any):
        self.network = network
        self.config = config
        self._lib = _native.load()
        self._slab: Optional[_TreeSlab] = None
        self.lanes_per_tree = 0          # 0 = auto
        self.schedule = _native.SCHED_AUTO   # launch tuning (no effect on results): see cumind_search_cfg
        self.teams_per_cta = 0
        self.net_warps_per_team = 0
        self.trees_per_team = 0
        self.stepwise = False            # True: one launch per phase (cumind_search_stepwise)
        # "reference" (mcts.py semantics, the default and the parity mode) or "minmax_discount": the OPTIONAL textbook rules
        # (min-max-normalised Q, reward + config.discount in the backup, leaf counted) — see CUMIND_TREE_MINMAX_DISCOUNT
        self.tree_mode = getattr(config, "tree_mode", "reference")
        self.last_cfg: Optional[_native.SearchCfg] = None

    # ---- helpers -------------------------------------------------------------------------------
    def _tree(self, B: int, S: int, A: int) -> _TreeSlab:
        H = self.network.hidden_dim
        s = self._slab
        if s is None or s.B != B or s.S_max < S or s.A != A or s.H != H:
            self._slab = None
            s = _TreeSlab(self._lib, self.network.device, B, S, A, H)
            self._slab = s
        return s

    def _cfg(self, noise_mode: int, seed: int = 0, tree_index_base: int = 0) -> _native.SearchCfg:
        c = self.config
        # the configuration is read at call time (tests mutate it after construction, tests/test_mcts.py:274); the struct of
        # the previous call is reused when nothing it is built from has changed
        key = (c.num_simulations, c.action_space_size, c.c_puct, c.exploration_fraction, c.dirichlet_alpha, noise_mode, self.network.net_mode,
               self.tree_mode, getattr(c, "discount", 1.0), self.lanes_per_tree, self.schedule, self.teams_per_cta, self.net_warps_per_team,
               self.trees_per_team, seed, tree_index_base)
        last = getattr(self, "_cfg_cache", None)
        if last is not None and last[0] == key:
            return _native.SearchCfg.from_buffer_copy(last[1])   # a copy: callers may edit their struct
        cfg = _native.SearchCfg()
        cfg.num_simulations = int(c.num_simulations)
        cfg.action_space_size = int(c.action_space_size)
        cfg.c_puct = float(c.c_puct)
        cfg.exploration_fraction = float(c.exploration_fraction)
        cfg.dirichlet_alpha = float(c.dirichlet_alpha)
        cfg.noise_mode = noise_mode
        cfg.net_mode = self.network.net_mode
        if self.tree_mode not in ("reference", "minmax_discount"):
            raise ValueError(f"unknown tree_mode {self.tree_mode!r}")
        cfg.tree_mode = _native.TREE_MINMAX_DISCOUNT if self.tree_mode == "minmax_discount" else _native.TREE_REFERENCE
        cfg.discount = float(getattr(c, "discount", 1.0))
        cfg.lanes_per_tree = int(self.lanes_per_tree)
        cfg.schedule = int(self.schedule)
        cfg.teams_per_cta = int(self.teams_per_cta)
        cfg.net_warps_per_team = int(self.net_warps_per_team)
        cfg.trees_per_team = int(self.trees_per_team)
        cfg.seed = int(seed) & (2**64 - 1)
        cfg.tree_index_base = int(tree_index_base)
        self._cfg_cache = (key, _native.SearchCfg.from_buffer_copy(cfg))
        return cfg

    # ---- reference API -------------------------------------------------------------------------
    def search(self, root_hidden_state, add_noise: bool = True) -> np.ndarray:
        """Action probabilities (float64 [A]) from one root hidden state [H] (mcts.py:115-155).
        With add_noise the Dirichlet draw comes from the global NumPy RNG exactly as the reference
        does (mcts.py:218), and is mixed in on the device."""
        h = root_hidden_state if isinstance(root_hidden_state, torch.Tensor) else torch.as_tensor(np.asarray(root_hidden_state))
        h = h.to(device=self.network.device, dtype=torch.float32).reshape(1, -1)
        noise = None
        if add_noise:
            A = min(int(self.config.action_space_size), self.network.action_space_size)
            noise = np.random.dirichlet([self.config.dirichlet_alpha] * A)[None, :]
        probs, _ = self.search_batch(h, noise=noise)
        return probs[0]

    def search_batch(self, root_hidden, noise=None, philox_seed: Optional[int] = None, tree_index_base: int = 0,
                     return_device: bool = False):
        """B independent searches in one launch.  root_hidden [B,H]; `noise` [B,A] float64 is mixed into
        the root priors when given; `philox_seed` draws Dirichlet(alpha) noise on the device instead.
        Returns (probs [B,A_cfg] float64, visits [B,A_cfg] int32) as NumPy arrays (or device tensors)."""
        net = self.network
        h = root_hidden if isinstance(root_hidden, torch.Tensor) else torch.as_tensor(np.asarray(root_hidden))
        h = h.to(device=net.device, dtype=torch.float32).contiguous()
        if h.dim() != 2 or h.shape[1] != net.hidden_dim:
            raise ValueError(f"expected root_hidden of shape (B, {net.hidden_dim}), got {tuple(h.shape)}")
        B = h.shape[0]
        A_cfg = int(self.config.action_space_size)
        A = min(A_cfg, net.action_space_size)
        S = int(self.config.num_simulations)
        if S <= 0:
            raise ValueError(f"num_simulations must be positive, got {S}")
        if B == 0:   # an empty batch is an empty result, not an error (no launch)
            self.last_root_value = torch.empty((0,), dtype=torch.float64, device=net.device)
            visits = torch.empty((0, A_cfg), dtype=torch.int32, device=net.device)
            probs = torch.empty((0, A_cfg), dtype=torch.float64, device=net.device)
            return (probs, visits) if return_device else (probs.cpu().numpy(), visits.cpu().numpy())
        tree = self._tree(B, S, A)
        d_noise = None
        mode = 0
        if noise is not None:
            d_noise = torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float64)).to(net.device)
            if tuple(d_noise.shape) != (B, A):
                raise ValueError(f"expected noise of shape ({B}, {A}), got {tuple(d_noise.shape)}")
            mode = 1
        elif philox_seed is not None:
            mode = 2
        cfg = self._cfg(mode, philox_seed or 0, tree_index_base)
        self.last_cfg = cfg
        visits = torch.empty((B, A_cfg), dtype=torch.int32, device=net.device)
        probs = torch.empty((B, A_cfg), dtype=torch.float64, device=net.device)
        root_value = torch.empty((B,), dtype=torch.float64, device=net.device)
        fn = self._lib.cumind_search_stepwise if self.stepwise else self._lib.cumind_search
        with torch.cuda.device(net.device):
            _native.check(fn(tree.handle, net._handle, _ptr(h), C.byref(cfg), _ptr(d_noise), _ptr(visits), _ptr(probs),
                             _ptr(root_value), _stream_ptr()))
        self.last_root_value = root_value
        if return_device:
            return probs, visits
        return probs.cpu().numpy(), visits.cpu().numpy()

    def search_from_obs(self, observations: torch.Tensor, noise: Optional[torch.Tensor] = None, philox_seed: Optional[int] = None,
                        tree_index_base: int = 0):
        """encode -> search (agent.py:73-79) for B device-resident observations through ONE C call
        (cumind_search_from_obs): both kernels are issued back to back from C, the search as a programmatic
        dependent launch behind the encoder.  `observations` [B,*obs] float32 and `noise` [B,A] float64 are CUDA
        tensors; returns device tensors (probs [B,A_cfg] float64, visits [B,A_cfg] int32).  Output and work
        buffers are kept per batch size, so repeated calls allocate nothing."""
        net = self.network
        B = int(observations.shape[0])
        A_cfg = int(self.config.action_space_size)
        A = min(A_cfg, net.action_space_size)
        tree = self._tree(B, int(self.config.num_simulations), A)
        buf = getattr(self, "_obs_buffers", None)
        if buf is None or buf[0].shape[0] != B or buf[1].shape[1] != A_cfg:
            net.reserve(B)
            buf = (torch.empty((B, net.hidden_dim), dtype=torch.float32, device=net.device),
                   torch.empty((B, A_cfg), dtype=torch.int32, device=net.device),
                   torch.empty((B, A_cfg), dtype=torch.float64, device=net.device),
                   torch.empty((B,), dtype=torch.float64, device=net.device))
            self._obs_buffers = buf
        hidden, visits, probs, root_value = buf
        mode = 1 if noise is not None else (2 if philox_seed is not None else 0)
        cfg = self._cfg(mode, philox_seed or 0, tree_index_base)
        self.last_cfg = cfg
        _native.check(self._lib.cumind_search_from_obs(tree.handle, net._handle, _ptr(observations), C.byref(cfg), _ptr(noise),
                                                       _ptr(hidden), _ptr(visits), _ptr(probs), _ptr(root_value), _stream_ptr()))
        self.last_root_value = root_value
        return probs, visits

    def plan(self, B: int) -> Dict[str, int]:
        """Launch configuration cumind_search would use for B trees with the current config."""
        A = min(int(self.config.action_space_size), self.network.action_space_size)
        tree = self._tree(B, int(self.config.num_simulations), A)
        out = (C.c_int32 * 8)()
        cfg = self._cfg(0)
        _native.check(self._lib.cumind_search_plan(tree.handle, self.network._handle, C.byref(cfg), C.byref(out)))
        v = [int(x) for x in out]
        plan = {"fused": 1 if v[0] else 0, "schedule": {0: "stepwise", 1: "warp", 2: "team", 4: "tensor-core"}[v[0]], "lanes_per_tree": v[1],
                "grid": v[3], "block": v[4], "smem_bytes": v[5]}
        if v[0] == _native.SCHED_TEAM:
            plan.update(trees_per_team=v[2], teams_per_cta=v[6], net_warps_per_team=v[7] % 256, layer_warpgroup_resident_k=v[7] // 256)
        else:
            plan.update(outputs_per_lane=v[2])
        return plan

    # ---- inspection ----------------------------------------------------------------------------
    def dump_trees(self) -> Dict[str, np.ndarray]:
        if self._slab is None:
            raise RuntimeError("no search has been run yet")
        return self._slab.dump()

    def tree_as_nodes(self, index: int = 0) -> Node:
        """Rebuild the reference-style Node graph of tree `index` from the device slab."""
        d = self.dump_trees()
        A = self._slab.A
        visit, wsum, prior = d["visit"][index], d["value_sum"][index], d["prior"][index]
        rprior, exp_node, hidden = d["root_prior"][index], d["exp_node"][index], d["hidden"][index]
        n_exp = int(d["n_expanded"][index])
        nodes: Dict[int, Node] = {0: Node(float(prior[0]))}
        nodes[0].visit_count, nodes[0].value_sum = int(visit[0]), float(wsum[0])
        for e in range(n_exp):
            parent_slot = int(exp_node[e])
            parent = nodes[parent_slot]
            parent.hidden_state = hidden[e].copy()
            for a in range(A):
                slot = 1 + e * A + a
                ch = Node(float(rprior[a]) if parent_slot == 0 else float(prior[slot]))
                ch.visit_count, ch.value_sum = int(visit[slot]), float(wsum[slot])
                parent.children[a] = ch
                nodes[slot] = ch
        return nodes[0]
