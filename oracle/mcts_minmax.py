"""TEST INFRASTRUCTURE — the OPTIONAL min-max-normalised / discounted tree mode as a Python object graph.

This mode (CUMIND_TREE_MINMAX_DISCOUNT in include/cumind_b200.h) is NOT what the reference computes:
src/cumind/core/mcts.py:51-65, 157-203 scores unvisited children +inf, has no min-max normalisation, no reward
and no discount in the backup, and never counts the leaf.  BASELINE.json's north_star names the textbook MuZero
rules ("min-max-normalised Q", "discounted backup") as a mode, so the mode is defined here and in
oracle/cumind_oracle.c (cmo_search_mm) — two independent restatements that must agree node for node — and the
CUDA path is tested against them.  "Parity unpinned": there is no reference implementation of this mode.

Rules (all statistics Python floats = IEEE float64, one rounded operation each, in this order):
  score(child)  = vs + ((c_puct * prior) * sqrt(N_parent)) / (1 + n)          N_parent = parent.visit_count
                  vs = 0.0 if n == 0 else normalise(reward + discount * (W / n))
  normalise(x)  = (x - mn) / (mx - mn) if mx > mn else x                       mn, mx: per-tree bounds (start +inf, -inf)
  select        = first maximum in action order (Python max keeps the first), from the root to an unexpanded node
  expand        = hidden, reward = dynamics(parent.hidden, action); priors, value = prediction(hidden); children = priors
  backup        = for node in reversed(path incl. the leaf): W += G; n += 1; q = reward + discount * (W / n);
                  mn = q if q < mn; mx = q if q > mx; G = reward + discount * G    (root reward 0, counted once per simulation)
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np

from .network_np import softmax


class MMNode:
    __slots__ = ("prior", "visit_count", "value_sum", "reward", "children", "hidden_state")

    def __init__(self, prior: float):
        self.prior = prior
        self.visit_count = 0
        self.value_sum = 0.0
        self.reward = 0.0
        self.children: Dict[int, "MMNode"] = {}
        self.hidden_state = None


class MinMax:
    def __init__(self):
        self.mn, self.mx = math.inf, -math.inf

    def update(self, q: float) -> None:
        if q < self.mn:
            self.mn = q
        if q > self.mx:
            self.mx = q

    def normalise(self, x: float) -> float:
        return (x - self.mn) / (self.mx - self.mn) if self.mx > self.mn else x


def _score(parent: MMNode, child: MMNode, c_puct: float, discount: float, mm: MinMax) -> float:
    explo = ((c_puct * child.prior) * math.sqrt(parent.visit_count)) / (1 + child.visit_count)
    vs = 0.0
    if child.visit_count != 0:
        vs = mm.normalise(child.reward + discount * (child.value_sum / child.visit_count))
    return vs + explo


def run_search(net, root_hidden: np.ndarray, num_simulations: int, n_actions: int, c_puct: float, discount: float,
               exploration_fraction: float = 0.25, noise: Optional[np.ndarray] = None):
    """One tree.  `net` is oracle.network_np.NumpyCuMindNetwork (batch-1 calls, like the reference makes them).
    Returns (root, MinMax)."""
    logits, _ = net.prediction_network(root_hidden[None])
    priors = softmax(logits)[0]
    A = min(n_actions, len(priors))
    root = MMNode(1.0)
    root.hidden_state = root_hidden
    for a in range(A):
        p = float(priors[a])
        if noise is not None:
            p = p * (1.0 - exploration_fraction) + float(noise[a]) * exploration_fraction
        root.children[a] = MMNode(p)
    mm = MinMax()
    for _ in range(num_simulations):
        node, path, action, parent = root, [root], -1, root
        while node.children:
            best_a, best_s = -1, 0.0
            for a, ch in node.children.items():
                s = _score(node, ch, c_puct, discount, mm)
                if best_a < 0 or s > best_s:
                    best_a, best_s = a, s
            parent, action = node, best_a
            node = node.children[best_a]
            path.append(node)
        hidden, reward = net.dynamics_network(parent.hidden_state[None], np.asarray([action], np.int32))
        logits, value = net.prediction_network(hidden)
        pri = softmax(logits)[0]
        node.hidden_state = hidden[0]
        node.reward = float(reward[0, 0])
        for a in range(A):
            node.children[a] = MMNode(float(pri[a]))
        G = float(value[0, 0])
        for nd in reversed(path):
            nd.value_sum = nd.value_sum + G
            nd.visit_count += 1
            mm.update(nd.reward + discount * (nd.value_sum / nd.visit_count))
            G = nd.reward + discount * G
    return root, mm


def flatten(root: MMNode, n_actions: int) -> List[tuple]:
    """Depth-first (action order) list of (visit_count, value_sum bits, reward, n_children) — the canonical form the
    tests compare with the array layout of the C oracle / CUDA slab."""
    out = []

    def walk(nd: MMNode):
        out.append((nd.visit_count, np.float64(nd.value_sum).view(np.uint64).item(), np.float32(nd.reward).view(np.uint32).item(),
                    len(nd.children)))
        for a in range(n_actions):
            if a in nd.children:
                walk(nd.children[a])

    walk(root)
    return out


def flatten_arrays(o: dict, b: int, n_actions: int) -> List[tuple]:
    """The same canonical list from the SoA arrays of tree b (visit / value_sum / reward / child_eidx)."""
    out = []
    visit, wsum, rew, eidx = o["visit"][b], o["value_sum"][b], o["reward"][b], o["child_eidx"][b]

    def walk(slot: int):
        e = int(eidx[slot])
        out.append((int(visit[slot]), wsum[slot:slot + 1].view(np.uint64)[0].item(), rew[slot:slot + 1].view(np.uint32)[0].item(),
                    n_actions if e >= 0 else 0))
        if e >= 0:
            for a in range(n_actions):
                walk(1 + e * n_actions + a)

    walk(0)
    return out
