"""TEST INFRASTRUCTURE — Python port of the reference tree search (the CPU baseline that travels).

/root/reference does not exist on the GPU box, so the "reference CPU path" timed by bench.py is
this restatement of src/cumind/core/mcts.py: one Python object per node, dict of children,
float64 statistics, one batch-1 network call per dynamics / prediction evaluation — the same cost
structure as the reference (minus its log.debug f-strings and the eager-JAX dispatch, so it is
FASTER than the real thing; every speed-up quoted against it is conservative).

  TreeNode            <- Node                     mcts.py:18-98
  TreeNode.score      <- Node.ucb_score           mcts.py:51-65   (+inf when unvisited)
  TreeNode.best_action<- Node.select_child        mcts.py:67-76   (first maximum wins)
  run_search          <- MCTS.search              mcts.py:115-155
  _one_simulation     <- MCTS._simulate           mcts.py:157-203 (parents-only backup, root twice)
  noise mixing        <- _add_exploration_noise   mcts.py:205-222

tests/test_oracle_golden.py checks it tree-for-tree against vectors produced by the verbatim
reference (oracle/reference_exec.py).
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from .network_np import softmax

_INF = float("inf")


class TreeNode:
    __slots__ = ("prior", "visit_count", "value_sum", "children", "hidden_state")

    def __init__(self, prior: float):
        self.prior = prior
        self.visit_count = 0
        self.value_sum = 0.0
        self.children: Dict[int, "TreeNode"] = {}
        self.hidden_state = None

    def score(self, parent_visits: int, c_puct: float) -> float:
        n = self.visit_count
        if n == 0:
            return _INF
        return self.value_sum / n + c_puct * self.prior * math.sqrt(parent_visits) / (1 + n)

    def best_action(self, c_puct: float) -> int:
        best_a, best_s = -1, 0.0
        total = self.visit_count
        for a, child in self.children.items():
            s = child.score(total, c_puct)
            if best_a < 0 or s > best_s:
                best_a, best_s = a, s
        return best_a

    def grow(self, n_actions: int, priors: np.ndarray, hidden) -> None:
        self.hidden_state = hidden
        for a in range(min(n_actions, len(priors))):
            self.children[a] = TreeNode(float(priors[a]))


def _one_simulation(network, root: TreeNode, n_actions: int, c_puct: float) -> int:
    trail: List[TreeNode] = []
    node, action = root, -1
    while node.children:
        action = node.best_action(c_puct)
        trail.append(node)
        node = node.children[action]
    parent = trail[-1]
    nxt, _reward = network.dynamics_network(parent.hidden_state[None, :], np.array([action]))
    hidden = np.asarray(nxt)[0]
    logits, value = network.prediction_network(hidden[None, :])
    priors = np.asarray(softmax(logits)[0], dtype=np.float32)
    leaf_value = float(np.asarray(value)[0, 0])
    node.grow(n_actions, priors, hidden)
    for parent in trail:
        parent.visit_count += 1
        parent.value_sum += leaf_value
    root.visit_count += 1
    root.value_sum += leaf_value
    return len(trail)


def run_search(network, root_hidden: np.ndarray, num_simulations: int, action_space_size: int, c_puct: float,
               exploration_fraction: float = 0.25, noise: Optional[np.ndarray] = None) -> Tuple[np.ndarray, TreeNode, int]:
    """Returns (action_probs float64 [A], root node, sum of path lengths)."""
    root_hidden = np.asarray(root_hidden, dtype=np.float32)
    logits, _ = network.prediction_network(root_hidden[None, :])
    priors = softmax(logits)[0]
    root = TreeNode(1.0)
    root.grow(action_space_size, np.asarray(priors, dtype=np.float32), root_hidden)
    if noise is not None:
        f = exploration_fraction
        for i, child in enumerate(root.children.values()):
            child.prior = child.prior * (1 - f) + float(noise[i]) * f
    depth_sum = 0
    for _ in range(num_simulations):
        depth_sum += _one_simulation(network, root, action_space_size, c_puct)
    counts = np.zeros(action_space_size)
    for a, child in root.children.items():
        counts[a] = child.visit_count
    total = counts.sum()
    if total == 0:
        return np.ones(action_space_size) / action_space_size, root, depth_sum
    return counts / total, root, depth_sum
