"""TEST INFRASTRUCTURE — canonical form of a search tree, for bit-exact comparison.

A tree is reduced to one record per node, keyed by the action path from the root and sorted by
(depth, path): visit_count (int32), value_sum (float64, compared as raw bits), prior (float64 bits),
expanded flag.  Works for the reference's Python `Node` graph (mcts.py:18-31) and for one row of
the structure-of-arrays slab described in include/cumind_b200.h.
"""

from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np


def _pack(records: List[Tuple[Tuple[int, ...], int, float, float, bool]]) -> Dict[str, np.ndarray]:
    records.sort(key=lambda r: (len(r[0]), r[0]))
    maxd = max((len(r[0]) for r in records), default=0)
    M = len(records)
    paths = np.full((M, max(maxd, 1)), -1, dtype=np.int16)
    plen = np.zeros(M, dtype=np.int16)
    for i, r in enumerate(records):
        plen[i] = len(r[0])
        paths[i, : len(r[0])] = r[0]
    return {
        "path_len": plen,
        "paths": paths,
        "visit": np.array([r[1] for r in records], dtype=np.int32),
        "value_sum": np.array([r[2] for r in records], dtype=np.float64),
        "prior": np.array([r[3] for r in records], dtype=np.float64),
        "expanded": np.array([r[4] for r in records], dtype=np.bool_),
    }


def canon_from_nodes(root: Any) -> Dict[str, np.ndarray]:
    """Canonical form of a Node graph (attributes prior, visit_count, value_sum, children)."""
    records = []
    stack = [((), root)]
    while stack:
        path, node = stack.pop()
        records.append((path, int(node.visit_count), float(node.value_sum), float(node.prior), len(node.children) > 0))
        for a, child in node.children.items():
            stack.append((path + (int(a),), child))
    return _pack(records)


def canon_from_soa(visit, value_sum, prior, root_prior, child_eidx, exp_node, A: int) -> Dict[str, np.ndarray]:
    """Canonical form of one tree of the SoA slab (1-D rows of each array)."""
    n_exp = int(np.sum(np.asarray(exp_node) >= 0))
    paths: Dict[int, Tuple[int, ...]] = {0: ()}
    records = [((), int(visit[0]), float(value_sum[0]), float(prior[0]), child_eidx[0] >= 0)]
    # expansions happen parent-before-child, so walking e in order always finds the parent's path
    for e in range(n_exp):
        parent = int(exp_node[e])
        ppath = paths[parent]
        for a in range(A):
            slot = 1 + e * A + a
            p = ppath + (a,)
            paths[slot] = p
            pr = float(root_prior[a]) if parent == 0 else float(np.float64(prior[slot]))
            records.append((p, int(visit[slot]), float(value_sum[slot]), pr, child_eidx[slot] >= 0))
    return _pack(records)


def canon_equal(a: Dict[str, np.ndarray], b: Dict[str, np.ndarray]) -> Tuple[bool, str]:
    """Bit-exact equality (float64 compared through their uint64 images)."""
    for k in ("path_len", "paths", "visit", "expanded"):
        if a[k].shape != b[k].shape or not np.array_equal(a[k], b[k]):
            return False, f"{k} differs"
    for k in ("value_sum", "prior"):
        if not np.array_equal(a[k].view(np.uint64), b[k].view(np.uint64)):
            idx = int(np.nonzero(a[k].view(np.uint64) != b[k].view(np.uint64))[0][0])
            return False, f"{k} differs at node {idx}: {a[k][idx]!r} vs {b[k][idx]!r}"
    return True, ""
