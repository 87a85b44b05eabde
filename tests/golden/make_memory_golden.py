"""Generates tests/golden/memory_cases.npz by EXECUTING the reference's TreeBuffer (oracle/memory_exec.py).

Run in the build container (where /root/reference exists):  python tests/golden/make_memory_golden.py
Each case: a capacity, `n_add` adds, a batch of update_priorities on random TREE indices (duplicates included, so
the order of the updates matters), then one sample() with seeded `random`.  Stored: the sum tree after the adds
and after the updates (raw float64), max_priority, the uniforms `random.random()` returned, and the tree indices
`_retrieve` returned.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import memory_exec  # noqa: E402

CASES = [  # capacity, n_add, n_update, batch, alpha, epsilon, seed
    (1, 1, 3, 1, 0.6, 1e-6, 1), (2, 2, 5, 2, 0.6, 1e-6, 2), (5, 3, 7, 2, 0.6, 1e-6, 3), (5, 12, 9, 4, 1.0, 0.0, 4),
    (8, 8, 20, 8, 0.6, 1e-6, 5), (8, 19, 33, 5, 0.5, 1e-3, 6), (100, 60, 64, 16, 0.6, 1e-6, 7), (100, 250, 300, 32, 0.6, 1e-6, 8),
    (1000, 1000, 1024, 256, 0.6, 1e-6, 9), (4096, 5000, 2048, 1024, 0.7, 1e-5, 10), (20000, 20000, 512, 128, 0.6, 1e-6, 11),
]


def main():
    ref = memory_exec.load()
    out = {}
    for ci, (cap, n_add, n_upd, batch, alpha, eps, seed) in enumerate(CASES):
        rng = np.random.default_rng(seed)
        buf = ref.TreeBuffer(capacity=cap, alpha=alpha, epsilon=eps)
        for i in range(n_add):
            buf.add(i)
        tree_after_add = buf.sum_tree.copy()
        n_leaf_used = min(n_add, cap)
        upd_idx = (rng.integers(0, n_leaf_used, size=n_upd) + buf.tree_size - 1).astype(np.int64)
        upd_pri = np.abs(rng.standard_normal(n_upd)) * 3.0
        upd_pri[rng.random(n_upd) < 0.1] = 0.0
        buf.update_priorities(list(upd_idx), list(upd_pri))
        tree_after_upd = buf.sum_tree.copy()
        # sample(): record what _retrieve returns; the uniforms are the seeded random.random() draws
        random.seed(seed)
        uniforms = np.array([random.random() for _ in range(batch)], dtype=np.float64)
        got = []
        orig = buf._retrieve

        def rec(idx, s, _orig=orig, _got=got):
            r = _orig(idx, s)
            if idx == 0:
                _got.append(r)
            return r

        buf._retrieve = rec
        random.seed(seed)
        try:
            buf.sample(batch)
        except IndexError:
            pass  # reference quirk: a retrieved leaf past the end of the deque; the retrieval itself is recorded
        buf._retrieve = orig
        p = f"c{ci}_"
        out[p + "meta"] = np.array([cap, n_add, n_upd, batch, seed], dtype=np.int64)
        out[p + "alpha_eps"] = np.array([alpha, eps], dtype=np.float64)
        out[p + "tree_after_add"] = tree_after_add
        out[p + "upd_idx"] = upd_idx
        out[p + "upd_pri"] = upd_pri
        out[p + "tree_after_upd"] = tree_after_upd
        out[p + "max_priority"] = np.array([buf.max_priority], dtype=np.float64)
        out[p + "uniforms"] = uniforms
        out[p + "retrieved"] = np.array(got, dtype=np.int64)
    out["n_cases"] = np.array([len(CASES)], dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "memory_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
