"""Replay buffers: cumind_b200.memory against the reference's own data/memory.py.

CPU: the reference's buffer tests (tests/test_data.py:24-97) re-pointed at the drop-in; golden vectors produced by
executing the reference's TreeBuffer (tests/golden/make_memory_golden.py); where /root/reference exists, a
differential run of all three classes against the reference module itself under seeded RNGs.
GPU: the device sum tree (csrc/replay.cu through the C ABI) against the same golden vectors and against the host
tree, bit for bit (float64 nodes compared as raw bits, indices exactly).
"""
import os
import random

import numpy as np
import pytest

from cumind_b200.config import Config
from cumind_b200.memory import Memory, MemoryBuffer, PrioritizedMemoryBuffer, TreeBuffer
from oracle import memory_exec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "memory_cases.npz")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _create(cls, capacity, config, **kw):
    if cls is PrioritizedMemoryBuffer:
        return cls(capacity=capacity, alpha=config.per_alpha, epsilon=config.per_epsilon, beta=config.per_beta)
    if cls is TreeBuffer:
        return cls(capacity=capacity, alpha=config.per_alpha, epsilon=config.per_epsilon, **kw)
    return cls(capacity=capacity)


BUFFERS = [MemoryBuffer, PrioritizedMemoryBuffer, TreeBuffer]


# ---- the reference's own tests (tests/test_data.py:24-97) ---------------------------------------------
@pytest.mark.parametrize("cls", BUFFERS)
def test_buffer_initialization(cls):
    b = _create(cls, 100, Config())
    assert b.capacity == 100 and len(b) == 0 and isinstance(b, Memory)


@pytest.mark.parametrize("cls", BUFFERS)
def test_add_len_capacity_clear(cls):
    b = _create(cls, 5, Config())
    b.add({"obs": 1, "action": 0})
    assert len(b) == 1
    for i in range(10):
        b.add({"obs": i})
    assert len(b) == 5 and b.get_size() == 5 and b.get_pct() == 100.0
    b.clear()
    assert len(b) == 0


@pytest.mark.parametrize("cls", BUFFERS)
def test_is_ready(cls):
    b = _create(cls, 20, Config())
    assert not b.is_ready(min_size=10)
    for i in range(10):
        b.add({"obs": i})
    assert b.is_ready(min_size=10) and not b.is_ready(min_size=11)
    assert b.is_ready(min_fill_pct=0.5) and not b.is_ready(min_fill_pct=0.51)
    for bad in ({"min_size": -1}, {"min_fill_pct": 1.5}, {"min_fill_pct": -0.1}, {}):
        with pytest.raises(ValueError):
            b.is_ready(**bad)
    z = MemoryBuffer(0)
    with pytest.raises(ValueError):
        z.get_pct()
    with pytest.raises(ValueError):
        z.is_ready(min_size=1)


@pytest.mark.parametrize("cls", BUFFERS)
def test_sample_from_buffer(cls):
    b = _create(cls, 20, Config())
    for i in range(15):
        b.add({"id": i})
    s = b.sample(batch_size=5)
    assert isinstance(s, list) and len(s) == 5
    if cls is not TreeBuffer:
        assert len({x["id"] for x in s}) == 5
    assert b.sample(batch_size=99) == list(b.buffer)  # larger than the buffer: everything (memory.py:108-110)


# ---- golden vectors from the reference's TreeBuffer ----------------------------------------------------
def _cases():
    z = np.load(GOLDEN)
    for ci in range(int(z["n_cases"][0])):
        p = f"c{ci}_"
        yield ci, {k[len(p):]: z[k] for k in z.files if k.startswith(p)}


def _replay(case, device=None):
    cap, n_add, n_upd, batch, seed = (int(x) for x in case["meta"])
    alpha, eps = (float(x) for x in case["alpha_eps"])
    buf = TreeBuffer(capacity=cap, alpha=alpha, epsilon=eps, device=device)
    for i in range(n_add):
        buf.add(i)
    assert np.array_equal(_bits(buf.sum_tree), _bits(case["tree_after_add"])), "sum tree after the adds"
    buf.update_priorities(list(case["upd_idx"]), list(case["upd_pri"]))
    assert np.array_equal(_bits(buf.sum_tree), _bits(case["tree_after_upd"])), "sum tree after update_priorities"
    assert _bits([buf.max_priority])[0] == _bits(case["max_priority"])[0]
    if len(case["retrieved"]):
        data_idx = buf.sample_indices(batch, uniforms=case["uniforms"])
        assert np.array_equal(data_idx + buf.tree_size - 1, case["retrieved"]), "retrieved leaves"
    return buf


@pytest.mark.parametrize("ci,case", list(_cases()), ids=lambda v: f"c{v}" if isinstance(v, int) else "")
def test_host_tree_buffer_matches_reference_golden(ci, case):
    _replay(case)


def test_host_tree_sample_consumes_random_like_the_reference():
    """sample() with the `random` module seeded equals sample_indices() with the same draws injected."""
    buf = TreeBuffer(capacity=50, alpha=0.6, epsilon=1e-6)
    for i in range(50):
        buf.add({"id": i})
    buf.update_priorities([49 + 3, 49 + 7, 49 + 3], [2.0, 0.5, 4.0])
    random.seed(123)
    u = [random.random() for _ in range(8)]
    random.seed(123)
    got = buf.sample(8)
    assert [g["id"] for g in got] == [int(i) for i in buf.sample_indices(8, uniforms=u)]


# ---- differential run against the reference module itself (build container only) ------------------------
@pytest.mark.skipif(not memory_exec.available(), reason="/root/reference is not present (GPU box)")
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_all_buffers_match_the_reference_module(seed):
    ref = memory_exec.load()
    rng = np.random.default_rng(seed)
    cap = int(rng.integers(1, 40))
    cfg = Config()
    pairs = [(ref.MemoryBuffer(capacity=cap), MemoryBuffer(cap)),
             (ref.PrioritizedMemoryBuffer(capacity=cap, alpha=cfg.per_alpha, epsilon=cfg.per_epsilon, beta=cfg.per_beta),
              PrioritizedMemoryBuffer(cap, cfg.per_alpha, cfg.per_epsilon, cfg.per_beta)),
             (ref.TreeBuffer(capacity=cap, alpha=cfg.per_alpha, epsilon=cfg.per_epsilon), TreeBuffer(cap, cfg.per_alpha, cfg.per_epsilon))]
    for step in range(200):
        op = int(rng.integers(0, 10))
        arg = int(rng.integers(1, cap + 3))
        pri = [float(x) for x in np.abs(rng.standard_normal(arg))]
        for r, m in pairs:
            if op < 5:
                r.add(step), m.add(step)
            elif op < 8:
                outs = []
                for b in (r, m):
                    random.seed(seed * 1000 + step)
                    np.random.seed(seed * 1000 + step)
                    try:
                        outs.append(b.sample(arg))
                    except Exception as e:  # the reference's IndexError after the deque wrapped
                        outs.append(type(e))
                assert outs[0] == outs[1], (type(m).__name__, step)
            elif op < 9 and hasattr(r, "update_priorities"):
                if isinstance(m, TreeBuffer):
                    idx = [int(x) + m.tree_size - 1 for x in rng.integers(0, m.tree_size, size=arg)]
                else:
                    idx = [int(x) for x in rng.integers(0, cap + 2, size=arg)]
                r.update_priorities(idx, pri), m.update_priorities(idx, pri)
            elif op == 9 and step % 50 == 49:
                r.clear(), m.clear()
            assert len(r) == len(m) and list(r.buffer) == list(m.buffer)
            if isinstance(m, TreeBuffer):
                assert np.array_equal(_bits(r.sum_tree), _bits(m.sum_tree)) and r.max_priority == m.max_priority
                assert (r.data_pointer, r.n_entries) == (m.data_pointer, m.n_entries)
            if isinstance(m, PrioritizedMemoryBuffer):
                assert list(r.priorities) == list(m.priorities) and r.max_priority == m.max_priority


# ---- device sum tree -------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ci,case", list(_cases()), ids=lambda v: f"c{v}" if isinstance(v, int) else "")
def test_device_sum_tree_matches_reference_golden(ci, case):
    import torch

    _replay(case, device=torch.device("cuda", 0))


@pytest.mark.gpu
@pytest.mark.parametrize("cap", [1, 3, 64, 1000, 14000, 30000])
def test_device_sum_tree_matches_host_tree_on_random_batches(cap):
    """Batched, ordered updates with duplicate indices (shared-memory and global-memory kernels) and batched
    retrieval: the device array and the sampled indices equal the host tree's."""
    import torch

    rng = np.random.default_rng(cap)
    host = TreeBuffer(cap, 0.6, 1e-6)
    dev = TreeBuffer(cap, 0.6, 1e-6, device=torch.device("cuda", 0))
    for i in range(min(cap, 300) + 5):
        host.add(i), dev.add(i)
    for rnd in range(4):
        n = int(rng.integers(1, 700))
        idx = [int(x) + host.tree_size - 1 for x in rng.integers(0, min(cap, 300), size=n)]
        pri = [float(x) for x in np.abs(rng.standard_normal(n)) * 2.0]
        host.update_priorities(idx, pri), dev.update_priorities(idx, pri)
        assert np.array_equal(_bits(host.sum_tree), _bits(dev.sum_tree)), (cap, rnd)
        assert host.max_priority == dev.max_priority
        u = rng.random(257)
        assert np.array_equal(host.sample_indices(257, uniforms=u), dev.sample_indices(257, uniforms=u))
    with pytest.raises(IndexError):
        dev._dev.update([dev._dev.n_nodes], [1.0])
    host.clear(), dev.clear()
    assert not dev.sum_tree.any() and len(dev) == 0
