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


# ---- batched add / index draws (host) --------------------------------------------------------------------
@pytest.mark.parametrize("cls", BUFFERS)
def test_batch_indices_are_the_picks_of_sample(cls):
    """_batch_indices draws from the same RNG stream as sample() and names the same elements."""
    b = _create(cls, 40, Config())
    for i in range(55):
        b.add({"id": i})
    if hasattr(b, "update_priorities"):
        off = b.tree_size - 1 if cls is TreeBuffer else 0
        b.update_priorities([off + 3, off + 9, off + 3], [2.0, 0.25, 5.0])
    for k in (7, 40, 41):
        random.seed(5), np.random.seed(5)
        want = b.sample(k)
        random.seed(5), np.random.seed(5)
        idx = b._batch_indices(k)
        assert [b.buffer[int(i)] for i in idx] == want


@pytest.mark.parametrize("cls", BUFFERS)
def test_add_batch_equals_adds_in_order(cls):
    one, many = _create(cls, 13, Config()), _create(cls, 13, Config())
    items = [{"id": i} for i in range(31)]
    for lo, hi in ((0, 5), (5, 6), (6, 24), (24, 31)):
        for it in items[lo:hi]:
            one.add(it)
        many.add_batch(items[lo:hi])
        assert list(one.buffer) == list(many.buffer) and one._added == many._added
        if cls is TreeBuffer:
            assert np.array_equal(_bits(one.sum_tree), _bits(many.sum_tree))
            assert (one.data_pointer, one.n_entries) == (many.data_pointer, many.n_entries)
        if cls is PrioritizedMemoryBuffer:
            assert list(one.priorities) == list(many.priorities)


# ---- device sum tree: the parallel batched update -----------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cap", [5, 64, 1000, 5000, 30000])
def test_device_sum_tree_parallel_update_large_batches(cap):
    """Batches longer than one chunk (1024), heavy duplication, and chunks that fall back to the ordered path (an
    internal node or the last node among the indices): the tree equals the host tree bit for bit."""
    import torch

    rng = np.random.default_rng(cap + 1)
    host = TreeBuffer(cap, 0.6, 1e-6)
    dev = TreeBuffer(cap, 0.6, 1e-6, device=torch.device("cuda", 0))
    items = list(range(min(cap, 2500)))
    for it in items:
        host.add(it)
    dev.add_batch(items)
    assert np.array_equal(_bits(host.sum_tree), _bits(dev.sum_tree))
    for rnd, n in enumerate((1, 1024, 1025, 2600, 3333)):
        span = min(len(items), [len(items), 7, len(items), 40, len(items)][rnd])
        idx = [int(x) + host.tree_size - 1 for x in rng.integers(0, span, size=n)]
        if rnd == 3:
            idx[1700] = 0                        # an internal node in the second chunk
        if rnd == 4:
            idx[5] = 2 * host.tree_size - 2     # the last node in the first chunk
        pri = [float(x) for x in np.abs(rng.standard_normal(n)) * 3.0]
        host.update_priorities(idx, pri), dev.update_priorities(idx, pri)
        assert np.array_equal(_bits(host.sum_tree), _bits(dev.sum_tree)), (cap, rnd)
    u = rng.random(64)
    assert np.array_equal(host.sample_indices(64, uniforms=u), dev.sample_indices(64, uniforms=u))


# ---- trajectory records on the device ---------------------------------------------------------------------
def _trajectories(rng, n, obs_shape, A, max_len):
    out = []
    for _ in range(n):
        T = int(rng.integers(1, max_len + 1))
        out.append([{"observation": rng.standard_normal(obs_shape).astype(np.float32), "action": int(rng.integers(0, A)),
                     "reward": float(rng.standard_normal()), "policy": rng.dirichlet([1.0] * A).astype(np.float32)} for _ in range(T)])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["uniform", "prioritized", "tree_host", "tree_device"])
@pytest.mark.parametrize("obs_shape,cap,n_add", [((4,), 64, 40), ((4,), 32, 75), ((6, 5, 2), 16, 16)], ids=["filling", "wrapped", "image"])
def test_device_batch_equals_prepare_batch(kind, obs_shape, cap, n_add):
    """Trainer._device_batch (indices drawn like sample(), rows gathered from the records in HBM, n-step targets on the
    device) == Trainer._prepare_batch(memory.sample(...)) under the same RNG seeds, bit for bit — including the
    reference's quirk that a wrapped TreeBuffer indexes the deque with tree slots."""
    import torch

    from cumind_b200 import Agent
    from cumind_b200.memory import DeviceTrajectoryStore
    from cumind_b200.trainer import Trainer

    A, K, n = 3, 4, 5
    cfg = Config(hidden_dim=16, num_blocks=1, action_space_size=A, observation_shape=obs_shape, conv_channels=8, num_unroll_steps=K, td_steps=n,
                 batch_size=12, num_simulations=2, discount=0.97, min_memory_size=1)
    agent = Agent(cfg)
    dev = torch.device("cuda", 0)
    mk = {"uniform": lambda: MemoryBuffer(cap), "prioritized": lambda: PrioritizedMemoryBuffer(cap, 0.6, 1e-6, 0.4),
          "tree_host": lambda: TreeBuffer(cap, 0.6, 1e-6), "tree_device": lambda: TreeBuffer(cap, 0.6, 1e-6, device=dev)}[kind]
    host_buf, dev_buf = mk(), mk()
    dev_buf.attach_device_store(DeviceTrajectoryStore(cap, obs_shape, A, K, n, cfg.discount, device=dev))
    trajs = _trajectories(np.random.default_rng(3), n_add, obs_shape, A, 9)
    for t in trajs[:7]:
        host_buf.add(t), dev_buf.add(t)
    for t in trajs[7:]:
        host_buf.add(t)
    dev_buf.add_batch(trajs[7:])
    assert list(map(id, host_buf.buffer)) == list(map(id, dev_buf.buffer))
    if hasattr(host_buf, "update_priorities"):
        off = host_buf.tree_size - 1 if isinstance(host_buf, TreeBuffer) else 0
        idx, pri = [off + 1, off + 5, off + 2], [3.0, 0.5, 1.5]
        host_buf.update_priorities(idx, pri), dev_buf.update_priorities(idx, pri)
    t_host, t_dev = Trainer(agent, host_buf, cfg), Trainer(agent, dev_buf, cfg)
    for rnd in range(3):
        random.seed(rnd), np.random.seed(rnd)
        try:
            want = t_host._prepare_batch(host_buf.sample(cfg.batch_size))
        except IndexError:
            want = IndexError
        random.seed(rnd), np.random.seed(rnd)
        if want is IndexError:
            with pytest.raises(IndexError):
                t_dev._device_batch()
            continue
        obs, act, tg = t_dev._device_batch()
        assert np.array_equal(obs.cpu().numpy().view(np.uint32), want[0].view(np.uint32))
        assert np.array_equal(act.cpu().numpy(), want[1])
        for k in ("values", "rewards", "policies"):
            assert np.array_equal(tg[k].cpu().numpy().view(np.uint32), want[2][k].view(np.uint32)), k


@pytest.mark.gpu
def test_device_store_flags_and_train_step():
    """An empty trajectory raises the reference's RuntimeError when sampled; a train step from the device store equals
    the step from the host deque under the same seeds."""
    import torch

    from cumind_b200 import Agent
    from cumind_b200.memory import DeviceTrajectoryStore
    from cumind_b200.trainer import Trainer

    A, K, n, cap = 2, 3, 4, 50
    cfg = Config(hidden_dim=32, num_blocks=2, action_space_size=A, observation_shape=(4,), num_unroll_steps=K, td_steps=n, batch_size=16,
                 num_simulations=2, min_memory_size=1)
    dev = torch.device("cuda", 0)
    trajs = _trajectories(np.random.default_rng(8), 30, (4,), A, 12)
    losses, blobs = [], []
    for use_store in (False, True):
        agent = Agent(cfg)                       # seeded by cfg.seed: the same initial parameters both times
        buf = MemoryBuffer(cap)
        if use_store:
            buf.attach_device_store(DeviceTrajectoryStore(cap, (4,), A, K, n, cfg.discount, device=dev))
        buf.add_batch(trajs)
        tr = Trainer(agent, buf, cfg)
        random.seed(1)
        losses.append([tr.train_step()["total_loss"] for _ in range(3)])
        blobs.append(agent.network.blob.copy())
    assert losses[0] == losses[1] and np.isfinite(losses[0]).all()
    assert np.array_equal(blobs[0].view(np.uint32), blobs[1].view(np.uint32))
    bad = MemoryBuffer(4)
    bad.attach_device_store(DeviceTrajectoryStore(4, (4,), A, K, n, cfg.discount, device=dev))
    bad.add_batch([trajs[0], [], trajs[1]])
    with pytest.raises(RuntimeError, match="empty item"):
        Trainer(Agent(cfg), bad, cfg)._device_batch()
    with pytest.raises(RuntimeError):
        MemoryBuffer(4).sample_batch_device(2)
    full = MemoryBuffer(4)
    full.add([1])
    with pytest.raises(RuntimeError):
        full.attach_device_store(DeviceTrajectoryStore(4, (4,), A, K, n, cfg.discount, device=dev))


@pytest.mark.gpu
def test_self_play_feeds_the_device_store_and_the_trainer_reads_it():
    """The whole loop on the device side: SelfPlay.run_batch ends episodes into memory.add_batch (lazy Trajectory views packed
    into HBM records), Trainer.train_step samples / gathers / targets on the GPU; the device batch equals the host batch of
    the same draw."""
    import torch

    from cumind_b200 import Agent
    from cumind_b200.envs import CartPoleVec
    from cumind_b200.memory import DeviceTrajectoryStore
    from cumind_b200.self_play import SelfPlay
    from cumind_b200.trainer import Trainer

    cfg = Config(hidden_dim=32, num_simulations=6, batch_size=16, td_steps=3, num_unroll_steps=3, min_memory_size=4)
    agent = Agent(cfg)
    dev = torch.device("cuda", 0)
    memory = TreeBuffer(256, cfg.per_alpha, cfg.per_epsilon, device=dev)
    memory.attach_device_store(DeviceTrajectoryStore(256, cfg.observation_shape, cfg.action_space_size, cfg.num_unroll_steps, cfg.td_steps,
                                                     cfg.discount, device=dev))
    SelfPlay(agent, memory).run_batch(CartPoleVec(64, seed=3), num_steps=30, philox_seed=5)
    assert len(memory) > 16
    trainer = Trainer(agent, memory, cfg)
    random.seed(7)
    obs, act, tg = trainer._device_batch()
    random.seed(7)
    idx = memory._batch_indices(cfg.batch_size).cpu().numpy()
    want = trainer._prepare_batch([memory.buffer[int(i)] for i in idx])
    assert np.array_equal(obs.cpu().numpy().view(np.uint32), want[0].view(np.uint32)) and np.array_equal(act.cpu().numpy(), want[1])
    for k in ("values", "rewards", "policies"):
        assert np.array_equal(tg[k].cpu().numpy().view(np.uint32), want[2][k].view(np.uint32)), k
    info = trainer.train_step()
    assert np.isfinite(info["total_loss"])
