"""TreeBuffer priority update + stratified sampling: host tree (the reference's Python/NumPy structure) vs the
device sum tree (csrc/replay.cu).  PYTHONPATH=. python tools/bench_replay.py"""
import time

import numpy as np
import torch

from cumind_b200.memory import DeviceSumTree, TreeBuffer

for cap, batch in [(1000, 1024), (100000, 1024), (1000000, 4096)]:
    rng = np.random.default_rng(0)
    host = TreeBuffer(cap, 0.6, 1e-6)
    n_fill = min(cap, 20000)
    for i in range(n_fill):
        host.add(i)
    idx = rng.integers(0, n_fill, size=batch) + host.tree_size - 1
    pri = np.abs(rng.standard_normal(batch))
    u = rng.random(batch)
    t0 = time.perf_counter()
    host.update_priorities(list(idx), list(pri))
    t1 = time.perf_counter()
    host.sample_indices(batch, uniforms=u)
    t2 = time.perf_counter()
    dev = DeviceSumTree(cap)
    leaves = torch.arange(n_fill, dtype=torch.int64, device="cuda") + dev.tree_size - 1
    dev.update(leaves, torch.ones(n_fill, dtype=torch.float64, device="cuda"))
    d_idx = torch.as_tensor(idx, device="cuda")
    d_pri = torch.as_tensor((pri + 1e-6) ** 0.6, device="cuda")
    d_u = torch.as_tensor(u, device="cuda")
    for _ in range(3):
        dev.update(d_idx, d_pri)
        dev.sample(d_u)
    torch.cuda.synchronize()
    s, m, e = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record()
    for _ in range(20):
        dev.update(d_idx, d_pri, check=False)
    m.record()
    for _ in range(20):
        dev.sample(d_u)
    e.record()
    torch.cuda.synchronize()
    print(f"capacity {cap:8d} batch {batch}: host update {1e3 * (t1 - t0):8.2f} ms  sample {1e3 * (t2 - t1):8.2f} ms | "
          f"device update {s.elapsed_time(m) / 20:7.3f} ms  sample {m.elapsed_time(e) / 20:7.3f} ms")
