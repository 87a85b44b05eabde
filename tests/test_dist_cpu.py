"""CPU tests of the multi-GPU bookkeeping: shard arithmetic, and a world_size-2 gloo run in which each
rank searches its shard with the oracle (standing in for the GPU search, which needs CUDA) and the
gathered result equals the single-process result — i.e. the layout has no cross-shard dependence."""

import os
import socket
import sys

import numpy as np
import pytest

from cumind_b200 import dist as cdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            b = cdist.all_bounds(total, world)
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert cdist.shard_bounds(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        cdist.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import liboracle as lo
    from oracle import network_np as nn

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, H, S = 3, 16, 12
        blob = nn.flatten_params(nn.random_params((4,), A, H, 2, 32, 21), (4,))
        desc = lo.make_desc((4,), A, H, 2)
        rng = np.random.default_rng(0)
        h = rng.standard_normal((total, H)).astype(np.float32)          # every rank derives the same global inputs
        noise = rng.dirichlet([0.3] * A, size=total)
        a, b = cdist.shard_bounds(total, rank, world)
        local = lo.search(desc, blob, h[a:b], S, A, 1.25, 0.25, noise[a:b])
        probs = cdist.gather_rows(local["out_probs"], total)
        visits = cdist.gather_rows(local["out_visits"], total)
        if rank == 0:
            full = lo.search(desc, blob, h, S, A, 1.25, 0.25, noise)
            np.save(os.path.join(out_dir, "ok.npy"),
                    np.array([np.array_equal(probs, full["out_probs"]), np.array_equal(visits, full["out_visits"])]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_reassemble_to_the_single_process_result(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, 101, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all()
