"""The OPTIONAL tree mode CUMIND_TREE_MINMAX_DISCOUNT (min-max-normalised Q, reward + discount in the backup, leaf
counted).  It is not the reference's search (mcts.py:51-65, 157-203 have none of these), so there is no reference
vector to pin it to: the mode is DEFINED by two independent restatements, oracle/cumind_oracle.c (cmo_search_mm,
arrays) and oracle/mcts_minmax.py (Python object graph), which must agree node for node (CPU test), and the CUDA
path must equal the C restatement bit for bit on the same case grid as the reference mode (GPU test).  The default
mode stays CUMIND_TREE_REFERENCE."""

import numpy as np
import pytest

from oracle import liboracle as lo
from oracle import mcts_minmax as mm
from oracle import network_np as nn

GRID = [(2, 64, 50), (4, 64, 50), (18, 32, 25), (3, 16, 30), (2, 128, 25), (5, 24, 20), (1, 32, 7)]


def _case(A, H, S, B, seed=0, bias_scale=0.3):
    params = nn.random_params((4,), A, H, 2, 32, 11 + A + seed, bias_scale=bias_scale)
    blob = nn.flatten_params(params, (4,))
    rng = np.random.default_rng(A * 1000 + H + seed)
    obs = rng.standard_normal((B, 4)).astype(np.float32)
    desc = lo.make_desc((4,), A, H, 2)
    h, _, _ = lo.initial_inference(desc, blob, obs)
    noise = rng.dirichlet([0.3] * A, size=B)
    return params, blob, desc, h, noise


@pytest.mark.parametrize("shape", GRID, ids=lambda s: f"A{s[0]}H{s[1]}S{s[2]}")
def test_c_and_python_restatements_agree(shape):
    A, H, S = shape
    B = 4
    params, blob, desc, h, noise = _case(A, H, S, B)
    for discount, c_puct, use_noise in ((0.997, 1.25, True), (0.9, 1.0, False), (1.0, 2.5, True)):
        nz = noise if use_noise else None
        o = lo.search_minmax(desc, blob, h, S, A, c_puct, discount, 0.25, nz)
        net = nn.NumpyCuMindNetwork(params, (4,))
        for b in range(B):
            root, stats = mm.run_search(net, h[b], S, A, c_puct, discount, 0.25, None if nz is None else nz[b])
            assert mm.flatten(root, A) == mm.flatten_arrays(o, b, A)
            assert stats.mn == o["minmax"][b, 0] and stats.mx == o["minmax"][b, 1]
            assert root.visit_count == S                      # the root is counted once per simulation
            assert o["out_visits"][b].sum() == S              # every simulation passes through exactly one root child
            assert abs(o["out_probs"][b].sum() - 1.0) < 1e-12


def test_mode_differs_from_the_reference_mode():
    """The two modes are different searches: same inputs, different trees (otherwise the mode flag would be vacuous)."""
    A, H, S = 4, 64, 50
    _, blob, desc, h, noise = _case(A, H, S, 8)
    ref = lo.search(desc, blob, h, S, A, 1.25, 0.25, noise)
    alt = lo.search_minmax(desc, blob, h, S, A, 1.25, 0.997, 0.25, noise)
    assert np.all(ref["visit"][:, 0] == 2 * S) and np.all(alt["visit"][:, 0] == S)
    assert not np.array_equal(ref["out_visits"], alt["out_visits"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", GRID, ids=lambda s: f"A{s[0]}H{s[1]}S{s[2]}")
def test_cuda_minmax_mode_matches_the_c_restatement(shape):
    from cumind_b200 import params as P
    from cumind_b200.mcts import MCTS
    from cumind_b200.network import CuMindNetwork

    A, H, S = shape
    B = 301
    _, blob, desc, h, noise = _case(A, H, S, B, bias_scale=0.05)

    class Cfg:
        num_simulations, action_space_size, c_puct, dirichlet_alpha, exploration_fraction, discount = S, A, 1.25, 0.3, 0.25, 0.997

    net = CuMindNetwork((4,), A, H, 2, 32, params=P.unflatten(blob, (4,), A, H, 2, 32))
    o = lo.search_minmax(desc, blob, h, S, A, 1.25, 0.997, 0.25, noise)
    mcts = MCTS(net, Cfg())
    mcts.tree_mode = "minmax_discount"
    probs, visits = mcts.search_batch(h, noise=noise)
    d = mcts.dump_trees()
    used = 1 + A * (S + 1)
    assert np.array_equal(d["visit"][:, :used], o["visit"][:, :used])
    assert np.array_equal(d["value_sum"][:, :used].view(np.uint64), o["value_sum"][:, :used].view(np.uint64))
    assert np.array_equal(d["reward"][:, :used].view(np.uint32), o["reward"][:, :used].view(np.uint32))
    assert np.array_equal(d["child_eidx"][:, :used], o["child_eidx"][:, :used])
    assert np.array_equal(d["minmax"].view(np.uint64), o["minmax"].view(np.uint64))
    assert np.array_equal(visits, o["out_visits"]) and np.array_equal(probs, o["out_probs"])
    assert np.array_equal(mcts.last_root_value.cpu().numpy().view(np.uint64), o["out_root_value"].view(np.uint64))
    # the default mode is untouched by the flag
    mcts.tree_mode = "reference"
    probs_r, visits_r = mcts.search_batch(h, noise=noise)
    ref = lo.search(desc, blob, h, S, A, 1.25, 0.25, noise)
    assert np.array_equal(visits_r, ref["out_visits"]) and np.array_equal(probs_r, ref["out_probs"])


@pytest.mark.gpu
def test_minmax_mode_rejects_the_tensor_core_network():
    from cumind_b200 import _native
    from cumind_b200.mcts import MCTS
    from cumind_b200.network import CuMindNetwork

    class Cfg:
        num_simulations, action_space_size, c_puct, dirichlet_alpha, exploration_fraction, discount = 5, 2, 1.25, 0.3, 0.25, 0.997

    net = CuMindNetwork((4,), 2, 64, 2, 32, net_mode=_native.NET_BF16_TC)
    mcts = MCTS(net, Cfg())
    mcts.tree_mode = "minmax_discount"
    with pytest.raises(_native.CuMindNativeError):
        mcts.search_batch(np.zeros((3, 64), np.float32))
