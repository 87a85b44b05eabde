"""CPU tests: the oracle (C restatement, NumPy network, Python port) against the golden vectors
produced by the reference's own mcts.py executed verbatim (tests/golden/make_golden.py)."""

import math

import numpy as np
import pytest

from oracle import liboracle as lo
from oracle import mcts_port as mp
from oracle import network_np as nn
from oracle import reference_exec as rx
from oracle import tree_canon as tc
from tests import helpers as hp

SEARCH = hp.search_cases()
NETS = hp.network_cases()


def _c_search(case, d):
    desc = lo.make_desc(case["obs_shape"], case["A_net"], case["H"], case["L"])
    noise = None if d["noise"] is None else d["noise"][None]
    o = lo.search(desc, d["params"], d["root_h"][None], case["S"], case["A_cfg"], case["c_puct"], case["frac"], noise)
    A = o["A"]
    canon = tc.canon_from_soa(o["visit"][0], o["value_sum"][0], o["prior"][0], o["root_prior"][0], o["child_eidx"][0], o["exp_node"][0], A)
    return o, canon


@pytest.mark.parametrize("case", SEARCH, ids=[c["id"] for c in SEARCH])
def test_c_oracle_matches_reference_trees(case):
    d = hp.search_case_data(case)
    o, canon = _c_search(case, d)
    ok, msg = tc.canon_equal(d["canon"], canon)
    assert ok, msg
    assert np.array_equal(o["out_probs"][0], d["probs"])
    # invariants of the reference semantics
    assert o["visit"][0, 0] == 2 * case["S"]
    assert int(np.sum(o["exp_node"][0] >= 0)) == case["S"] + 1


def _unflatten(blob, case):
    """Rebuild the parameter tree from the flat blob (vector-observation nets)."""
    D, H, L, A = case["obs_shape"][0], case["H"], case["L"], case["A_net"]
    o = 0

    def take(*shape):
        nonlocal o
        n = int(np.prod(shape))
        a = blob[o : o + n].reshape(shape)
        o += n
        return a

    enc, cur = {}, D
    for i in range(L):
        enc[i] = {"kernel": take(cur, H), "bias": take(H)}
        cur = H
    dyn = {"action_embedding": {"embedding": take(A, H)}, "layers": {i: {"kernel": take(H, H), "bias": take(H)} for i in range(L)},
           "reward_head": {"kernel": take(H, 1), "bias": take(1)}}
    pred = {"policy_head": {"kernel": take(H, A), "bias": take(A)}, "value_head": {"kernel": take(H, 1), "bias": take(1)}}
    assert o == blob.size
    return {"representation_network": {"encoder": {"layers": enc}}, "dynamics_network": dyn, "prediction_network": pred}


@pytest.mark.parametrize("case", [c for c in SEARCH if c["S"] <= 50], ids=[c["id"] for c in SEARCH if c["S"] <= 50])
def test_python_port_matches_reference_trees(case):
    d = hp.search_case_data(case)
    net = nn.NumpyCuMindNetwork(_unflatten(d["params"], case), case["obs_shape"])
    probs, root, _ = mp.run_search(net, d["root_h"], case["S"], case["A_cfg"], case["c_puct"], case["frac"], d["noise"])
    ok, msg = tc.canon_equal(d["canon"], tc.canon_from_nodes(root))
    assert ok, msg
    assert np.array_equal(probs, d["probs"])


@pytest.mark.skipif(not rx.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("case", SEARCH[::7], ids=[c["id"] for c in SEARCH[::7]])
def test_golden_is_reproducible_from_the_verbatim_reference(case):
    d = hp.search_case_data(case)
    net = nn.NumpyCuMindNetwork(_unflatten(d["params"], case), case["obs_shape"])
    cfg = rx.SearchConfig(case["S"], case["A_cfg"], case["c_puct"], case["alpha"], case["frac"])
    probs, root = rx.search_with_tree(net, cfg, d["root_h"], d["noise"])
    ok, msg = tc.canon_equal(d["canon"], tc.canon_from_nodes(root))
    assert ok, msg
    assert np.array_equal(probs, d["probs"])


@pytest.mark.parametrize("case", NETS, ids=[c["id"] for c in NETS])
def test_c_network_matches_numpy_network(case):
    d = hp.network_case_data(case)
    desc = lo.make_desc(case["obs_shape"], case["A"], case["H"], case["L"], case["Cc"])
    assert lo.param_count(desc) == d["params"].size
    h, lg, v = lo.initial_inference(desc, d["params"], d["obs"])
    assert hp.bits_equal(h, d["h"]) and hp.bits_equal(lg, d["logits"]) and hp.bits_equal(v, d["value"])
    nx, rw, lg2, v2 = lo.recurrent_inference(desc, d["params"], d["h"], d["action"])
    assert hp.bits_equal(nx, d["next"]) and hp.bits_equal(rw, d["reward"])
    assert hp.bits_equal(lg2, d["logits2"]) and hp.bits_equal(v2, d["value2"])
    assert hp.bits_equal(lo.softmax(lg2), d["probs2"])


def test_pinned_exp_is_bit_reproducible_and_accurate():
    rng = np.random.default_rng(0)
    d32 = np.concatenate([-rng.exponential(3.0, 20000), -rng.uniform(0, 110, 20000), [0.0, -0.0, -1e-30, -87.3, -103.9, -150.0, -250.0]]).astype(np.float32)
    got_np = nn.pexp(d32.astype(np.float64))
    got_c = np.array([lo.pexp(float(x)) for x in d32])
    assert np.array_equal(got_np.view(np.uint64), got_c.view(np.uint64))
    exact = np.array([math.exp(float(x)) for x in d32])
    rel = np.abs(got_np - exact) / np.maximum(exact, 1e-300)
    assert np.max(rel[d32 >= -200]) < 4e-16
    # rounded once to fp32 it is within 1 ulp of the correctly rounded exp
    f = got_np.astype(np.float32)
    e = exact.astype(np.float32)
    assert np.all(np.abs(f.view(np.int32).astype(np.int64) - e.view(np.int32).astype(np.int64)) <= 1)


def test_numpy_fma_emulation_is_exact():
    """network_np.fma32 == fmaf bit for bit: random operands over wide exponent ranges, exact
    cancellation, specials, and a family of true double-rounding cases that a naive float64
    evaluation gets wrong."""
    rng = np.random.default_rng(0)

    def rnd(n, span):
        return (rng.standard_normal(n) * np.exp2(rng.integers(-span, span, n))).astype(np.float32)

    for span in (2, 10, 40, 120):
        a, b, c = rnd(300000, span), rnd(300000, span), rnd(300000, span)
        assert hp.bits_equal(nn.fma32(a, b, c), lo.fmaf(a, b, c))
    a, b = rnd(100000, 3), rnd(100000, 3)
    c = (-(a.astype(np.float64) * b.astype(np.float64))).astype(np.float32)
    assert hp.bits_equal(nn.fma32(a, b, c), lo.fmaf(a, b, c))
    sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, -3.4e38, 1.0, -1.0, 1e-38], np.float32)
    A, B, Cc = (m.ravel() for m in np.meshgrid(sp, sp, sp, indexing="ij"))
    assert hp.bits_equal(nn.fma32(A, B, Cc), lo.fmaf(A, B, Cc))
    u = np.arange(1, 200, dtype=np.float64)
    a = (1 + u * 2.0**-23).astype(np.float32)
    b = (2.0**-24 * (1 - u * 2.0**-23)).astype(np.float32)
    c = np.full(a.shape, 1 + 2.0**-23, np.float32)      # a*b + c sits just below a float32 midpoint
    want = lo.fmaf(a, b, c)
    assert hp.bits_equal(nn.fma32(a, b, c), want) and hp.bits_equal(nn.fma32(-a, b, -c), lo.fmaf(-a, b, -c))
    naive = (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)
    assert not np.array_equal(naive, want)  # double rounding is real; the emulation avoids it


def test_reference_quirks_are_reproduced():
    """root.n == 2S, leaf never backed up, sum of root-children visits == S - A once S >= 2A-1."""
    case = next(c for c in SEARCH if c["A_net"] == 4 and c["A_cfg"] == 4 and c["S"] == 25)
    d = hp.search_case_data(case)
    o, _ = _c_search(case, d)
    assert o["visit"][0, 0] == 50
    assert o["out_visits"][0].sum() == 25 - 4
    one = next(c for c in SEARCH if c["A_net"] == 2 and c["A_cfg"] == 2 and c["S"] == 1)
    o1, _ = _c_search(one, hp.search_case_data(one))
    assert o1["out_visits"][0].sum() == 0
    assert np.array_equal(o1["out_probs"][0], np.array([0.5, 0.5]))
