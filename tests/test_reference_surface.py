"""The reference's own hot-path tests, re-pointed at the drop-in.

Node known-answer tests follow /root/reference/tests/test_mcts.py:27-179 (CPU: `Node` is a host
class); the MCTS / Agent / network-shape tests follow tests/test_mcts.py:181-279,
tests/test_agent.py:17-100, tests/test_basic.py:24-94, tests/test_network.py:448-470 and need the GPU."""

import math

import numpy as np
import pytest

from cumind_b200 import Config
from cumind_b200.mcts import Node


# ---- Node (tests/test_mcts.py:27-179) ------------------------------------------------------------
def test_node_initial_state():
    node = Node(0.5)
    assert node.prior == 0.5 and node.visit_count == 0 and node.value_sum == 0.0
    assert node.children == {} and node.hidden_state is None
    assert not node.is_expanded()
    node.children[0] = Node(0.3)
    assert node.is_expanded()


def test_node_value_known_answers():
    node = Node(0.5)
    assert node.value() == 0.0
    node.visit_count, node.value_sum = 1, 0.8
    assert node.value() == 0.8
    node.visit_count, node.value_sum = 4, 2.0
    assert node.value() == 0.5
    node.value_sum = -1.0
    assert node.value() == -0.25


def test_ucb_score_known_answers():
    child = Node(0.3)
    assert child.ucb_score(10, 1.0) == float("inf")
    child.visit_count, child.value_sum = 5, 2.5
    expected = 0.5 + 1.0 * 0.3 * math.sqrt(10) / 6
    assert abs(child.ucb_score(10, 1.0) - expected) < 1e-6


def test_select_child_picks_larger_ucb_and_first_on_ties():
    parent = Node(0.5)
    parent.visit_count = 10
    c1, c2 = Node(0.3), Node(0.7)
    c1.visit_count, c1.value_sum = 2, 1.0
    c2.visit_count, c2.value_sum = 1, 0.8
    parent.children = {0: c1, 1: c2}
    u1, u2 = c1.ucb_score(10, 1.0), c2.ucb_score(10, 1.0)
    assert parent.select_child(1.0) == (0 if u1 > u2 else 1)
    fresh = Node(1.0)
    fresh.children = {0: Node(0.1), 1: Node(0.9)}
    assert fresh.select_child(1.0) == 0  # inf == inf -> first action


def test_node_expand_and_backup():
    node = Node(0.5)
    node.expand([0, 1], np.array([0.6, 0.4], dtype=np.float32), np.zeros(4, np.float32))
    assert len(node.children) == 2 and abs(node.children[0].prior - 0.6) < 1e-6
    node.expand([0], np.array([0.6, 0.4], dtype=np.float32), None)  # zip truncates
    node.backup(0.8)
    assert node.visit_count == 1 and node.value() == 0.8


# ---- Config (config.py) ---------------------------------------------------------------------------
def test_config_defaults_json_round_trip_and_validation(tmp_path):
    c = Config()
    assert (c.hidden_dim, c.num_blocks, c.num_simulations, c.c_puct, c.dirichlet_alpha, c.exploration_fraction) == (128, 2, 10, 1.0, 0.3, 0.25)
    assert c.observation_shape == (4,) and c.action_space_size == 2 and c.conv_channels == 32 and c.seed == 42
    c.num_simulations = 25
    p = tmp_path / "cfg" / "c.json"
    c.to_json(str(p))
    import json
    doc = json.loads(p.read_text())
    assert list(doc) == ["Network architecture", "Training", "MCTS", "Environment", "Self-Play", "Memory", "Data Types", "Device", "Other"]
    assert doc["MCTS"]["num_simulations"] == 25
    assert Config.from_json(str(p)) == c
    with pytest.raises(FileNotFoundError):
        Config.from_json(str(tmp_path / "missing.json"))
    for field, bad in (("hidden_dim", 0), ("action_space_size", -1), ("num_simulations", 0), ("learning_rate", 0.0), ("batch_size", 0)):
        with pytest.raises(ValueError, match=f"{field} must be positive"):
            Config(**{field: bad}).validate()
    with pytest.raises(ValueError, match="1D or 3D"):
        Config(observation_shape=(4, 4)).validate()
    with pytest.raises(ValueError, match="model_dtype"):
        Config(model_dtype="int8").validate()
    Config().validate()


# ---- GPU: MCTS / network / agent surface ------------------------------------------------------------
def _setup(hidden=16, A=2, obs=(4,), S=100):
    from cumind_b200 import CuMindNetwork, MCTS, Rngs, key

    config = Config()
    config.num_simulations, config.action_space_size, config.observation_shape, config.hidden_dim = S, A, obs, hidden
    key.seed(config.seed)
    rngs = Rngs(params=key())
    network = CuMindNetwork(observation_shape=config.observation_shape, action_space_size=config.action_space_size,
                            hidden_dim=config.hidden_dim, num_blocks=config.num_blocks, conv_channels=config.conv_channels, rngs=rngs)
    return MCTS(network, config), network, config


@pytest.mark.gpu
def test_mcts_search_returns_valid_policy():
    mcts, _, config = _setup()
    policy = mcts.search(np.ones(16, np.float32), add_noise=False)
    assert isinstance(policy, np.ndarray) and policy.dtype == np.float64 and policy.shape == (2,)
    assert np.isclose(policy.sum(), 1.0) and np.all(policy >= 0)
    noisy = mcts.search(np.ones(16, np.float32), add_noise=True)
    assert np.isclose(noisy.sum(), 1.0)


@pytest.mark.gpu
def test_mcts_three_actions_and_single_action_edge_case():
    mcts, network, config = _setup(A=3, obs=(8,))
    assert mcts.search(np.ones(16, np.float32), add_noise=False).shape == (3,)
    mcts2, network2, config2 = _setup(A=2)
    config2.action_space_size = 1          # tests/test_mcts.py:271-279: 2-logit network, 1 configured action
    from cumind_b200 import MCTS
    single = MCTS(network2, config2).search(np.ones(16, np.float32), add_noise=False)
    assert single.shape == (1,) and np.isclose(single[0], 1.0)


@pytest.mark.gpu
def test_network_shapes_and_errors():
    from cumind_b200 import CuMindNetwork

    _, net, config = _setup(hidden=32)
    h, lg, v = net.initial_inference(np.ones((2, 4), np.float32))
    assert tuple(h.shape) == (2, 32) and tuple(lg.shape) == (2, 2) and tuple(v.shape) == (2, 1)
    nx, rw, lg2, v2 = net.recurrent_inference(h, np.array([0, 1]))
    assert tuple(nx.shape) == (2, 32) and tuple(rw.shape) == (2, 1) and tuple(lg2.shape) == (2, 2) and tuple(v2.shape) == (2, 1)
    assert len(net.representation_network.encoder.layers) == config.num_blocks
    with pytest.raises(ValueError, match="Unsupported observation shape"):
        CuMindNetwork((4, 4), 2, 16, 2, 32, 0)
    img = CuMindNetwork((84, 84, 3), 4, 64, 2, 32, 0)
    h2, lg3, _ = img.initial_inference(np.ones((2, 84, 84, 3), np.float32))
    assert tuple(h2.shape) == (2, 64) and tuple(lg3.shape) == (2, 4)
    small = CuMindNetwork((10, 10, 4), 4, 16, 2, 32, 0)
    assert tuple(small.initial_inference(np.ones((3, 10, 10, 4), np.float32))[0].shape) == (3, 16)


@pytest.mark.gpu
def test_agent_surface():
    from cumind_b200 import Agent

    config = Config()
    config.num_simulations, config.action_space_size, config.observation_shape = 5, 2, (4,)
    agent = Agent(config)
    assert agent.config == config
    for attr in ("network", "target_network", "optimizer", "optimizer_state", "mcts"):
        assert hasattr(agent, attr)
    obs = np.ones(4)
    action, policy = agent.select_action(obs, training=True)
    assert isinstance(action, int) and isinstance(policy, np.ndarray) and len(policy) == 2 and np.isclose(policy.sum(), 1.0)
    a1, _ = agent.select_action(obs, training=False)
    a2, _ = agent.select_action(obs, training=False)
    assert a1 == a2  # eval mode is deterministic (tests/test_agent.py:46-58)
    state = agent.save_state()
    assert set(state) == {"network_state", "optimizer_state"}
    agent2 = Agent(config)
    agent2.load_state(state)
    assert np.allclose(agent.network.blob, agent2.network.blob, rtol=1e-6)
    assert np.array_equal(agent2.target_network.blob, agent2.network.blob)


@pytest.mark.gpu
def test_agent_with_image_observations():
    from cumind_b200 import Agent

    config = Config()
    config.observation_shape, config.action_space_size = (84, 84, 4), 4
    agent = Agent(config)
    action, probs = agent.select_action(np.ones((84, 84, 4)), training=True)
    assert isinstance(action, int) and 0 <= action < 4 and probs.shape == (4,)
    config8 = Config()
    config8.observation_shape = (8,)
    a, _ = Agent(config8).select_action(np.ones(8), training=True)
    assert isinstance(a, int)
