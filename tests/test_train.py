"""Train step: the NumPy oracle's hand-written backward pass against finite differences (CPU),
and the GPU step (torch autograd over the flat blob + cumind_adamw_step) against the oracle.
Tolerances: gradients rtol 1e-4 (fp32 GEMM reductions vs float64), AdamW update rtol 1e-5."""

import os

import numpy as np
import pytest

from oracle import network_np as nn
from oracle import train_np as tn


def _problem(B=24, A=3, H=16, L=2, K=3, D=5, seed=0):
    rng = np.random.default_rng(seed)
    params = nn.random_params((D,), A, H, L, 32, seed, bias_scale=0.1)
    obs = rng.standard_normal((B, D)).astype(np.float32)
    actions = rng.integers(0, A, size=(B, K)).astype(np.int32)
    pol = rng.dirichlet([0.5] * A, size=B).astype(np.float32)
    targets = {"values": rng.standard_normal(B).astype(np.float32), "rewards": rng.standard_normal((B, K)).astype(np.float32), "policies": pol}
    return params, obs, actions, targets, (D, A, H, L, K)


def test_oracle_backward_matches_finite_differences():
    params, obs, actions, targets, (D, A, H, L, K) = _problem()
    total, losses, grads = tn.loss_and_grad(params, obs, actions, targets, K)
    assert abs(total - sum(losses.values())) < 1e-12
    flat = nn.flatten_params(params, (D,)).astype(np.float64)
    gflat = nn.flatten_params(grads, (D,)) if False else np.concatenate([np.asarray(a, np.float64).reshape(-1) for a in _leaves(grads, (D,))])
    rng = np.random.default_rng(1)
    from cumind_b200 import params as P

    for idx in rng.choice(flat.size, size=40, replace=False):
        def loss_at(delta):
            f2 = flat.copy()
            f2[idx] += delta
            tree = P.unflatten(f2.astype(np.float32), (D,), A, H, L, 32)
            tree = _as64(tree, f2, (D,), A, H, L)
            return tn.loss_and_grad(tree, obs, actions, targets, K)[0]
        eps = 1e-5
        fd = (loss_at(eps) - loss_at(-eps)) / (2 * eps)
        assert abs(fd - gflat[idx]) <= 1e-6 + 1e-4 * abs(fd), (idx, fd, gflat[idx])


def test_oracle_adamw_equals_an_independent_implementation():
    """The optimiser semantics are pinned to something not written here: torch.optim.AdamW (decoupled weight decay,
    bias-corrected moments, eps outside the square root) is the update optax.adamw documents (agent.py:48:
    optax.adamw(learning_rate, weight_decay)); the oracle's adamw must follow it step for step in float64."""
    import torch

    from oracle import train_np as tn

    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(257)
    lr, wd = 3e-3, 1e-2
    tp = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.AdamW([tp], lr=lr, weight_decay=wd, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 40):
        g = rng.standard_normal(257) * (10.0 ** rng.integers(-3, 2))
        tp.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
        p, m, v = tn.adamw(p, g, m, v, step, lr, wd)
        assert np.allclose(p, tp.detach().numpy(), rtol=1e-12, atol=1e-14), step


def _leaves(tree, obs_shape):
    from cumind_b200 import params as P

    A, H = tree["dynamics_network"]["action_embedding"]["embedding"].shape
    L = len(tree["dynamics_network"]["layers"])
    return [P._get(tree, path) for path, _ in P._schema(obs_shape, A, H, L, 32)]


def _as64(tree, flat64, obs_shape, A, H, L):
    """Parameter tree whose leaves are float64 views of flat64 (so a 1e-5 perturbation is not rounded away)."""
    from cumind_b200 import params as P

    out, o = {}, 0
    for path, shape in P._schema(obs_shape, A, H, L, 32):
        n = int(np.prod(shape))
        P._set(out, path, flat64[o:o + n].reshape(shape))
        o += n
    return out


@pytest.mark.gpu
def test_gpu_train_step_matches_the_oracle():
    import torch

    from cumind_b200 import Agent, Config
    from cumind_b200 import params as P
    from cumind_b200.trainer import MemoryBuffer, Trainer

    params, obs, actions, targets, (D, A, H, L, K) = _problem(B=64, A=2, H=32, L=2, K=5, D=4, seed=3)
    cfg = Config(hidden_dim=H, num_blocks=L, action_space_size=A, observation_shape=(D,), num_unroll_steps=K, learning_rate=1e-2,
                 weight_decay=1e-4, num_simulations=5)
    agent = Agent(cfg)
    agent.load_state({"network_state": params, "optimizer_state": {"count": 0}})
    trainer = Trainer(agent, MemoryBuffer(10), cfg)
    want_total, want_losses, want_grads = tn.loss_and_grad(params, obs, actions, targets, K)
    losses = trainer.compute_losses(obs, actions, targets)
    for k in want_losses:
        assert abs(float(losses[k]) - want_losses[k]) <= 1e-5 * max(1.0, abs(want_losses[k])), k
    before = agent.network.blob.copy()
    info = trainer.train_step((obs, actions, targets))
    assert abs(info["total_loss"] - want_total) <= 1e-5 * max(1.0, abs(want_total))
    g_gpu = trainer.last_grad.cpu().numpy().astype(np.float64)
    g_ref = np.concatenate([np.asarray(a, np.float64).reshape(-1) for a in _leaves(want_grads, (D,))])
    scale = np.abs(g_ref).max()
    assert np.all(np.abs(g_gpu - g_ref) <= 1e-4 * np.abs(g_ref) + 1e-6 * scale)
    want_p, _, _ = tn.adamw(before.astype(np.float64), g_ref, np.zeros_like(g_ref), np.zeros_like(g_ref), 1, cfg.learning_rate, cfg.weight_decay)
    after = agent.network.blob.astype(np.float64)
    assert np.all(np.abs(after - want_p) <= 1e-5 * np.abs(want_p) + 1e-6)
    assert not np.array_equal(after, before)
    # the search kernels see the new parameters
    h_new, _, _ = agent.network.initial_inference(obs[:4])
    net2 = agent.network.clone()
    assert np.array_equal(net2.initial_inference(obs[:4])[0].cpu().numpy(), h_new.cpu().numpy())
    # a few more steps reduce the loss on the same batch
    first = info["total_loss"]
    for _ in range(30):
        info = trainer.train_step((obs, actions, targets))
    assert info["total_loss"] < first
    trainer.sync_optimizer_state()
    assert agent.optimizer_state["count"] == 31 and int(trainer._step_dev.item()) == 31
    # the CUDA-graph replay and the eager path take the same steps
    agent_e = Agent(cfg)
    agent_e.load_state({"network_state": params, "optimizer_state": {"count": 0}})
    eager = Trainer(agent_e, MemoryBuffer(10), cfg)
    eager.use_cuda_graph = False
    for _ in range(31):
        info_e = eager.train_step((obs, actions, targets))
    # (Adam normalises every gradient, so round-off in near-zero gradients moves individual weights by ~lr;
    # the trajectories agree in loss, and step 1 of the graph path was checked against the oracle above)
    assert abs(info_e["total_loss"] - info["total_loss"]) <= 0.02 * abs(info["total_loss"])
    assert int(eager._step_dev.item()) == 31


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1024, 2, 64, 2, 5, 4), (203, 4, 64, 2, 5, 4), (37, 3, 32, 1, 1, 7), (64, 2, 64, 3, 4, 11), (9, 14, 32, 2, 14, 64)],
                         ids=lambda s: "B%dA%dH%dL%dK%dD%d" % s)
def test_fused_loss_and_gradient_kernel_matches_the_oracle(shape):
    """cumind_train_loss_grad (train_fused.cu: the whole unroll forward + backward in one kernel) against the float64
    restatement oracle/train_np.py: the three losses within 1e-5, every gradient entry within 1e-4 relative (+ 1e-6 of the
    largest entry), at the BASELINE configs[4] shape (batch 1024, K=5, H=64) and at ragged / other shapes; and the
    result is reproducible (fixed summation order: two runs give identical bits)."""
    import torch

    from cumind_b200 import Agent, Config
    from cumind_b200.trainer import MemoryBuffer, Trainer

    B, A, H, L, K, D = shape
    params, obs, actions, targets, _ = _problem(B=B, A=A, H=H, L=L, K=K, D=D, seed=B + H)
    cfg = Config(hidden_dim=H, num_blocks=L, action_space_size=A, observation_shape=(D,), num_unroll_steps=K, num_simulations=3)
    agent = Agent(cfg)
    agent.load_state({"network_state": params, "optimizer_state": {"count": 0}})
    tr = Trainer(agent, MemoryBuffer(4), cfg)
    assert tr.use_fused_kernel
    dev = tr._flat.device
    t_obs, t_act = torch.as_tensor(obs).to(dev), torch.as_tensor(actions).to(dev)
    t_tg = {k: torch.as_tensor(v).to(dev) for k, v in targets.items()}
    out, grad = tr._fwd_bwd(t_obs, t_act, t_tg)
    want_total, want_losses, want_grads = tn.loss_and_grad(params, obs, actions, targets, K)
    for k in want_losses:
        assert abs(float(out[k]) - want_losses[k]) <= 1e-5 * max(1.0, abs(want_losses[k])), (k, float(out[k]), want_losses[k])
    assert abs(float(out["total_loss"]) - want_total) <= 1e-5 * max(1.0, abs(want_total))
    g_gpu = grad.cpu().numpy().astype(np.float64)
    g_ref = np.concatenate([np.asarray(a, np.float64).reshape(-1) for a in _leaves(want_grads, (D,))])
    scale = np.abs(g_ref).max()
    bad = np.abs(g_gpu - g_ref) > 1e-4 * np.abs(g_ref) + 1e-6 * scale
    assert not bad.any(), (int(bad.sum()), np.where(bad)[0][:10], g_gpu[bad][:5], g_ref[bad][:5])
    first = grad.clone()
    _, again = tr._fwd_bwd(t_obs, t_act, t_tg)
    assert torch.equal(first, again)
    # and the autograd path (what image encoders use) computes the same thing
    tr.use_fused_kernel = False
    out_a, grad_a = tr._fwd_bwd(t_obs, t_act, t_tg)
    assert np.all(np.abs(grad_a.cpu().numpy() - g_gpu) <= 2e-4 * np.abs(g_ref) + 2e-6 * scale)


@pytest.mark.gpu
def test_trainer_runs_from_replay_with_bootstrapped_targets_and_image_observations():
    from cumind_b200 import Agent, Config
    from cumind_b200.envs import CartPoleVec
    from cumind_b200.self_play import SelfPlay
    from cumind_b200.trainer import MemoryBuffer, Trainer

    cfg = Config(hidden_dim=32, num_simulations=6, batch_size=16, td_steps=3, num_unroll_steps=3, min_memory_size=4)
    agent = Agent(cfg)
    memory = MemoryBuffer(200)
    per_env = SelfPlay(agent, memory).run_batch(CartPoleVec(32, seed=2), num_steps=25, philox_seed=4)
    for eps in per_env:                       # unfinished tails are trajectories too
        for ep in eps:
            if len(ep) > 1:
                memory.add(ep)
    trainer = Trainer(agent, memory, cfg)
    info = trainer.train_step()
    assert np.isfinite(info["total_loss"]) and set(info) >= {"total_loss", "train/value_loss", "train/policy_loss", "train/reward_loss"}
    # image observations: conv trunk + BatchNorm (running statistics stay fixed)
    icfg = Config(hidden_dim=16, observation_shape=(12, 12, 3), action_space_size=3, conv_channels=8, num_unroll_steps=2, num_simulations=4)
    iagent = Agent(icfg)
    itr = Trainer(iagent, MemoryBuffer(4), icfg)
    rng = np.random.default_rng(0)
    batch = (rng.random((8, 12, 12, 3)).astype(np.float32), rng.integers(0, 3, size=(8, 2)).astype(np.int32),
             {"values": rng.standard_normal(8).astype(np.float32), "rewards": rng.standard_normal((8, 2)).astype(np.float32),
              "policies": rng.dirichlet([1.0] * 3, size=8).astype(np.float32)})
    before = iagent.network.blob.copy()
    l0 = itr.train_step(batch)["total_loss"]
    for _ in range(20):
        l1 = itr.train_step(batch)["total_loss"]
    assert np.isfinite(l1) and l1 < l0
    from cumind_b200.trainer import trainable_mask
    mask = trainable_mask(iagent.network._dims).astype(bool)
    after = iagent.network.blob
    assert np.array_equal(after[~mask], before[~mask]) and not np.array_equal(after[mask], before[mask])


@pytest.mark.gpu
def test_distributed_two_graph_step_equals_the_eager_step():
    """torch.distributed initialised (world_size 1 over gloo, so it runs on one GPU): the step replayed as two CUDA graphs
    around an eager all-reduce follows exactly the same training trajectory as the eager step."""
    import torch
    import torch.distributed as dist

    from cumind_b200 import Agent, Config
    from cumind_b200.trainer import MemoryBuffer, Trainer

    params, obs, actions, targets, (D, A, H, L, K) = _problem(B=32, A=3, H=16, L=2, K=3, D=5, seed=9)
    cfg = Config(hidden_dim=H, num_blocks=L, action_space_size=A, observation_shape=(D,), num_unroll_steps=K, learning_rate=1e-2,
                 weight_decay=1e-4, num_simulations=4)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29731")
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        runs = []
        for graphs in (True, False):
            agent = Agent(cfg)
            agent.load_state({"network_state": params, "optimizer_state": {"count": 0}})
            tr = Trainer(agent, MemoryBuffer(4), cfg)
            tr.use_cuda_graph = graphs
            losses = [tr.train_step((obs, actions, targets))["total_loss"] for _ in range(6)]
            runs.append((losses, agent.network.blob.copy()))
            if graphs:
                assert isinstance(tr._graph, tuple) and len(tr._graph) == 2   # the two-graph path was taken
        assert runs[0][0] == runs[1][0]
        assert np.array_equal(runs[0][1], runs[1][1])
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_checkpoint_round_trip_resumes_the_same_trajectory(tmp_path):
    """save -> load -> the next steps are bit-equal to an uninterrupted run: the AdamW moments and the step count travel
    with the checkpoint (Trainer.load_checkpoint reloads them in place), Agent.save_state pulls them from the device, and a
    direct Agent.load_state is picked up by the next train step."""
    from cumind_b200 import Agent, Config
    from cumind_b200.trainer import MemoryBuffer, Trainer

    params, obs, actions, targets, (D, A, H, L, K) = _problem(B=48, A=2, H=32, L=2, K=4, D=4, seed=5)
    cfg = Config(hidden_dim=H, num_blocks=L, action_space_size=A, observation_shape=(D,), num_unroll_steps=K, learning_rate=1e-2,
                 weight_decay=1e-4, num_simulations=4, checkpoint_root_dir=str(tmp_path))
    batch = (obs, actions, targets)
    a1 = Agent(cfg)
    a1.load_state({"network_state": params, "optimizer_state": {"count": 0}})
    t1 = Trainer(a1, MemoryBuffer(4), cfg)
    for _ in range(3):
        t1.train_step(batch)
    path = t1.save_checkpoint(3)
    state = a1.save_state()                                   # not stale: the moments come from the device
    assert state["optimizer_state"]["count"] == 3 and np.any(state["optimizer_state"]["mu"] != 0)
    for _ in range(2):
        t1.train_step(batch)
    want = a1.network.blob.copy()
    a2 = Agent(cfg)
    t2 = Trainer(a2, MemoryBuffer(4), cfg)
    t2.train_step(batch)                                      # some unrelated history, then resume from the file
    t2.load_checkpoint(path)
    assert t2._count == 3 and int(t2._step_dev.item()) == 3
    for _ in range(2):
        t2.train_step(batch)
    assert np.array_equal(a2.network.blob, want)
    # Agent.load_state directly: the trainer notices at the next step
    a3 = Agent(cfg)
    t3 = Trainer(a3, MemoryBuffer(4), cfg)
    t3.train_step(batch)
    a3.load_state(state)
    for _ in range(2):
        t3.train_step(batch)
    assert np.array_equal(a3.network.blob, want)


@pytest.mark.gpu
def test_run_training_loop_follows_the_reference_schedule(tmp_path):
    """trainer.py:44-81: train every train_frequency episodes, target network every target_update_frequency train steps,
    checkpoint every checkpoint_interval episodes."""
    from cumind_b200 import Agent, Config
    from cumind_b200.trainer import MemoryBuffer, Trainer
    from tests.test_selfplay import _Single

    cfg = Config(hidden_dim=32, num_simulations=4, batch_size=8, td_steps=2, num_unroll_steps=2, min_memory_size=1, min_memory_pct=0.0,
                 num_episodes=6, train_frequency=2, target_update_frequency=2, checkpoint_interval=3, checkpoint_root_dir=str(tmp_path))
    agent = Agent(cfg)
    tr = Trainer(agent, MemoryBuffer(100), cfg)
    before = agent.target_network.blob.copy()
    tr.run_training_loop(_Single())
    assert tr.train_step_count == 3 and len(tr.memory) == 6
    assert not np.array_equal(agent.target_network.blob, before)      # updated at train step 2
    assert sorted(os.listdir(tr.checkpoint_dir)) == ["episode_00003.pkl", "episode_00006.pkl"]


@pytest.mark.gpu
def test_fused_peer_memory_allreduce_adamw_on_two_gpus():
    """csrc/p2p_adamw.cu: the gradient all-reduce fused into AdamW over NVLink peer memory, one process per GPU (torchrun,
    world_size 2): replicas stay bit-identical, the fused sum equals ncclAllReduce within 1e-6, the CUDA-graph step equals the
    eager one.  Needs two GPUs in the box (skipped on a single-GPU box; tools/test_p2p_train.py is the script)."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", os.path.join(root, "tools", "test_p2p_train.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
