#!/usr/bin/env python
"""bench.py — batched MCTS simulations/sec of the CuMind search hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--trees B] [--sims S]

One "step" = one full search pass over one batch of synthetic trees: initial_inference on the
resident observations, root prediction + Dirichlet mix, S simulations (select -> dynamics ->
prediction -> expand -> backup), visit-count normalisation.  Workload at every N: BASELINE.json
configs[1] per GPU (B=4096 CartPole-shaped trees, S=50, fp32-exact network); trees shard by
environment across ranks with no collective in the data path ("weak" scaling).

`value`    = all trees x simulations of all ranks / device time of the K steps (CUDA events on the
             launching stream, max over ranks), inputs resident in HBM.
`e2e`      = the same metric through Agent.select_actions -> cumind_select_actions_host: host (pinned)
             observations + noise in, host actions + probabilities out, copies inside the timed region.
`roofline` = HBM roofline of the dominant kernel (the fused search kernel): algorithmic bytes per
             simulation x simulations per launch / its CUDA-event duration, against
             MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` = the oracle's Python port of the reference search (one tree per call, NumPy/BLAS
             network), one process per host core, bounded sample, timed in this run on rank 0.
--impl reference times only that CPU path (rank 0; other ranks exit 0).
"""

from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ---- workload (BASELINE.json configs[1]) -------------------------------
OBS_SHAPE, A, H, L, CC = (4,), 2, 64, 2, 32
C_PUCT, ALPHA, FRAC = 1.25, 0.3, 0.25
WORKLOAD, MODEL_DTYPE = "cartpole", "float32"


def set_workload(name: str):
    """cartpole = BASELINE configs[1] (the bench line); atari = configs[2] (3x84x84 image observations, NHWC
    (84,84,3), A=4, residual conv trunk, bf16 tensor-core mode) — measured for profiles/, not the bench line."""
    global OBS_SHAPE, A, H, L, CC, WORKLOAD, MODEL_DTYPE
    WORKLOAD = name
    if name == "atari":
        OBS_SHAPE, A, H, L, CC, MODEL_DTYPE = (84, 84, 3), 4, 64, 2, 32, "bfloat16"
    elif name == "atari-literal":
        OBS_SHAPE, A, H, L, CC, MODEL_DTYPE = (3, 84, 84), 4, 64, 2, 32, "bfloat16"
    elif name == "cartpole-h128":   # the shipped checkpoint's width (configuration.json: hidden_dim 128), SURVEY §8(d) config #2
        H = 128
    elif name != "cartpole":
        raise ValueError(name)


def set_workload_raw(state):
    global OBS_SHAPE, A, H, L, CC, WORKLOAD, MODEL_DTYPE
    OBS_SHAPE, A, H, L, CC, WORKLOAD, MODEL_DTYPE = state


def encoder_flops_per_obs() -> float:
    if len(OBS_SHAPE) == 1:
        return 2.0 * OBS_SHAPE[0] * H + 2.0 * (L - 1) * H * H
    hh, ww, c = OBS_SHAPE
    npix = ((hh - 1) // 2 + 1) * ((ww - 1) // 2 + 1)
    return 2.0 * npix * 9 * c * CC + L * 2 * (2.0 * npix * 9 * CC * CC) + 2.0 * CC * H
PARAM_SEED, OBS_SEED, NOISE_SEED = 42, 0, 1
FALLBACK_HBM_GBS = 6650.0


def synthetic_params(seed: int):
    """Random-init weights of the named architecture (lecun-normal-scale kernels, zero biases, identity BatchNorm), the
    nested layout of cumind_b200/params.py.  Same draw order as the parity tests' generator, so both arms and the
    tests search the same networks; the native arm does not import oracle/."""
    rng = np.random.default_rng(seed)
    f32 = np.float32

    def dense(i, o):
        k = (rng.standard_normal((i, o)) / np.sqrt(i)).astype(f32)
        return {"kernel": k, "bias": (rng.standard_normal(o) * 0.0).astype(f32)}   # signed zeros, as the tests' generator draws them

    def conv(i, o):
        k = (rng.standard_normal((3, 3, i, o)) / np.sqrt(9 * i)).astype(f32)
        return {"kernel": k, "bias": (rng.standard_normal(o) * 0.0).astype(f32)}

    def bn(c):
        return {"scale": np.ones(c, f32), "bias": np.zeros(c, f32), "mean": np.zeros(c, f32), "var": np.ones(c, f32)}

    if len(OBS_SHAPE) == 1:
        layers, cur = {}, OBS_SHAPE[0]
        for i in range(L):
            layers[i] = dense(cur, H)
            cur = H
        enc = {"layers": layers}
    else:
        enc = {"initial_conv": conv(OBS_SHAPE[2], CC),
               "residual_blocks": {i: {"conv1": conv(CC, CC), "bn1": bn(CC), "conv2": conv(CC, CC), "bn2": bn(CC)} for i in range(L)},
               "final_dense": dense(CC, H)}
    dyn = {"action_embedding": {"embedding": (rng.standard_normal((A, H)) / np.sqrt(H)).astype(f32)},
           "layers": {i: dense(H, H) for i in range(L)}, "reward_head": dense(H, 1)}
    pred = {"policy_head": dense(H, A), "value_head": dense(H, 1)}
    return {"representation_network": {"encoder": enc}, "dynamics_network": dyn, "prediction_network": pred}


def make_inputs(B: int, rank: int):
    from cumind_b200 import params as P

    params = synthetic_params(PARAM_SEED)
    blob = P.flatten(params, OBS_SHAPE, A, H, L, CC)
    rng_obs = np.random.default_rng(OBS_SEED + 1000 * rank)
    if len(OBS_SHAPE) == 1:
        obs = rng_obs.standard_normal((B,) + OBS_SHAPE).astype(np.float32)
    else:   # image observations are emulator frames: uint8, scaled to [0, 1] (the device-resident leg holds the floats)
        obs = rng_obs.integers(0, 256, size=(B,) + OBS_SHAPE, dtype=np.uint8).astype(np.float32) / np.float32(255)
    noise = np.random.default_rng(NOISE_SEED + 1000 * rank).dirichlet([ALPHA] * A, size=B)
    return params, blob, obs, noise


def algorithmic_bytes_per_sim(dbar: float) -> float:
    """select d*(16A+8) + expand 8H+16A+4 + backup 24*d."""
    return dbar * (16 * A + 8) + 8 * H + 16 * A + 4 + 24 * dbar


# ---- CPU baseline: the oracle's Python port, one process per core -----------------------------------
_W = {}


def _cpu_init(S: int, max_trees: int):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    from oracle import mcts_port as port
    from oracle import network_np as nn

    params = synthetic_params(PARAM_SEED)
    net = nn.NumpyCuMindNetwork(params, OBS_SHAPE, blas=True)
    rng = np.random.default_rng(os.getpid())
    obs = rng.standard_normal((max_trees,) + OBS_SHAPE).astype(np.float32)
    noise = rng.dirichlet([ALPHA] * A, size=max_trees)
    hidden = net.representation_network(obs)
    port.run_search(net, hidden[0], 5, A, C_PUCT, FRAC, noise[0])  # warm caches
    _W.update(port=port, net=net, hidden=hidden, noise=noise, S=S, next=0, max=max_trees)


def _cpu_work(seconds: float):
    w = _W
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        i = w["next"] % w["max"]
        w["port"].run_search(w["net"], w["hidden"][i], w["S"], A, C_PUCT, FRAC, w["noise"][i])
        w["next"] += 1
        n += 1
    return n, time.perf_counter() - t0


class CpuPool:
    """One worker process per host core, each running single-tree searches (the reference's shape of work)."""

    def __init__(self, S: int):
        self.S = S
        self.cores = len(os.sched_getaffinity(0))
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init, initargs=(S, 256))
        self.pool.map(_cpu_work, [0.05] * self.cores)

    def run(self, seconds: float):
        res = self.pool.map(_cpu_work, [seconds] * self.cores, chunksize=1)
        trees = sum(r[0] for r in res)
        wall = max(r[1] for r in res)
        return trees, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(S: int, seconds: float):
    pool = CpuPool(S)
    try:
        trees, wall = pool.run(seconds)
    finally:
        pool.close()
    return {"value": trees * S / wall, "unit": "simulations/s", "cores": pool.cores, "kind": "port",
            "sample": f"{trees} single-tree searches (S={S}, A={A}, H={H}) over {pool.cores} processes, {wall:.1f} s wall; " + CPU_NOTE}


CPU_NOTE = ("oracle/mcts_port.py (Python port of mcts.py) + NumPy/BLAS network, one tree per call, one process per core; "
            "no eager-JAX dispatch or log formatting, so faster than the real reference")


def c_oracle_rate(S: int, trees: int = 512):
    from oracle import liboracle as lo

    _, blob, obs, noise = make_inputs(trees, 0)
    desc = lo.make_desc(OBS_SHAPE, A, H, L)
    h, _, _ = lo.initial_inference(desc, blob, obs)
    t0 = time.perf_counter()
    lo.search(desc, blob, h, S, A, C_PUCT, FRAC, noise)
    return trees * S / (time.perf_counter() - t0)


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int, period: float = 0.025):
        super().__init__(daemon=True)
        self.period, self.samples, self.stop_flag, self.live = period, [], False, False
        self.max_mhz, self.h, self.nv = None, None, None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.live:
                    self.samples.append((mhz, int(reasons)))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        mask = 0
        for _, r in self.samples:
            mask |= r
        reasons = [n for b, n in names.items() if mask & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(m for m, _ in self.samples), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


# ---- the arms ----------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    """The reference's CPU implementation of the path (Python port; /root/reference cannot travel)."""
    if rank != 0:
        return
    S, K, W = args.sims, args.steps, args.warmup
    box = min(4.0, max(0.3, 150.0 / (K + W)))  # whole run stays within a few minutes
    t0 = time.perf_counter()
    pool = CpuPool(S)
    try:
        for _ in range(W):
            pool.run(box)
        trees_total, wall_total = 0, 0.0
        for _ in range(K):
            trees, wall = pool.run(box)
            trees_total += trees
            wall_total += wall
    finally:
        pool.close()
    value = trees_total * S / wall_total
    base = {"value": value, "unit": "simulations/s", "cores": pool.cores, "kind": "port",
            "sample": f"{K} steps x {box:.2f} s time box per core: {trees_total} single-tree searches (S={S}, A={A}, H={H}); " + CPU_NOTE}
    line = {"impl": "reference", "metric": "batched MCTS simulations/sec", "value": value, "unit": "simulations/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * wall_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args, per_gpu=args.trees), "cpu_baseline": base,
            "e2e": {"value": value, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu: int, total_trees: int = 0):
    title = ("BASELINE configs[3]: large-batch search, %d CartPole-shaped trees sharded over the GPUs" % total_trees if total_trees == 65536 and WORKLOAD == "cartpole" else
             "CartPole-shaped trees, %d in total sharded over the GPUs (strong scaling)" % total_trees if total_trees > 0 and WORKLOAD == "cartpole" else
             "BASELINE configs[1]: CartPole-shaped synthetic self-play" if WORKLOAD == "cartpole" and per_gpu == 4096 else
             "CartPole-shaped synthetic self-play (the configs[1] network at another batch size)" if WORKLOAD == "cartpole" else
             "CartPole-shaped, shipped-checkpoint width" if WORKLOAD == "cartpole-h128" else f"BASELINE configs[2] ({WORKLOAD}): image observations")
    mode = "fp32-exact network" if MODEL_DTYPE == "float32" else "bf16 tensor-core (tcgen05) network mode (search kernel: see plan)"
    return {"workload": f"{title}, B={per_gpu} trees per GPU, S={args.sims} simulations, "
                        f"obs {OBS_SHAPE}, A={A}, hidden_dim={H}, num_blocks={L}, {mode}, injected Dirichlet({ALPHA}) noise, c_puct={C_PUCT}",
            "trees_per_gpu": per_gpu, "num_simulations": args.sims, "parallelism": f"trees sharded by environment x{args.gpus}, no collective",
            "l2": "between timed steps a 256 MiB device buffer is written (L2 flush) and a second 256 MiB buffer is read (drains the dirty lines); per-step CUDA events"}


def measure_search(args, rank: int, world: int, local_rank: int, dist, sampler, K: int, W: int, sustain: float, total_trees: int,
                   trees: int):
    """One workload (the globals set by set_workload): device-resident steps, the search kernel alone, the sustained leg and
    the end-to-end path.  Every rank runs it; the returned dict is complete on every rank (timings are max over ranks)."""
    import torch

    from cumind_b200 import Config, _native
    from cumind_b200 import params as P
    from cumind_b200.agent import Agent

    S = args.sims
    tree_base = rank * trees
    if total_trees > 0:
        from cumind_b200.dist import shard_bounds

        lo_, hi_ = shard_bounds(total_trees, rank, world)
        trees, tree_base = hi_ - lo_, lo_
    B = trees
    _, blob, obs, noise = make_inputs(B, rank)
    cfg = Config(hidden_dim=H, num_blocks=L, num_simulations=S, action_space_size=A, observation_shape=OBS_SHAPE, c_puct=C_PUCT,
                 dirichlet_alpha=ALPHA, exploration_fraction=FRAC, conv_channels=CC, model_dtype=MODEL_DTYPE)
    agent = Agent(cfg)
    agent.network.load_state(P.unflatten(blob, OBS_SHAPE, A, H, L, CC))
    agent.update_target_network()
    agent.mcts.lanes_per_tree = args.lanes
    agent.mcts.schedule = {"auto": 0, "warp": 1, "team": 2, "tensor": 4}[args.schedule]
    agent.mcts.teams_per_cta, agent.mcts.net_warps_per_team, agent.mcts.trees_per_team = args.teams, args.net_warps, args.trees_per_team
    net, mcts = agent.network, agent.mcts
    dev = torch.device("cuda", local_rank)
    d_obs = torch.as_tensor(obs).to(dev)
    d_noise = torch.as_tensor(noise).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    drain = torch.zeros(64 << 20, dtype=torch.int32, device=dev)   # 256 MiB read after the write: evicts the dirty lines

    def flush_l2(tag: int):
        """Write a buffer larger than L2, then read another one so the write-backs of the dirty lines are
        over before the next timed step starts (they would otherwise be billed to it)."""
        flush.fill_(tag & 0xFF)
        drain.sum()

    def step_dev():
        # encode -> search through ONE C call (cumind_search_from_obs): no host work between the two launches
        return mcts.search_from_obs(d_obs, d_noise, tree_index_base=tree_base)

    for _ in range(max(W, 3)):
        step_dev()
        flush_l2(1)
    torch.cuda.synchronize()
    if args.phase_clocks:
        if rank == 0:
            print(json.dumps(phase_clocks(mcts, net.representation_network(d_obs), d_noise, tree_base, S)), flush=True)
        return None

    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    k_start = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    k_end = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _native.launch_count()
    if sampler is not None:
        sampler.live = True
    wall0 = time.perf_counter()
    for i in range(K):
        starts[i].record()
        step_dev()
        ends[i].record()
        flush_l2(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = _native.launch_count() - launches0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = sum(step_ms)
    # the dominant kernel alone (roofline): K launches of the search on the resident root states, same L2 hygiene
    hidden = net.representation_network(d_obs)
    for i in range(K):
        k_start[i].record()
        _search_device(mcts, hidden, d_noise, tree_base)
        k_end[i].record()
        flush_l2(i)
    torch.cuda.synchronize()
    kern_ms = [s.elapsed_time(e) for s, e in zip(k_start, k_end)]
    enc_ms = [max(a - b, 0.0) for a, b in zip(step_ms, kern_ms)]
    # sustained leg: the same step back to back for >= `sustain` seconds (clocks under load for the samplers)
    sustained = None
    if sustain > 0:
        n_sus, s0, s1 = 0, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_sus = time.perf_counter()
        s0.record()
        while time.perf_counter() - t_sus < sustain:
            for _ in range(50):
                step_dev()
            n_sus += 50
            torch.cuda.synchronize()
        s1.record()
        torch.cuda.synchronize()
        sus_ms = s0.elapsed_time(s1)
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_sus, "ms_per_step": sus_ms / n_sus,
                     "note": "the timed step back to back without the L2 flush (the trees stay L2-resident); not the headline"}
    if sampler is not None:
        sampler.live = False
    dump = mcts.dump_trees()
    dbar = float(dump["depth_sum"].mean()) / S
    visits_ok = bool(np.all(dump["visit"][:, 0] == 2 * S))

    # ---- end to end through the public API: host buffers in, host results out -------------------
    # the caller (an environment loop) writes its observations and noise into the agent's pinned host buffers; every
    # step the GPU reads them from there (H2D inside the timed region) and writes actions + probabilities back (D2H)
    hb = agent.host_buffers(B)
    if len(OBS_SHAPE) == 3:   # frames cross PCIe as the uint8 the emulator produced; the device scales them (same floats)
        hb["obs_u8"][...] = np.rint(obs * np.float32(255)).astype(np.uint8)
        assert np.array_equal(hb["obs_u8"].astype(np.float32) / np.float32(255), obs)
        obs = hb["obs_u8"]
    else:
        hb["obs"][...] = obs
        obs = hb["obs"]
    hb["noise"][...] = noise
    noise = hb["noise"]
    for _ in range(3):
        agent.select_actions(obs, training=False, noise=noise)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0 = time.perf_counter()
    for _ in range(K):
        actions, probs = agent.select_actions(obs, training=False, noise=noise)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0

    t = torch.tensor([total_ms, e2e_s * 1e3, statistics.mean(kern_ms), statistics.mean(enc_ms)], dtype=torch.float64, device=dev)
    if sustained is not None:
        ts = torch.tensor([sustained["ms_per_step"]], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained["ms_per_step"] = float(ts.item())
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_mean, enc_ms_mean = (float(x) for x in t.cpu())
    t_trees = torch.tensor([B], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(t_trees)
    sims_per_step = int(t_trees.item()) * S
    value = sims_per_step * K / (total_ms * 1e-3)
    e2e_value = sims_per_step * K / (e2e_ms * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)")
    bytes_per_launch = algorithmic_bytes_per_sim(dbar) * B * S
    achieved = bytes_per_launch / (kern_ms_mean * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and WORKLOAD == "cartpole" and B == 4096:
        traffic = json.load(open(tpath)).get("search_fused_dram_bytes_per_launch")
    plan = mcts.plan(B)
    kernel_name = {"team": "search_team_lw_kernel" if plan.get("layer_warpgroup_resident_k") else "search_team_kernel", "warp": "search_fused_kernel",
                   "tensor-core": "search_tc_kernel"}.get(plan["schedule"], "stepwise")
    # SM-side view of the same kernel: the fp32 fused multiply-adds of the in-search network against the FFMA peak of the part
    # (148 SMs x 128 lanes x 2 flop x SM clock; MEASURED_PEAKS.json has no fp32 entry)
    fma_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
    fma_tflops = (2.0 * L * H * H + 2.0 * H * (A + 1)) * B * S / (kern_ms_mean * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel_name, "kernel_ms": kern_ms_mean, "algorithmic_bytes_per_sim": algorithmic_bytes_per_sim(dbar),
                "mean_path_length": dbar, "peak_source": peak_src, "encoder_ms": enc_ms_mean,
                "kernel_ms_how": "K launches of the search kernel alone after the timed steps (CUDA events, same L2 flush); encoder_ms = step - kernel",
                "sm_side": {"unit": "TFLOP/s", "achieved": fma_tflops, "peak": fma_peak, "frac": fma_tflops / fma_peak,
                            "what": "fp32 FMA work of dynamics + prediction per simulation (2LH^2 + 2H(A+1) flop) / kernel time, against 148 SMs x 128 FFMA lanes x 2 x max SM clock"}}
    if WORKLOAD.startswith("atari"):
        # image workload: the conv trunk dominates and is tensor-pipe bound
        tf_peak = float(peaks.get("bf16_tflops", 1590.0))
        tf = encoder_flops_per_obs() * B / (enc_ms_mean * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak, "traffic": None,
                    "kernel": "conv_tc kernels (all launches of one initial_inference)", "kernel_ms": enc_ms_mean,
                    "flops_per_obs": encoder_flops_per_obs(), "search_ms": kern_ms_mean, "search_kernel": kernel_name, "mean_path_length": dbar,
                    "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "fallback"}
    line = {
        "metric": "batched MCTS simulations/sec", "value": value, "unit": "simulations/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if total_trees > 0 else "weak", "vs_baseline": None, "dtype": "fp32" if MODEL_DTYPE == "float32" else "bf16",
        "data": "synthetic", "config": workload_config(args, per_gpu=B, total_trees=total_trees),
        "e2e": {"value": e2e_value, "unit": "simulations/s", "h2d_bytes_per_step": int(obs.nbytes + noise.nbytes),
                "d2h_bytes_per_step": int(actions.nbytes + probs.nbytes), "ms_per_step": e2e_ms / K},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "plan": plan, "wall_s": wall, "root_visits_ok": visits_ok,
    }
    if sustained is not None:
        sustained["value"] = sims_per_step / (sustained["ms_per_step"] * 1e-3)
        line["sustained"] = sustained
    del agent, flush, drain
    torch.cuda.empty_cache()
    return line


def measure_train(rank: int, world: int, local_rank: int, dist, K: int, W: int):
    """BASELINE configs[4]: one MuZero train step = K=5 unroll over a replay batch of 1024 per GPU (loss + gradient kernel),
    one NCCL all-reduce of the flat gradient when world > 1, AdamW, parameter push to the search kernels — through
    Trainer.train_step (host batch in, host losses out)."""
    import torch

    from cumind_b200 import Agent, Config
    from cumind_b200.trainer import MemoryBuffer, Trainer

    Ku, At, Dt, Bt, Ht = 5, 2, 4, 1024, 64
    cfg = Config(hidden_dim=Ht, num_blocks=2, action_space_size=At, observation_shape=(Dt,), num_unroll_steps=Ku, batch_size=Bt)
    agent = Agent(cfg)
    trainer = Trainer(agent, MemoryBuffer(8), cfg)
    rng = np.random.default_rng(100 + rank)
    batch = (rng.standard_normal((Bt, Dt)).astype(np.float32), rng.integers(0, At, size=(Bt, Ku)).astype(np.int32),
             {"values": rng.standard_normal(Bt).astype(np.float32), "rewards": rng.standard_normal((Bt, Ku)).astype(np.float32),
              "policies": rng.dirichlet([0.5] * At, size=Bt).astype(np.float32)})
    for _ in range(max(W, 3)):
        trainer.train_step(batch)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    for _ in range(K):
        info = trainer.train_step(batch)
    t1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1e3 / K
    ms = torch.tensor([t0.elapsed_time(t1) / K, wall_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_dev, ms_wall = float(ms[0]), float(ms[1])
    return {"workload": f"BASELINE configs[4]: train step, K={Ku} unroll, replay batch {Bt} per GPU, hidden_dim={Ht}, AdamW, "
                        + ("gradient all-reduce over the GPUs" if world > 1 else "single GPU (no collective)"),
            "metric": "train steps through Trainer.train_step (host batch in, host losses out)", "ms_per_step": ms_wall,
            "device_ms_per_step": ms_dev, "n_gpus": world, "global_batch": Bt * world, "trajectories_per_s": Bt * world / (ms_wall * 1e-3),
            "loss_grad_kernel": "train_fwd_bwd_kernel (hand-written)" if trainer.use_fused_kernel else "torch autograd",
            "allreduce": ("fused into AdamW over NVLink peer memory (p2p_allreduce_adamw_kernel), whole step one CUDA graph" if trainer._p2p is not None
                          else ("NCCL all-reduce of the flat gradient between two CUDA graphs" if world > 1 else "none (1 GPU)")),
            "params": int(trainer._flat.numel()), "loss": info["total_loss"]}


def run_native(args, rank: int, world: int, local_rank: int):
    import torch

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    sampler.start()
    line = measure_search(args, rank, world, local_rank, dist, sampler, args.steps, args.warmup, args.sustain, args.total_trees, args.trees)
    sampler.stop_flag = True
    if line is None:
        return
    line["clocks"] = sampler.summary()
    line["clocks"]["note"] = "NVML, 25 ms period, over the timed steps and the sustained leg"
    # ---- the other BASELINE configurations, measured after the headline (same process, same box) ---------------------
    if args.secondary and WORKLOAD == "cartpole" and args.total_trees == 0:
        secondary = {}
        keep = (OBS_SHAPE, A, H, L, CC, WORKLOAD, MODEL_DTYPE)
        ks = max(5, min(args.steps, 10))

        def brief(d):
            return {k: d[k] for k in ("value", "unit", "ms_per_step", "scaling", "dtype", "config", "e2e", "roofline", "plan", "root_visits_ok")}
        try:
            set_workload("atari")
            secondary["configs[2]"] = brief(measure_search(args, rank, world, local_rank, dist, None, ks, 3, 0.0, 0, 1024))
        finally:
            set_workload_raw(keep)
        if world > 1:
            secondary["configs[3]"] = brief(measure_search(args, rank, world, local_rank, dist, None, ks, 3, 0.0, 65536, 0))
        secondary["configs[4]"] = measure_train(rank, world, local_rank, dist, 50, 5)
        line["secondary"] = secondary
    if rank == 0:
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.sims, args.cpu_seconds)
            line["cpu_baseline_c"] = {"value": c_oracle_rate(args.sims), "unit": "simulations/s", "cores": 1, "kind": "port",
                                      "sample": "oracle/cumind_oracle.c, 512 trees, one thread (pinned fp32 order)"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def phase_clocks(mcts, hidden, d_noise, tree_index_base, S):
    """cumind_search_phase_clocks: SM cycles per phase of the team kernel (tree warp 0 / network warp 0 of
    team 0 in every CTA), averaged over the CTAs and divided by the number of simulations."""
    import ctypes as C

    import torch

    from cumind_b200 import _native
    from cumind_b200.network import _ptr, _stream_ptr

    net = mcts.network
    B = hidden.shape[0]
    A_cfg = int(mcts.config.action_space_size)
    tree = mcts._tree(B, int(mcts.config.num_simulations), min(A_cfg, net.action_space_size))
    cfg = mcts._cfg(1, 0, tree_index_base)
    plan = mcts.plan(B)
    visits = torch.empty((B, A_cfg), dtype=torch.int32, device=net.device)
    probs = torch.empty((B, A_cfg), dtype=torch.float64, device=net.device)
    clk = torch.zeros((plan["grid"], 16), dtype=torch.int64, device=net.device)
    _native.check(mcts._lib.cumind_search_phase_clocks(tree.handle, net._handle, _ptr(hidden), C.byref(cfg), _ptr(d_noise), _ptr(visits),
                                                       _ptr(probs), None, _ptr(clk), _stream_ptr()))
    torch.cuda.synchronize()
    c = clk.cpu().numpy().astype(np.float64)
    names = ["tree.select", "tree.expand_children", "tree.wait_network_backup", "tree.root_and_final", "tree.softmax_priors", "tree.barrier_A",
             "net.gather", "net.layer_barriers", "net.heads", "net.wait_trees", "net.layer_compute", "net.backup_rescore",
             "layerwg.wait_rows_all_teams", "layerwg.layer0_all_teams"]
    cols = [0, 1, 2, 3, 5, 6, 8, 9, 10, 11, 12, 13, 7, 15]
    out = {"plan": plan, "cycles_per_simulation_mean": {n: float(c[:, k].mean()) / S for n, k in zip(names, cols)},
           "cycles_per_simulation_max": {n: float(c[:, k].max()) / S for n, k in zip(names, cols)}}
    out["tree_total_per_sim"] = sum(out["cycles_per_simulation_mean"][n] for n in names[:6])
    return out


def _search_device(mcts, hidden, d_noise, tree_index_base):
    """cumind_search with every buffer already on the device (no host copies)."""
    import ctypes as C

    import torch

    from cumind_b200 import _native
    from cumind_b200.network import _ptr, _stream_ptr

    net = mcts.network
    B = hidden.shape[0]
    A_cfg = int(mcts.config.action_space_size)
    tree = mcts._tree(B, int(mcts.config.num_simulations), min(A_cfg, net.action_space_size))
    cfg = mcts._cfg(1, 0, tree_index_base)
    if not hasattr(mcts, "_bench_out") or mcts._bench_out[0].shape[0] != B:
        mcts._bench_out = (torch.empty((B, A_cfg), dtype=torch.int32, device=net.device),
                           torch.empty((B, A_cfg), dtype=torch.float64, device=net.device))
    visits, probs = mcts._bench_out
    _native.check(mcts._lib.cumind_search(tree.handle, net._handle, _ptr(hidden), C.byref(cfg), _ptr(d_noise), _ptr(visits), _ptr(probs),
                                          None, _stream_ptr()))
    return probs, visits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--trees", type=int, default=4096, help="trees per GPU (weak scaling)")
    ap.add_argument("--total-trees", type=int, default=0,
                    help="global batch sharded across the ranks (strong scaling; BASELINE configs[3] uses 65536)")
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per tree (0 = auto)")
    ap.add_argument("--schedule", default="auto", choices=["auto", "warp", "team", "tensor"], help="cumind_search_cfg.schedule")
    ap.add_argument("--teams", type=int, default=0, help="team schedule: teams per CTA (0 = auto)")
    ap.add_argument("--net-warps", type=int, default=0, help="team schedule: network warps per team (0 = auto)")
    ap.add_argument("--trees-per-team", type=int, default=0, help="team schedule: tree slots per team (0 = auto)")
    ap.add_argument("--phase-clocks", action="store_true", help="print the team kernel's per-phase SM cycles and exit")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the secondary legs (BASELINE configs[2], [3] when N > 1, [4]) measured after the headline")
    ap.add_argument("--sustain", type=float, default=2.5, help="seconds of back-to-back steps after the timed region (0 = skip)")
    ap.add_argument("--workload", default="cartpole", choices=["cartpole", "cartpole-h128", "atari", "atari-literal"],
                    help="cartpole = BASELINE configs[1] (the bench line); atari = configs[2], measured for profiles/ only")
    ap.add_argument("--dtype", default="", choices=["", "float32", "bfloat16"], help="override the workload's model_dtype (bfloat16 = tcgen05 network)")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.dtype:
        global MODEL_DTYPE
        MODEL_DTYPE = args.dtype
    if args.workload != "cartpole":
        args.no_cpu = True
        if args.trees == 4096 and args.workload.startswith("atari"):
            args.trees = 1024
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
