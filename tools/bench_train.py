#!/usr/bin/env python
"""BASELINE configs[4] (train step): K=5 unroll over a replay batch of 1024, AdamW, gradient all-reduce.
Not the bench line (bench.py measures the search); run for profiles/:
    python tools/bench_train.py [--steps 50]            # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py
Reports ms per train step (CUDA events, max over ranks) and trajectories/s over all ranks."""

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cumind_b200 import Agent, Config  # noqa: E402
from cumind_b200.trainer import MemoryBuffer, Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--hidden", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, A, D, B = 5, 2, 4, args.batch
    cfg = Config(hidden_dim=args.hidden, num_blocks=2, action_space_size=A, observation_shape=(D,), num_unroll_steps=K, batch_size=B)
    agent = Agent(cfg)
    trainer = Trainer(agent, MemoryBuffer(8), cfg)
    rng = np.random.default_rng(rank)
    batch = (rng.standard_normal((B, D)).astype(np.float32), rng.integers(0, A, size=(B, K)).astype(np.int32),
             {"values": rng.standard_normal(B).astype(np.float32), "rewards": rng.standard_normal((B, K)).astype(np.float32),
              "policies": rng.dirichlet([0.5] * A, size=B).astype(np.float32)})
    for _ in range(args.warmup):
        trainer.train_step(batch)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    dev_batch = (torch.as_tensor(batch[0]).cuda(), torch.as_tensor(batch[1]).cuda(), {k: torch.as_tensor(v).cuda() for k, v in batch[2].items()})
    for _ in range(args.steps):
        info = trainer.train_step(batch)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "train steps (K=5 unroll, AdamW, flat-bucket all-reduce)", "ms_per_step": float(ms), "n_gpus": world,
                          "global_batch": B * world, "trajectories_per_s": B * world / (float(ms) * 1e-3), "hidden_dim": args.hidden,
                          "params": int(trainer._flat.numel()), "loss": info["total_loss"]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
