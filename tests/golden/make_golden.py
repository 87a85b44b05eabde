"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

search_cases.npz   trees produced by the reference's own mcts.py executed verbatim
                   (oracle/reference_exec.py) over the NumPy network of oracle/network_np.py, for
                   A in {1,2,3,4,18}, S in {1,2,3,5,25,50,100}, with and without injected Dirichlet
                   noise, plus the action_space_size edge cases of tests/test_mcts.py:271-279 and
                   the shipped CartPole checkpoint with its configuration.json search settings.
network_cases.npz  known-answer vectors of the NumPy network restatement (the reference's own tests
                   pin shapes only — network numerics are "parity unpinned" against flax).
cartpole_ckpt.npz  the fp32 arrays of checkpoints/CartPole-v1/20250713-141205/episode_01200.pkl,
                   read with a stub unpickler (no JAX needed), as a flat blob + config.
"""

from __future__ import annotations

import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import network_np as nn  # noqa: E402
from oracle import reference_exec as rx  # noqa: E402
from oracle import tree_canon as tc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = "/root/reference/checkpoints/CartPole-v1/20250713-141205/episode_01200.pkl"
CKPT_CFG = "/root/reference/configuration.json"


# ---- checkpoint reader (no JAX): arrays are pickled through numpy's _reconstruct inside
# jax._src.array._reconstruct_array; every other class is replaced by a stub ------------------
class _Stub:
    def __init__(self, *a, **k):
        self.args, self.kw = a, k

    def __setstate__(self, st):
        self.state = st


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy"):
            return super().find_class(module, name)
        if module + "." + name == "jax._src.array._reconstruct_array":
            def rebuild(fun, args, arr_state, aval_state):
                arr = fun(*args)
                arr.__setstate__(arr_state)
                return arr
            return rebuild
        return type(name, (_Stub,), {"__module__": module})


def read_checkpoint_params(path):
    with open(path, "rb") as f:
        obj = _Unpickler(f).load()
    mapping = obj["network_state"].state["_mapping"]

    def conv(node):
        if isinstance(node, dict):
            return {k: conv(v) for k, v in node.items() if k != "rngs"}
        if isinstance(node, _Stub):  # VariableState: state = (None, {"type":..., "value": ndarray})
            return np.asarray(node.state[1]["value"], dtype=np.float32)
        raise TypeError(type(node))

    return conv(mapping)


def tree_record(prefix, canon, out):
    for k, v in canon.items():
        out[f"{prefix}_{k}"] = v


def main():
    assert rx.available(), "needs /root/reference"
    out, manifest = {}, []
    nets = {}

    def get_net(A_net, H, L, seed):
        key = f"net_A{A_net}_H{H}_L{L}_s{seed}"
        if key not in nets:
            params = nn.random_params((4,), A_net, H, L, 32, seed, bias_scale=0.05)
            nets[key] = params
            out[key] = nn.flatten_params(params, (4,))
        return key, nets[key]

    idx = 0
    combos = [(A, A) for A in (1, 2, 3, 4, 18)] + [(2, 1), (4, 6), (4, 2)]
    for A_net, A_cfg in combos:
        H = 64 if A_net == 2 else 16
        for S in (1, 2, 3, 5, 25, 50, 100):
            for noise_on in (False, True):
                seed = 100 + idx
                key, params = get_net(A_net, H, 2, 7)
                net = nn.NumpyCuMindNetwork(params, (4,))
                rng = np.random.default_rng(seed)
                obs = rng.uniform(-0.05, 0.05, size=(1, 4)).astype(np.float32) if idx % 2 else np.ones((1, 4), np.float32)
                root_h = net.representation_network(obs)[0]
                A = min(A_net, A_cfg)
                alpha = 0.3
                noise = rng.dirichlet([alpha] * A) if noise_on else None
                c_puct = 1.25 if idx % 3 else 1.0
                cfg = rx.SearchConfig(S, A_cfg, c_puct, alpha, 0.25)
                probs, root = rx.search_with_tree(net, cfg, root_h, noise)
                pre = f"c{idx}"
                tree_record(pre, tc.canon_from_nodes(root), out)
                out[f"{pre}_probs"] = np.asarray(probs, dtype=np.float64)
                out[f"{pre}_root_h"] = root_h
                out[f"{pre}_obs"] = obs[0]
                if noise is not None:
                    out[f"{pre}_noise"] = noise
                manifest.append({"id": pre, "net": key, "obs_shape": [4], "A_net": A_net, "A_cfg": A_cfg, "H": H, "L": 2, "S": S,
                                 "c_puct": c_puct, "frac": 0.25, "alpha": alpha, "noise": bool(noise_on)})
                idx += 1

    # shipped checkpoint, configuration.json search settings (S=25, c_puct=1.25, alpha=.25, f=.25, H=128)
    params = read_checkpoint_params(CKPT)
    with open(CKPT_CFG) as f:
        cj = json.load(f)
    blob = nn.flatten_params(params, (4,))
    assert blob.size == 50948, blob.size
    np.savez_compressed(os.path.join(HERE, "cartpole_ckpt.npz"), params=blob, config=json.dumps(cj))
    net = nn.NumpyCuMindNetwork(params, (4,))
    rng = np.random.default_rng(2025)
    m = cj["MCTS"]
    for j in range(8):
        obs = rng.uniform(-0.05, 0.05, size=(1, 4)).astype(np.float32)
        if j >= 4:
            obs = (obs * np.float32(20.0)).astype(np.float32)  # states away from the reset range
        root_h = net.representation_network(obs)[0]
        noise = rng.dirichlet([m["dirichlet_alpha"]] * 2) if j % 2 else None
        cfg = rx.SearchConfig(m["num_simulations"], 2, m["c_puct"], m["dirichlet_alpha"], m["exploration_fraction"])
        probs, root = rx.search_with_tree(net, cfg, root_h, noise)
        pre = f"c{idx}"
        tree_record(pre, tc.canon_from_nodes(root), out)
        out[f"{pre}_probs"] = np.asarray(probs, dtype=np.float64)
        out[f"{pre}_root_h"] = root_h
        out[f"{pre}_obs"] = obs[0]
        if noise is not None:
            out[f"{pre}_noise"] = noise
        manifest.append({"id": pre, "net": "ckpt", "obs_shape": [4], "A_net": 2, "A_cfg": 2, "H": 128, "L": 2, "S": m["num_simulations"],
                         "c_puct": m["c_puct"], "frac": m["exploration_fraction"], "alpha": m["dirichlet_alpha"], "noise": noise is not None})
        idx += 1
    out["manifest"] = np.array(json.dumps(manifest))
    np.savez_compressed(os.path.join(HERE, "search_cases.npz"), **out)
    print("search cases:", idx)

    # ---- network known-answer vectors (NumPy restatement) --------------------------------------
    nout, nman = {}, []
    specs = [((4,), 2, 64, 2, 32, False), ((4,), 2, 128, 2, 32, False), ((8,), 3, 16, 3, 32, False), ((4,), 18, 32, 1, 32, False),
             ((10, 10, 4), 4, 16, 2, 8, True), ((12, 9, 3), 4, 32, 1, 8, True), ((3, 20, 20), 4, 16, 2, 16, False)]
    for i, (shape, A, H, L, Cc, bnr) in enumerate(specs):
        params = nn.random_params(shape, A, H, L, Cc, 900 + i, bias_scale=0.1, bn_random=bnr)
        net = nn.NumpyCuMindNetwork(params, shape)
        rng = np.random.default_rng(i)
        B = 5
        obs = rng.standard_normal((B,) + tuple(shape)).astype(np.float32)
        h, lg, v = net.initial_inference(obs)
        act = rng.integers(0, A, size=B).astype(np.int32)
        nx, rw, lg2, v2 = net.recurrent_inference(h, act)
        pre = f"n{i}"
        nout.update({f"{pre}_params": nn.flatten_params(params, shape), f"{pre}_obs": obs, f"{pre}_h": h, f"{pre}_logits": lg,
                     f"{pre}_value": v[:, 0], f"{pre}_action": act, f"{pre}_next": nx, f"{pre}_reward": rw[:, 0],
                     f"{pre}_logits2": lg2, f"{pre}_value2": v2[:, 0], f"{pre}_probs2": nn.softmax(lg2)})
        nman.append({"id": pre, "obs_shape": list(shape), "A": A, "H": H, "L": L, "Cc": Cc})
    nout["manifest"] = np.array(json.dumps(nman))
    np.savez_compressed(os.path.join(HERE, "network_cases.npz"), **nout)
    print("network cases:", len(specs))


if __name__ == "__main__":
    main()
