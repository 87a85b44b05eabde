"""TEST INFRASTRUCTURE — float64 NumPy restatement of the reference train step for vector observations.

Loss and gradient of src/cumind/agent/trainer.py:172-205 (value MSE, policy cross-entropy, reward MSE over
a K-step unroll, same value/policy target at every step, losses divided by K+1 / K) with a hand-written
backward pass (checked against finite differences in tests/test_train_cpu.py), and optax.adamw
(agent.py:48: b1=0.9, b2=0.999, eps=1e-8, decoupled weight decay).  Parameter tree layout = oracle/network_np.py.
"""

from __future__ import annotations

import copy
from typing import Any, Dict, Tuple

import numpy as np


def _log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def loss_and_grad(params: Dict[str, Any], obs, actions, targets, K: int) -> Tuple[float, Dict[str, float], Dict[str, Any]]:
    p = params
    enc = p["representation_network"]["encoder"]["layers"]
    dyn = p["dynamics_network"]
    pred = p["prediction_network"]
    L = len(enc)
    B = obs.shape[0]
    f = lambda a: np.asarray(a, dtype=np.float64)
    tv, tp, tr = f(targets["values"]), f(targets["policies"]), f(targets["rewards"])
    g = copy.deepcopy(p)

    def zero(node):
        for k, v in node.items():
            if isinstance(v, dict):
                zero(v)
            else:
                node[k] = np.zeros_like(f(v))
    zero(g)

    # ---- forward, keeping what the backward pass needs --------------------------------------------
    enc_in, x = [], f(obs)
    for i in range(L):
        enc_in.append(x)
        z = x @ f(enc[i]["kernel"]) + f(enc[i]["bias"])
        x = np.maximum(z, 0) if i < L - 1 else z
    states = [x]
    dyn_cache = []
    for s in range(K):
        a = np.asarray(actions[:, s], dtype=np.int64)
        xin = states[-1] + f(dyn["action_embedding"]["embedding"])[a]
        layer_in, layer_z, cur = [], [], xin
        for l in range(L):
            layer_in.append(cur)
            z = cur @ f(dyn["layers"][l]["kernel"]) + f(dyn["layers"][l]["bias"])
            layer_z.append(z)
            cur = np.maximum(z, 0) + cur
        dyn_cache.append((a, layer_in, layer_z))
        states.append(cur)
    value_loss = policy_loss = reward_loss = 0.0
    d_states = [np.zeros_like(s) for s in states]
    Wp, bp = f(pred["policy_head"]["kernel"]), f(pred["policy_head"]["bias"])
    Wv, bv = f(pred["value_head"]["kernel"]), f(pred["value_head"]["bias"])
    Wr, br = f(dyn["reward_head"]["kernel"]), f(dyn["reward_head"]["bias"])
    nv = K + 1 if K > 0 else 1
    nr = K if K > 0 else 1
    for s, h in enumerate(states):
        logits = h @ Wp + bp
        value = (h @ Wv + bv)[:, 0]
        ls = _log_softmax(logits)
        value_loss += np.mean((value - tv) ** 2) / nv
        policy_loss += -np.mean(np.sum(tp * ls, axis=-1)) / nv
        dvalue = 2.0 * (value - tv) / (B * nv)
        dlogits = (np.exp(ls) * tp.sum(axis=-1, keepdims=True) - tp) / (B * nv)
        g["prediction_network"]["policy_head"]["kernel"] += h.T @ dlogits
        g["prediction_network"]["policy_head"]["bias"] += dlogits.sum(axis=0)
        g["prediction_network"]["value_head"]["kernel"] += h.T @ dvalue[:, None]
        g["prediction_network"]["value_head"]["bias"] += dvalue.sum(keepdims=True)
        d_states[s] += dlogits @ Wp.T + dvalue[:, None] @ Wv.T
        if s >= 1:
            reward = (h @ Wr + br)[:, 0]
            reward_loss += np.mean((reward - tr[:, s - 1]) ** 2) / nr
            dr = 2.0 * (reward - tr[:, s - 1]) / (B * nr)
            g["dynamics_network"]["reward_head"]["kernel"] += h.T @ dr[:, None]
            g["dynamics_network"]["reward_head"]["bias"] += dr.sum(keepdims=True)
            d_states[s] += dr[:, None] @ Wr.T
    # ---- backward through the unroll ----------------------------------------------------------------
    for s in range(K - 1, -1, -1):
        a, layer_in, layer_z = dyn_cache[s]
        d = d_states[s + 1]
        for l in range(L - 1, -1, -1):
            dz = d * (layer_z[l] > 0)
            g["dynamics_network"]["layers"][l]["kernel"] += layer_in[l].T @ dz
            g["dynamics_network"]["layers"][l]["bias"] += dz.sum(axis=0)
            d = d + dz @ f(dyn["layers"][l]["kernel"]).T
        np.add.at(g["dynamics_network"]["action_embedding"]["embedding"], a, d)
        d_states[s] += d
    d = d_states[0]
    for i in range(L - 1, -1, -1):
        z_in = enc_in[i]
        if i < L - 1:
            z = z_in @ f(enc[i]["kernel"]) + f(enc[i]["bias"])
            d = d * (z > 0)
        g["representation_network"]["encoder"]["layers"][i]["kernel"] += z_in.T @ d
        g["representation_network"]["encoder"]["layers"][i]["bias"] += d.sum(axis=0)
        d = d @ f(enc[i]["kernel"]).T
    losses = {"value_loss": value_loss, "policy_loss": policy_loss, "reward_loss": reward_loss}
    return value_loss + policy_loss + reward_loss, losses, g


def adamw(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, lr: float, wd: float,
          b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mh, vh = m / (1 - b1 ** step), v / (1 - b2 ** step)
    return p - lr * (mh / (np.sqrt(vh) + eps) + wd * p), m, v
