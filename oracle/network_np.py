"""TEST INFRASTRUCTURE — NumPy restatement of CuMind's networks (forward only).

Restates /root/reference/src/cumind/core/network.py with the canonical numerics of DESIGN.md:
  ResidualBlock.__call__      network.py:29-45
  VectorEncoder.__call__      network.py:94-108
  ConvEncoder.__call__        network.py:131-147
  DynamicsNetwork.__call__    network.py:217-236
  PredictionNetwork.__call__  network.py:254-266
  initial/recurrent_inference network.py:288-315
flax.nnx / jax are absent from this image (flax 0.10.6, jax 0.6.2 pinned in uv.lock); their layer
semantics are restated from their published behaviour (Linear: x@kernel[in,out]+bias; Embed: row
lookup; Conv: NHWC cross-correlation, symmetric padding; BatchNorm inference; softmax).

Canonical dot product = sequential chain of IEEE fp32 fused multiply-adds (fma32 below emulates
fmaf exactly in NumPy: exact float64 product, TwoSum error term, round-to-odd, one rounding to
float32); everything else is a single float32 op, which NumPy never fuses.  The loops below are
therefore bit-identical to oracle/cumind_oracle.c (checked by tests/test_oracle_golden.py) and to
the CUDA fp32-exact kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; cumind_b200/ never does.
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

_LOG2E = np.float64(1.4426950408889634)
_LN2_HI = np.float64(6.93147180369123816490e-01)
_LN2_LO = np.float64(1.90821492927058770002e-10)
_EXP_C = [np.float64(1.0), np.float64(1.0), np.float64(0.5)] + [
    np.float64(1.0) / np.float64(float(__import__("math").factorial(i))) for i in range(3, 14)
]


def pexp(d: np.ndarray) -> np.ndarray:
    """Pinned float64 exp for d <= 0 (same operation sequence as cmo_pexp in cumind_oracle.c)."""
    d = np.asarray(d, dtype=np.float64)
    out = np.zeros_like(d)
    nan = np.isnan(d)
    live = d >= -200.0
    dd = np.where(live, np.minimum(d, 0.0), 0.0)
    kf = np.rint(dd * _LOG2E)
    r = (dd - kf * _LN2_HI) - kf * _LN2_LO
    p = np.full_like(dd, _EXP_C[13])
    for i in range(12, -1, -1):
        p = p * r + _EXP_C[i]
    bits = (kf.astype(np.int64) + 1023).astype(np.uint64) << np.uint64(52)
    scale = bits.view(np.float64)
    out = np.where(live, p * scale, 0.0)
    out = np.where(nan, d, out)
    return out


def softmax(logits: np.ndarray) -> np.ndarray:
    """jax.nn.softmax(x, axis=-1) in fp32 with the pinned exp (mcts.py:128,191)."""
    x = np.asarray(logits, dtype=f32)
    m = np.max(x, axis=-1, keepdims=True)
    e = pexp((x - m).astype(np.float64)).astype(f32)
    s = np.zeros(e.shape[:-1], dtype=f32)
    for a in range(e.shape[-1]):
        s = s + e[..., a]
    return e / s[..., None]


def fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Exact IEEE-754 binary32 fused multiply-add fl32(a*b + c) for float32 arrays.

    a*b is exact in float64 (24+24 significand bits).  The float64 sum s = p + c may round; the
    TwoSum error term tells whether and in which direction, s is turned into the round-to-odd
    result (truncate toward zero, OR a sticky bit into the last place) and a single final rounding
    to float32 is then exact because float64 carries more than two extra bits.  Checked against
    fmaf() in tests/test_oracle_golden.py, including true double-rounding cases."""
    a64, b64, c64 = (np.asarray(v, dtype=f32).astype(np.float64) for v in (a, b, c))
    with np.errstate(over="ignore", invalid="ignore"):
        p = a64 * b64
        s = np.asarray(p + c64)
        bb = s - p
        err = (p - (s - bb)) + (c64 - bb)
    bits = s.view(np.uint64)
    inexact = np.isfinite(s) & np.isfinite(err) & (err != 0)
    toward_zero = inexact & ((err < 0) != (s < 0)) & (s != 0)
    bits = np.where(toward_zero, bits - np.uint64(1), bits)
    bits = np.where(inexact, bits | np.uint64(1), bits)
    with np.errstate(over="ignore"):
        return bits.view(np.float64).astype(f32)


def linear(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """nnx.Linear: sequential-k chain of fp32 FMAs, then + bias. x [B,in], kernel [in,out], bias [out]."""
    x = np.asarray(x, dtype=f32)
    kernel = np.asarray(kernel, dtype=f32)
    acc = np.zeros((x.shape[0], kernel.shape[1]), dtype=f32)
    for k in range(kernel.shape[0]):
        acc = fma32(x[:, k : k + 1], kernel[k][None, :], acc)
    return acc + np.asarray(bias, dtype=f32)[None, :]


def relu(x: np.ndarray) -> np.ndarray:
    return np.where(x > 0, x, f32(0.0)).astype(f32)


def conv3x3(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride: int) -> np.ndarray:
    """nnx.Conv(kernel_size=(3,3), padding=1, strides=stride), NHWC, taps in the padding skipped."""
    x = np.asarray(x, dtype=f32)
    B, Hh, Ww, Cin = x.shape
    Cout = kernel.shape[3]
    Ho = (Hh - 1) // stride + 1
    Wo = (Ww - 1) // stride + 1
    acc = np.zeros((B, Ho, Wo, Cout), dtype=f32)
    oy = np.arange(Ho)
    ox = np.arange(Wo)
    for ky in range(3):
        iy = oy * stride - 1 + ky
        vy = (iy >= 0) & (iy < Hh)
        for kx in range(3):
            ix = ox * stride - 1 + kx
            vx = (ix >= 0) & (ix < Ww)
            if not vy.any() or not vx.any():
                continue
            ys, xs = oy[vy], ox[vx]
            patch = x[:, iy[vy]][:, :, ix[vx]]  # [B, ny, nx, Cin]
            sub = acc[:, ys[0] : ys[-1] + 1, xs[0] : xs[-1] + 1]  # valid output positions are contiguous
            for ci in range(Cin):
                sub = fma32(patch[..., ci : ci + 1], kernel[ky, kx, ci][None, None, None, :], sub)
            acc[:, ys[0] : ys[-1] + 1, xs[0] : xs[-1] + 1] = sub
    return acc + np.asarray(bias, dtype=f32)


def batchnorm(x: np.ndarray, bn: Dict[str, np.ndarray]) -> np.ndarray:
    """nnx.BatchNorm(use_running_average=True), eps=1e-5 (network.py:25,27)."""
    scale, beta, mean, var = (np.asarray(bn[k], dtype=f32) for k in ("scale", "bias", "mean", "var"))
    g = scale * (f32(1.0) / np.sqrt(var + f32(1e-5)))
    return ((x - mean) * g) + beta


class _Callable:
    def __init__(self, fn):
        self._fn = fn

    def __call__(self, *a):
        return self._fn(*a)


class NumpyCuMindNetwork:
    """Same attribute surface as CuMindNetwork (network.py:269-315), NumPy arrays in and out."""

    def __init__(self, params: Dict[str, Any], observation_shape: Sequence[int], blas: bool = False):
        """blas=True swaps the pinned sequential-k dot products for `x @ kernel` (BLAS order, what the
        reference's XLA dot would cost on a CPU): used ONLY to time the CPU baseline, never for parity."""
        self.params = params
        self.observation_shape = tuple(int(s) for s in observation_shape)
        self._linear = (lambda x, k, b: (np.asarray(x, f32) @ np.asarray(k, f32) + np.asarray(b, f32)).astype(f32)) if blas else linear
        self.representation_network = _Callable(self._representation)
        self.dynamics_network = _Callable(self._dynamics)
        self.prediction_network = _Callable(self._prediction)

    # --- sub-networks -------------------------------------------------------------------------
    def _representation(self, obs: np.ndarray) -> np.ndarray:
        enc = self.params["representation_network"]["encoder"]
        x = np.asarray(obs, dtype=f32)
        if len(self.observation_shape) == 1:
            layers = enc["layers"]
            n = len(layers)
            for i in range(n):
                x = self._linear(x, layers[i]["kernel"], layers[i]["bias"])
                if i < n - 1:
                    x = relu(x)
            return x
        if len(self.observation_shape) != 3:
            raise ValueError(f"Unsupported observation shape: {self.observation_shape}")
        x = relu(conv3x3(x, enc["initial_conv"]["kernel"], enc["initial_conv"]["bias"], 2))
        blocks = enc["residual_blocks"]
        for i in range(len(blocks)):
            blk = blocks[i]
            t = conv3x3(x, blk["conv1"]["kernel"], blk["conv1"]["bias"], 1)
            t = relu(batchnorm(t, blk["bn1"]))
            t = conv3x3(t, blk["conv2"]["kernel"], blk["conv2"]["bias"], 1)
            t = batchnorm(t, blk["bn2"])
            x = relu(t + x)
        B, Ho, Wo, Cc = x.shape
        flat = x.reshape(B, Ho * Wo, Cc)
        s = np.zeros((B, Cc), dtype=f32)
        for i in range(Ho * Wo):
            s = s + flat[:, i]
        pooled = s / f32(Ho * Wo)
        return self._linear(pooled, enc["final_dense"]["kernel"], enc["final_dense"]["bias"])

    def _dynamics(self, state: np.ndarray, action: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        dyn = self.params["dynamics_network"]
        emb = np.asarray(dyn["action_embedding"]["embedding"], dtype=f32)
        a = np.asarray(action, dtype=np.int32)
        x = np.asarray(state, dtype=f32) + emb[a]
        layers = dyn["layers"]
        for i in range(len(layers)):
            x = relu(self._linear(x, layers[i]["kernel"], layers[i]["bias"])) + x
        reward = self._linear(x, dyn["reward_head"]["kernel"], dyn["reward_head"]["bias"])
        return x, reward

    def _prediction(self, hidden: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        pred = self.params["prediction_network"]
        h = np.asarray(hidden, dtype=f32)
        logits = self._linear(h, pred["policy_head"]["kernel"], pred["policy_head"]["bias"])
        value = self._linear(h, pred["value_head"]["kernel"], pred["value_head"]["bias"])
        return logits, value

    # --- CuMindNetwork API ----------------------------------------------------------------------
    def initial_inference(self, observation: np.ndarray):
        h = self._representation(observation)
        logits, value = self._prediction(h)
        return h, logits, value

    def recurrent_inference(self, hidden_state: np.ndarray, action: np.ndarray):
        nxt, reward = self._dynamics(hidden_state, action)
        logits, value = self._prediction(nxt)
        return nxt, reward, logits, value


# ------------------------------------------------------------------------------------------------
# synthetic parameters (the reference initialises with flax lecun_normal / zeros through a JAX
# threefry key, which cannot be reproduced without JAX; shapes and scale are the same)
# ------------------------------------------------------------------------------------------------
def random_params(observation_shape: Sequence[int], action_space_size: int, hidden_dim: int, num_blocks: int,
                  conv_channels: int, seed: int, bias_scale: float = 0.0, bn_random: bool = False) -> Dict[str, Any]:
    rng = np.random.default_rng(seed)
    H, L, A, Cc = hidden_dim, num_blocks, action_space_size, conv_channels

    def dense(i, o):
        k = (rng.standard_normal((i, o)) / np.sqrt(i)).astype(f32)
        b = (rng.standard_normal(o) * bias_scale).astype(f32)
        return {"kernel": k, "bias": b}

    def conv(i, o):
        k = (rng.standard_normal((3, 3, i, o)) / np.sqrt(9 * i)).astype(f32)
        b = (rng.standard_normal(o) * bias_scale).astype(f32)
        return {"kernel": k, "bias": b}

    def bn(c):
        if bn_random:
            return {"scale": (1.0 + 0.1 * rng.standard_normal(c)).astype(f32), "bias": (0.1 * rng.standard_normal(c)).astype(f32),
                    "mean": (0.1 * rng.standard_normal(c)).astype(f32), "var": (1.0 + 0.2 * rng.random(c)).astype(f32)}
        return {"scale": np.ones(c, f32), "bias": np.zeros(c, f32), "mean": np.zeros(c, f32), "var": np.ones(c, f32)}

    shape = tuple(observation_shape)
    if len(shape) == 1:
        layers, cur = {}, shape[0]
        for i in range(L):
            layers[i] = dense(cur, H)
            cur = H
        enc = {"layers": layers}
    elif len(shape) == 3:
        enc = {"initial_conv": conv(shape[2], Cc),
               "residual_blocks": {i: {"conv1": conv(Cc, Cc), "bn1": bn(Cc), "conv2": conv(Cc, Cc), "bn2": bn(Cc)} for i in range(L)},
               "final_dense": dense(Cc, H)}
    else:
        raise ValueError(f"Unsupported observation shape: {shape}")
    dyn = {"action_embedding": {"embedding": (rng.standard_normal((A, H)) / np.sqrt(H)).astype(f32)},
           "layers": {i: dense(H, H) for i in range(L)},
           "reward_head": dense(H, 1)}
    pred = {"policy_head": dense(H, A), "value_head": dense(H, 1)}
    return {"representation_network": {"encoder": enc}, "dynamics_network": dyn, "prediction_network": pred}


def flatten_params(params: Dict[str, Any], observation_shape: Sequence[int]) -> np.ndarray:
    """Parameter tree -> flat fp32 blob in the order documented in include/cumind_b200.h."""
    out = []
    enc = params["representation_network"]["encoder"]
    if len(observation_shape) == 1:
        for i in range(len(enc["layers"])):
            out += [enc["layers"][i]["kernel"], enc["layers"][i]["bias"]]
    else:
        out += [enc["initial_conv"]["kernel"], enc["initial_conv"]["bias"]]
        for i in range(len(enc["residual_blocks"])):
            blk = enc["residual_blocks"][i]
            for c, b in (("conv1", "bn1"), ("conv2", "bn2")):
                out += [blk[c]["kernel"], blk[c]["bias"], blk[b]["scale"], blk[b]["bias"], blk[b]["mean"], blk[b]["var"]]
        out += [enc["final_dense"]["kernel"], enc["final_dense"]["bias"]]
    dyn = params["dynamics_network"]
    out.append(dyn["action_embedding"]["embedding"])
    for i in range(len(dyn["layers"])):
        out += [dyn["layers"][i]["kernel"], dyn["layers"][i]["bias"]]
    out += [dyn["reward_head"]["kernel"], dyn["reward_head"]["bias"]]
    pred = params["prediction_network"]
    out += [pred["policy_head"]["kernel"], pred["policy_head"]["bias"], pred["value_head"]["kernel"], pred["value_head"]["bias"]]
    return np.concatenate([np.asarray(a, dtype=f32).reshape(-1) for a in out])
