"""Checkpoint ingest: the reference's pickle is read without JAX and yields exactly
the committed blob; this package's own format round-trips."""

import os

import numpy as np
import pytest

from cumind_b200 import checkpoint as ck
from cumind_b200 import params as P
from tests import helpers as hp

REF_CKPT = "/root/reference/checkpoints/CartPole-v1/20250713-141205/episode_01200.pkl"


@pytest.mark.skipif(not os.path.exists(REF_CKPT), reason="/root/reference not present (GPU box)")
def test_reference_pickle_reads_without_jax_and_matches_the_committed_blob():
    state = ck.load_checkpoint(REF_CKPT)
    assert set(state) == {"network_state", "optimizer_state"}
    dims = ck.infer_dims(state["network_state"])
    assert dims == {"hidden_dim": 128, "action_space_size": 2, "num_blocks": 2, "observation_shape": (4,)}
    blob = P.flatten(state["network_state"], (4,), 2, 128, 2, 32)
    assert blob.size == 50948 and np.array_equal(blob, hp.ckpt_params())
    # the optimizer state is not dropped: optax.ScaleByAdamState(count, mu, nu) -> step count + moments in blob layout
    opt = state["optimizer_state"]
    assert opt["count"] > 0
    assert opt["mu"].shape == (50948,) and opt["nu"].shape == (50948,) and opt["mu"].dtype == np.float32
    assert np.any(opt["mu"] != 0) and np.all(opt["nu"] >= 0) and np.any(opt["nu"] > 0)


def test_unpickler_refuses_code(tmp_path):
    """A checkpoint is data: a pickle that asks for a callable outside the allow-list must be refused, in both formats."""
    import pickle

    class Evil:
        def __reduce__(self):
            return (eval, ("__import__('os').getpid()",))

    bad = tmp_path / "evil.pkl"
    bad.write_bytes(pickle.dumps({"network_state": Evil(), "optimizer_state": None}))
    with pytest.raises(pickle.UnpicklingError):
        ck.load_checkpoint(str(bad))
    bad2 = tmp_path / "evil2.pkl"
    bad2.write_bytes(pickle.dumps({"format": "cumind_b200/1", "network_state": Evil()}))
    with pytest.raises(pickle.UnpicklingError):
        ck.load_checkpoint(str(bad2))


def test_own_format_round_trips(tmp_path):
    tree = P.unflatten(hp.ckpt_params(), (4,), 2, 128, 2, 32)
    path = tmp_path / "CartPole-v1" / "x" / "episode_00001.pkl"
    ck.save_checkpoint({"network_state": tree, "optimizer_state": {"count": 3, "mu": None, "nu": None}}, str(path))
    back = ck.load_checkpoint(str(path))
    assert back["optimizer_state"]["count"] == 3
    assert np.array_equal(P.flatten(back["network_state"], (4,), 2, 128, 2, 32), hp.ckpt_params())
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.pkl"
        import pickle
        bad.write_bytes(pickle.dumps({"nope": 1}))
        ck.load_checkpoint(str(bad))
