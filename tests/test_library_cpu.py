"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/*.h
declares (no compute without a GPU), the blob layout agrees with the oracle's, and the product
never imports oracle/."""

import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cumind_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cumind_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cumind_b200 import _build, _native

    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    lib = _native.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_native.EXPORTS)
    assert lib.cumind_abi_version() == 3


def test_param_count_and_tree_bytes_without_a_gpu():
    from cumind_b200 import _native, params as P
    from oracle import liboracle as lo

    lib = _native.load()
    for shape, A, H, L, Cc in [((4,), 2, 64, 2, 32), ((8,), 3, 16, 3, 32), ((84, 84, 3), 4, 64, 2, 32), ((3, 84, 84), 4, 64, 2, 32)]:
        d = _native.make_desc(shape, A, H, L, Cc)
        n = int(lib.cumind_param_count(C.byref(d)))
        assert n == P.param_count(shape, A, H, L, Cc) == lo.param_count(lo.make_desc(shape, A, H, L, Cc))
    assert P.param_count((4,), 2, 128, 2, 32) == 50948  # the shipped checkpoint
    bad = _native.make_desc((4, 4), 2, 16, 2, 32)
    assert int(lib.cumind_param_count(C.byref(bad))) == 0
    nbytes = int(lib.cumind_tree_bytes(4096, 50, 2, 64))
    assert nbytes >= 4096 * (103 * 20 + 51 * 64 * 4)
    assert int(lib.cumind_tree_bytes(0, 50, 2, 64)) == 0


def test_blob_layout_matches_oracle_packing():
    from cumind_b200 import params as P
    from oracle import network_np as nn

    for shape, A, H, L, Cc in [((4,), 2, 64, 2, 32), ((10, 10, 4), 4, 16, 2, 8)]:
        tree = nn.random_params(shape, A, H, L, Cc, 5, bias_scale=0.1, bn_random=True)
        blob = P.flatten(tree, shape, A, H, L, Cc)
        assert np.array_equal(blob, nn.flatten_params(tree, shape))
        back = P.unflatten(blob, shape, A, H, L, Cc)
        assert np.array_equal(P.flatten(back, shape, A, H, L, Cc), blob)


def test_checkpoint_blob_loads_into_the_parameter_tree():
    from cumind_b200 import params as P
    from tests import helpers as hp

    blob = hp.ckpt_params()
    tree = P.unflatten(blob, (4,), 2, 128, 2, 32)
    assert tree["dynamics_network"]["action_embedding"]["embedding"].shape == (2, 128)
    assert tree["prediction_network"]["policy_head"]["kernel"].shape == (128, 2)
    assert tree["representation_network"]["encoder"]["layers"][0]["kernel"].shape == (4, 128)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cumind_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert not re.search(r'#include\s*[<"][^>"]*oracle', text), f
                assert "libcumind_oracle" not in text and "dlopen" not in text, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cumind_b200 import _build, _native

    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_build, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.delenv("CUMIND_B200_AUTOBUILD", raising=False)
    with pytest.raises(_native.CuMindNativeError, match="no CPU fallback"):
        _native.load()


def test_jax_ffi_shim_is_guarded():
    """cumind_b200/ffi: the XLA FFI handlers compile to an empty unit without jaxlib's headers, the Python front end
    raises ImportError (with the reason) without jax, and nothing in the package imports it."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "cumind_b200", "ffi", "cumind_ffi.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(root, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    for sym in ("cumind_initial_inference", "cumind_recurrent_inference", "cumind_search"):
        assert sym in text
    try:
        import jax  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="jax"):
            import cumind_b200.ffi.jax_ffi  # noqa: F401
    for name in os.listdir(os.path.join(root, "cumind_b200")):
        if name.endswith(".py"):
            assert "jax_ffi" not in open(os.path.join(root, "cumind_b200", name)).read(), name
