"""Builds cumind_b200/_lib/libcumind_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, so the
built library travels to the GPU box with the repo snapshot)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libcumind_b200.so")
SOURCES = ["abi.cu", "search_fused.cu", "search_team.cu", "tree_steps.cu", "network_fp32.cu", "conv_fp32.cu", "network_tc.cu", "search_tc.cu", "conv_tc.cu", "train.cu", "train_fused.cu", "p2p_adamw.cu", "replay.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcumind_b200.so")


def sources() -> List[str]:
    return [os.path.join(CSRC, s) for s in SOURCES]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "cumind_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _stale_obj(src: str, obj: str, headers: List[str]) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + headers)


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per .cu (compiled in parallel, only the stale ones), then one link step."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "cumind_b200.h"))
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src: str):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        if not force and not _stale_obj(src, obj, headers):
            return obj, ""
        r = subprocess.run([_nvcc()] + cflags + ["-c", "-o", obj, src], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    tmp = LIB_PATH + ".tmp"
    r = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-o", tmp] + [o for o, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
