"""TEST INFRASTRUCTURE — run the reference's own data/memory.py, unmodified (it only needs a logger stub).

/root/reference/src/cumind/data/memory.py is loaded from where it lies (never copied) with
`cumind.utils.logger.log` replaced by a no-op.  Used by tests/golden/make_memory_golden.py to produce
tests/golden/memory_cases.npz and by the container-only parity tests of cumind_b200/memory.py; nothing that
runs on the GPU box imports it (the reference is not there).
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CUMIND_REFERENCE_ROOT", "/root/reference")
_PATH = os.path.join(REFERENCE_ROOT, "src", "cumind", "data", "memory.py")
_loaded = None


def available() -> bool:
    return os.path.isfile(_PATH)


class _NoLog:
    def __getattr__(self, name):
        return lambda *a, **k: None


def load():
    """The reference module object (Memory, MemoryBuffer, PrioritizedMemoryBuffer, TreeBuffer)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError(f"reference not present at {_PATH}")
    names = ["cumind", "cumind.data", "cumind.utils", "cumind.utils.logger"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names[:3]:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
        logger = types.ModuleType("cumind.utils.logger")
        logger.log = _NoLog()
        sys.modules["cumind.utils.logger"] = logger
        spec = importlib.util.spec_from_file_location("cumind.data.memory", _PATH)
        mod = importlib.util.module_from_spec(spec)
        mod.__package__ = "cumind.data"
        sys.modules["cumind.data.memory"] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop("cumind.data.memory", None)
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    _loaded = mod
    return mod
