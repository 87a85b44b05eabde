"""Hyper-parameter container with the reference's field names, defaults and JSON layout.

Mirrors cumind.config.Config (src/cumind/config.py:13-186): a mutable dataclass (the reference's
tests assign fields after construction, tests/test_mcts.py:184-188), `from_json` / `to_json` with
the sectioned layout of configuration.json, and `validate()` raising ValueError with the same
messages.  chex is not needed: a plain dataclass gives the same behaviour.  Only the MCTS / network
/ environment fields are consumed by this package; the rest ride along so reference JSON files load.
"""

from __future__ import annotations

import dataclasses
import json
from pathlib import Path
from typing import Any, Dict, Tuple

# section name -> ((field, default), ...); order is the reference's to_json order (config.py:115-125)
_SECTIONS: Dict[str, Tuple[Tuple[str, Any], ...]] = {
    "Network architecture": (("hidden_dim", 128), ("num_blocks", 2), ("conv_channels", 32)),
    "Training": (("batch_size", 32), ("learning_rate", 1e-3), ("weight_decay", 1e-4), ("target_update_frequency", 10000000),
                 ("checkpoint_interval", 100), ("num_episodes", 500), ("train_frequency", 5), ("checkpoint_root_dir", "checkpoints")),
    "MCTS": (("num_simulations", 10), ("c_puct", 1.0), ("dirichlet_alpha", 0.3), ("exploration_fraction", 0.25)),
    "Environment": (("env_name", "CartPole-v1"), ("action_space_size", 2), ("observation_shape", (4,))),
    "Self-Play": (("num_unroll_steps", 3), ("td_steps", 5), ("discount", 0.997)),
    "Memory": (("memory_capacity", 1000), ("min_memory_size", 3), ("min_memory_pct", 0.0), ("per_alpha", 0.6),
               ("per_epsilon", 1e-6), ("per_beta", 0.4)),
    "Data Types": (("model_dtype", "float32"), ("action_dtype", "int32"), ("target_dtype", "float32")),
    "Device": (("device_type", "cpu"),),
    "Other": (("seed", 42),),
}

_POSITIVE = ("hidden_dim", "action_space_size", "learning_rate", "batch_size", "num_simulations")
_CHOICES = {
    "model_dtype": ["float32", "float16", "bfloat16"],
    "action_dtype": ["int32", "int64"],
    "target_dtype": ["float32", "float16", "bfloat16"],
    "device_type": ["cpu", "gpu", "tpu"],
}


def _make_config_class():
    fields = []
    for section in _SECTIONS.values():
        for name, default in section:
            typ = type(default) if not isinstance(default, tuple) else Tuple[int, ...]
            fields.append((name, typ, dataclasses.field(default=default)))
    def _post_init(self):
        if isinstance(self.observation_shape, list):  # JSON has no tuples
            self.observation_shape = tuple(self.observation_shape)

    ns = {"__doc__": "CuMind hyper-parameters (config.py:21-73).", "__post_init__": _post_init}
    return dataclasses.make_dataclass("_ConfigFields", fields, namespace=ns, eq=False)


_Base = _make_config_class()


class Config(_Base):  # type: ignore[misc, valid-type]
    """Keyword-constructed, mutable configuration.  `device_type` keeps the reference's vocabulary
    ("cpu" | "gpu" | "tpu"); this package always computes on a CUDA device."""

    def __eq__(self, other):
        return isinstance(other, Config) and dataclasses.asdict(self) == dataclasses.asdict(other)

    __hash__ = None  # type: ignore[assignment]

    @classmethod
    def from_json(cls, json_path: str) -> "Config":
        """Sectioned or flat JSON -> Config (config.py:75-102)."""
        path = Path(json_path)
        if not path.exists():
            raise FileNotFoundError(f"Config file not found: {json_path}")
        with open(path, "r", encoding="utf-8") as f:
            raw = json.load(f)
        flat: Dict[str, Any] = {}
        for key, val in raw.items():
            if isinstance(val, dict):
                flat.update(val)
            else:
                flat[key] = val
        return cls(**flat)

    def to_json(self, json_path: str) -> None:
        """Config -> sectioned JSON, same section names and order as the reference (config.py:104-135)."""
        path = Path(json_path)
        path.parent.mkdir(parents=True, exist_ok=True)
        values = dataclasses.asdict(self)
        doc = {section: {name: values[name] for name, _ in members} for section, members in _SECTIONS.items()}
        with open(path, "w", encoding="utf-8") as f:
            json.dump(doc, f, indent=2)

    def validate(self) -> None:
        """Raises ValueError for the same conditions, with the same messages, as config.py:137-186."""
        for name in ("hidden_dim", "action_space_size"):
            if getattr(self, name) <= 0:
                raise ValueError(f"{name} must be positive, got {getattr(self, name)}")
        if not self.observation_shape:
            raise ValueError("observation_shape cannot be empty")
        if len(self.observation_shape) not in (1, 3):
            raise ValueError(f"observation_shape must be 1D or 3D, got {len(self.observation_shape)}D")
        for name in ("learning_rate", "batch_size", "num_simulations"):
            if getattr(self, name) <= 0:
                raise ValueError(f"{name} must be positive, got {getattr(self, name)}")
        for name, allowed in _CHOICES.items():
            if getattr(self, name) not in allowed:
                raise ValueError(f"{name} must be one of {allowed}, got {getattr(self, name)}")
