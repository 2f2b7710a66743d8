"""The reference's config front-end (configuration/__init__.py:26-73) without protoc or generated *_pb2
modules: a small reader for the textproto subset the reference's configs use, the proto2 defaults of
configuration/environment.proto and config.proto, and the same `load`, `gym_env_name`, `gym_env_config`
functions with the same keys, so `--config ppo` style entry points work against the batched engine.

    cfg = configuration.load("ppo")                      # or a path to a .textproto
    env = agarai_b200.make(configuration.gym_env_name(cfg.environment), num_envs=N,
                           **configuration.gym_env_config(cfg.environment))

Only host-side parsing lives here (no compute).  Unknown fields raise, as text_format.Merge does (the
reference's sweep template configs/sweeps/gamma_lam.textproto:35 sets a non-existent `learning_rate`
field and is rejected by protobuf too)."""
import os
import re
from typing import Any, Dict

# ---------------------------------------------------------------------------------- schema
# field -> default (scalar), or a nested schema dict (message), or ("enum", names, default)
_SCHEDULE = {
    "shift": 0.0, "scale": 1.0, "constant": None,
    "inverse_time_decay": {"base": 0.0, "decay_rate": 0.0, "decay_steps": 0.0},
    "exponential_decay": {"base": 0.0, "decay_rate": 0.0, "decay_steps": 0.0},
}
_ENVIRONMENT = {  # configuration/environment.proto:6-64
    "agario": {
        "difficulty": ("enum", ("UNKNOWN", "TRIVIAL", "EMPTY", "NORMAL"), "UNKNOWN"),
        "arena_size": 0, "num_pellets": 0, "num_viruses": 0, "num_bots": 0, "pellet_regen": False,
    },
    "num_agents": 1,
    "observation": {
        "type": ("enum", ("UNKNOWN", "GRID", "RAM", "SCREEN"), "UNKNOWN"),
        "num_frames": 1, "ticks_per_step": 4, "grid_size": 64,
        "pellets": True, "viruses": True, "cells": True, "others": True,
    },
    "action": {"num_directions": 0, "num_magnitudes": 0, "allow_splitting": False, "allow_feeding": False},
}
_CONFIG = {  # configuration/config.proto:8-90
    "environment": _ENVIRONMENT,
    "test_environment": _ENVIRONMENT,
    "hyperparameters": {
        "seed": 0, "feature_extractor_lr": _SCHEDULE, "actor_lr": _SCHEDULE, "critic_lr": _SCHEDULE,
        "num_episodes": 0, "episode_length": 0, "gamma": 0.0, "gae": {"lam": 0.0},
        "loss": {"ppo": {"clip_epsilon": 0.0}}, "batch": False, "batch_size": 0, "num_sgd_steps": 0,
        "num_environments": 0,
    },
    "model": {"feature_extractor": {"num_features": 0}},
    "training": {"save_frequency": 0},
}


class Message:
    """Attribute-style view of a parsed message with proto2 defaults; `HasField(name)` as in protobuf."""

    def __init__(self, schema: dict, name: str = "Config"):
        object.__setattr__(self, "_schema", schema)
        object.__setattr__(self, "_set", {})
        object.__setattr__(self, "_name", name)

    def HasField(self, field: str) -> bool:
        if field not in self._schema:
            raise ValueError(f'Protocol message {self._name} has no "{field}" field.')
        return field in self._set

    def __getattr__(self, field):
        schema = object.__getattribute__(self, "_schema")
        if field not in schema:
            raise AttributeError(field)
        if field in self._set:
            return self._set[field]
        spec = schema[field]
        if isinstance(spec, dict):
            return Message(spec, field)  # unset sub-message reads as all defaults
        if isinstance(spec, tuple):
            return spec[2]
        return spec

    def __setattr__(self, field, value):
        if field not in self._schema:
            raise AttributeError(f'{self._name} has no field "{field}"')
        self._set[field] = value

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, Message) else v) for k, v in self._set.items()}

    def __repr__(self):
        return f"{self._name}({self.to_dict()})"


# ---------------------------------------------------------------------------------- text format
_TOKEN = re.compile(r"""\s*(?:(\#[^\n]*)|([{}:<>;,])|("(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')|([^\s{}:<>;,\#"']+))""")


def _tokens(text: str):
    pos = 0
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip():
                raise ValueError(f"textproto: cannot tokenise at {text[pos:pos + 20]!r}")
            return
        pos = m.end()
        if m.group(1) is None:
            yield m.group(2) or m.group(3) or m.group(4)


def _scalar(spec, tok: str, field: str):
    if isinstance(spec, tuple):  # enum: name or number
        names = spec[1]
        if tok in names:
            return tok
        if tok.lstrip("-").isdigit() and 0 <= int(tok) < len(names):
            return names[int(tok)]
        raise ValueError(f'textproto: enum field "{field}" has no value named {tok}')
    if isinstance(spec, bool):
        if tok in ("true", "True", "t", "1"):
            return True
        if tok in ("false", "False", "f", "0"):
            return False
        raise ValueError(f'textproto: bad bool for "{field}": {tok}')
    if isinstance(spec, int):
        return int(tok, 0)
    if isinstance(spec, float) or spec is None:
        return float(tok.rstrip("fF"))
    if isinstance(spec, str):
        return tok[1:-1] if tok[:1] in "\"'" else tok
    raise ValueError(f'textproto: field "{field}" is a message; expected {{')


def _parse_message(toks, schema: dict, name: str, closer) -> Message:
    msg = Message(schema, name)
    while True:
        tok = next(toks, None)
        if tok is None:
            if closer is None:
                return msg
            raise ValueError(f"textproto: message {name} is not closed")
        if tok == closer:
            return msg
        if tok in (";", ","):
            continue
        if tok not in schema:
            raise ValueError(f'textproto: message type "{name}" has no field named "{tok}"')
        spec = schema[tok]
        nxt = next(toks)
        if nxt == ":":
            nxt = next(toks)
        if nxt in ("{", "<"):
            if not isinstance(spec, dict):
                raise ValueError(f'textproto: field "{tok}" of {name} is not a message')
            msg._set[tok] = _parse_message(toks, spec, tok, "}" if nxt == "{" else ">")
        else:
            msg._set[tok] = _scalar(spec, nxt, tok)


def parse(text: str, schema: dict = None, name: str = "Config") -> Message:
    return _parse_message(_tokens(text), _CONFIG if schema is None else schema, name, None)


# ---------------------------------------------------------------------------------- reference API
def _configs_dir() -> str:
    return os.path.join(os.path.dirname(__file__), "configs")


def load(name: str) -> Message:
    """configuration/__init__.py:31-35: a configuration by name (file `<name>.textproto` in the package's
    configs directory) or by path."""
    path = name if os.path.exists(name) else os.path.join(_configs_dir(), f"{name}.textproto")
    with open(path, "r") as f:
        return parse(f.read())


_gym_names = {"GRID": "agario-grid-v0", "RAM": "agario-ram-v0", "SCREEN": "agario-screen-v0"}


def gym_env_name(environment: Message) -> str:
    """configuration/__init__.py:38-46."""
    return _gym_names[environment.observation.type]


def gym_env_config(environment: Message) -> Dict[str, Any]:
    """configuration/__init__.py:49-73, same keys in the same order."""
    env_config = {
        "num_agents": environment.num_agents,
        "multi_agent": True,
        "difficulty": environment.agario.difficulty.lower(),
        "ticks_per_step": environment.observation.ticks_per_step,
        "arena_size": environment.agario.arena_size,
        "num_pellets": environment.agario.num_pellets,
        "num_viruses": environment.agario.num_viruses,
        "num_bots": environment.agario.num_bots,
        "pellet_regen": environment.agario.pellet_regen,
    }
    if environment.observation.type == "GRID":
        env_config["grid_size"] = environment.observation.grid_size
        env_config["num_frames"] = environment.observation.num_frames
        env_config["observe_cells"] = environment.observation.cells
        env_config["observe_others"] = environment.observation.others
        env_config["observe_pellets"] = environment.observation.pellets
        env_config["observe_viruses"] = environment.observation.viruses
    return env_config


def action_config(environment: Message):
    """The discrete action set of ppo/train.py:236-240: ActionConfig(num_directions, num_magnitudes, split, feed)."""
    from .actions import ActionConfig
    a = environment.action
    return ActionConfig(int(a.num_directions), int(a.num_magnitudes), bool(a.allow_splitting), bool(a.allow_feeding))
