"""Discrete action index -> (target_xy, act) tables.

Mirrors the two action decodes of the reference:
  * PPO:  ppo/action_conversion.py:9-13 (`action_size`) and :19-80 (`make_action_converter`)
  * A2C / DQN: a2c/train.py:67-79 (`agario_to_action`), dqn/training.py:220-228 (`to_action`),
    both `np.unravel_index` over `(num_directions, num_magnitudes, num_acts)`.
The tables are built once on the host (trigonometry is not reproducible bit for bit between
libms, so it never runs inside the simulator) and uploaded with `agar_set_action_table`; the
step kernel then does the decode as a table lookup fused into the step (agar_step_indexed).
"""
from dataclasses import dataclass
from typing import Tuple

import numpy as np


@dataclass(frozen=True)
class ActionConfig:
    """configuration/environment.proto:27-32 `Action`."""
    num_directions: int = 8
    num_magnitudes: int = 1
    allow_splitting: bool = False
    allow_feeding: bool = False


def action_size(cfg: ActionConfig) -> int:
    """ppo/action_conversion.py:9-13."""
    num_acts = int(cfg.allow_splitting) + int(cfg.allow_feeding)
    move_size = cfg.num_directions * cfg.num_magnitudes
    return 1 + move_size + cfg.num_directions * num_acts


def ppo_action_table(cfg: ActionConfig) -> Tuple[np.ndarray, np.ndarray]:
    """Table form of ppo/action_conversion.py:19-80, evaluated in float32 like the jitted
    original: index 0 = no-op; next D*M = moves (theta = 2*pi*k/D, mag = 1 - m/(1+M)); then
    D per enabled act at magnitude 1.  With both acts enabled the reference computes
    `1 + index / D` as a float (:52); the integer part (1 = split, 2 = feed) is what an
    integer `act` argument can mean, so that is what the table holds."""
    D, M = cfg.num_directions, cfg.num_magnitudes
    n = action_size(cfg)
    xy = np.zeros((n, 2), np.float32)
    act = np.zeros((n,), np.int32)
    two_pi = np.float32(2 * np.pi)

    def direction(mag, theta):
        return np.array([np.cos(theta) * mag, np.sin(theta) * mag], np.float32)

    num_moves = D * M
    for index in range(1, n):
        i = index - 1
        if i < num_moves:
            theta = two_pi * np.float32(i % D) / np.float32(D)
            mag = np.float32(1 - (i // D) / (1 + M))
            xy[index] = direction(mag, theta)
        else:
            j = i - num_moves
            theta = two_pi * np.float32(j % D) / np.float32(D)
            if cfg.allow_splitting and cfg.allow_feeding:
                a = 1 + j // D
            elif cfg.allow_splitting:
                a = 1
            elif cfg.allow_feeding:
                a = 2
            else:
                a = 0
            xy[index] = direction(np.float32(1.0), theta)
            act[index] = a
    return xy, act


def unravel_action_table(action_shape) -> Tuple[np.ndarray, np.ndarray]:
    """Table form of a2c/train.py:67-79 / dqn/training.py:220-228 (float64 trigonometry as in
    the originals, stored as the float32 the engine consumes)."""
    D, M, acts = action_shape
    n = D * M * acts
    xy = np.zeros((n, 2), np.float32)
    act = np.zeros((n,), np.int32)
    for index in range(n):
        i0, i1, i2 = np.unravel_index(index, action_shape)
        theta = (2 * np.pi * i0) / D
        mag = 1 - i1 / M
        xy[index] = (np.cos(theta) * mag, np.sin(theta) * mag)
        act[index] = int(i2)
    return xy, act


class ActionConverter:
    """Host-side callable with the contract of `make_action_converter(...)`'s result
    (ppo/action_conversion.py:72-78): index -> (np.ndarray[2], int)."""

    def __init__(self, xy: np.ndarray, act: np.ndarray):
        self.xy, self.act = xy, act

    @classmethod
    def ppo(cls, cfg: ActionConfig):
        return cls(*ppo_action_table(cfg))

    @classmethod
    def unravel(cls, action_shape):
        return cls(*unravel_action_table(action_shape))

    def __len__(self):
        return len(self.act)

    def __call__(self, index):
        index = int(index)
        return self.xy[index].copy(), int(self.act[index])
