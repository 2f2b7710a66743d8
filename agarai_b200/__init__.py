"""agarai_b200 — B200-native batched Agar.io engine behind the env interface of jondeaton/AgarAI.

The package holds only the env-step + grid-observation + GAE path: CUDA kernels and the C-ABI
(csrc/, include/agar_b200.h) and the host-side mirror of the reference's interface.  Importing
it does not need a GPU; constructing an env or calling an op without the built CUDA library or
without a CUDA device raises (there is no CPU fallback).
"""
from ._lib import AgarConfig, AgarError, AgarLayout, default_config
from .actions import ActionConfig, ActionConverter, action_size, ppo_action_table, unravel_action_table
from .env import BatchedAgarioEnv, GymCompatEnv, config_from_gym_kwargs, make
from .features import FeatureExtractor, GridFeatureExtractor, ScreenFeatureExtractor
from .gae import gae, gae_and_returns, make_returns
from .rollout import DeviceRollout
from .vector import Coordinator, RemoteEnvironment

__all__ = [
    "AgarConfig", "AgarError", "AgarLayout", "default_config", "ActionConfig", "ActionConverter",
    "action_size", "ppo_action_table", "unravel_action_table", "BatchedAgarioEnv", "GymCompatEnv",
    "config_from_gym_kwargs", "make", "gae", "gae_and_returns", "make_returns", "DeviceRollout",
    "FeatureExtractor", "GridFeatureExtractor", "ScreenFeatureExtractor", "Coordinator", "RemoteEnvironment",
]
