"""Device-resident rollout buffer with the semantics of rollout/multi_rollout.py:25-64.

`MultiRollout.record(dict)` appends one `[A]`-long entry per key per step and `batched(key)`
transposes `[T][A]` into per-agent `[T, ...]` arrays (`transpose_batch`, :17-22).  Here every key
is one preallocated time-major tensor `[T_max, B, ...]` (B = num_envs * num_agents columns), so
recording is a device copy into row t and GAE/returns run on the buffer in place; `batched(key)`
gives the `[B, T, ...]` view a reference-shaped consumer indexes per agent.
"""
from typing import Dict

import torch


class DeviceRollout:
    def __init__(self, capacity: int, device=None):
        self.capacity = int(capacity)
        self.device = device
        self._buf: Dict[str, torch.Tensor] = {}
        self._len: Dict[str, int] = {}

    def __getitem__(self, key):
        """Time-major `[t, B, ...]` slice of what has been recorded under `key`."""
        return self._buf[key][: self._len[key]]

    def __contains__(self, key):
        return key in self._buf

    def __len__(self):
        return max(self._len.values(), default=0)

    def record(self, item_dict):
        """rollout/multi_rollout.py:39-41: one step's value for each key (keys may advance at
        different rates: `values` gets one extra bootstrap row, ppo/train.py:189-195)."""
        for key, value in item_dict.items():
            v = torch.as_tensor(value, device=self.device)  # [B, ...]: one row per agent column
            if key not in self._buf:
                rows = self.capacity + 1
                self._buf[key] = torch.empty((rows,) + tuple(v.shape), dtype=v.dtype, device=v.device)
                self._len[key] = 0
            t = self._len[key]
            if t >= self._buf[key].shape[0]:
                raise IndexError(f"rollout buffer for {key!r} is full")
            self._buf[key][t].copy_(v, non_blocking=True)
            self._len[key] = t + 1

    def batched(self, key):
        """rollout/multi_rollout.py:46-47: per-agent view `[B, T, ...]` of a key."""
        return self[key].transpose(0, 1)

    def clear(self):
        for k in self._len:
            self._len[k] = 0
