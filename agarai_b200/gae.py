"""rl.gae / rl.make_returns of the reference (rl/__init__.py:94-159; a2c/losses.py:66-88) on
time-major device buffers, through `agar_gae` of the C-ABI library.  No fallback exists."""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _prep(t, dtype):
    return None if t is None else t.to(dtype=dtype).contiguous()


def gae_and_returns(rewards: torch.Tensor, values: Optional[torch.Tensor], gamma: float, lam: float = 0.0,
                    dones: Optional[torch.Tensor] = None, end_value: Optional[torch.Tensor] = None,
                    want_adv: bool = True, want_ret: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """rewards [T,N] f32 cuda, values [T+1,N], dones [T,N] uint8 (optional; the reference has no
    masking inside, see ppo/train.py:189-195 and :284-290), end_value [N] (returns bootstrap).
    One launch computes both what `rl.gae` and `rl.make_returns` compute per agent in the
    Python loop of ppo/train.py:272-277."""
    if not rewards.is_cuda:
        raise _lib.AgarError("gae_and_returns needs CUDA tensors (there is no CPU fallback)")
    r = _prep(rewards, torch.float32)
    if r.dim() != 2:
        raise ValueError("rewards must be [T, N]")
    T, N = r.shape
    v = _prep(values, torch.float32)
    if want_adv and (v is None or tuple(v.shape) != (T + 1, N)):
        raise ValueError("values must be [T+1, N]")
    d = _prep(dones, torch.uint8)
    ev = _prep(end_value, torch.float32)
    # the kernel (TMA bulk copies included) trusts these shapes: a wrong device or a short buffer would fault the context
    for name, t in (("values", v), ("dones", d), ("end_value", ev)):
        if t is not None and t.device != r.device:
            raise ValueError(f"{name} must live on {r.device} like rewards (got {t.device})")
    if v is not None and not want_adv and tuple(v.shape) != (T + 1, N):
        raise ValueError("values must be [T+1, N]")
    if d is not None and tuple(d.shape) != (T, N):
        raise ValueError("dones must be [T, N]")
    if ev is not None and ev.numel() != N:
        raise ValueError("end_value must hold N elements")
    adv = torch.empty_like(r) if want_adv else None
    ret = torch.empty_like(r) if want_ret else None
    stream = C.c_void_p(torch.cuda.current_stream(r.device).cuda_stream)
    with torch.cuda.device(r.device):
        _lib.check(_lib.lib().agar_gae(_ptr(r), _ptr(v), _ptr(d), _ptr(ev), float(gamma), float(lam),
                                       _ptr(adv), _ptr(ret), T, N, stream))
    return adv, ret


def gae(rewards, values, gamma, lam, dones=None):
    """rl.gae(rewards, values, gamma, lam) — rl/__init__.py:94-137 — for all columns at once."""
    return gae_and_returns(rewards, values, gamma, lam, dones=dones, want_ret=False)[0]


def make_returns(rewards, gamma, end_value=None, dones=None):
    """rl.make_returns(rewards, gamma, end_value) — rl/__init__.py:140-159; with end_value None it
    is a2c.losses.make_returns (a2c/losses.py:66-78)."""
    return gae_and_returns(rewards, None, gamma, 0.0, dones=dones, end_value=end_value, want_adv=False)[1]
