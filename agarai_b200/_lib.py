"""ctypes loader for libagar_b200.so (the C-ABI of include/agar_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing or cannot be loaded this
module raises, and every op of the package fails with it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGAR_B200_LIB") or os.path.join(_HERE, "libagar_b200.so")  # env override: tuning builds


class AgarConfig(C.Structure):
    """Field-for-field mirror of `AgarConfig` (include/agar_b200.h)."""
    _fields_ = [
        ("num_arenas", C.c_int32), ("num_agents", C.c_int32), ("num_bots", C.c_int32),
        ("num_pellets", C.c_int32), ("num_viruses", C.c_int32), ("max_foods", C.c_int32),
        ("ticks_per_step", C.c_int32), ("grid_size", C.c_int32), ("obs_flags", C.c_int32),
        ("pellet_regen", C.c_int32), ("auto_reset", C.c_int32), ("event_capacity", C.c_int32),
        ("bots_split", C.c_int32), ("reserved0", C.c_int32), ("arena_id_base", C.c_int64),
        ("arena_size", C.c_float), ("view_size", C.c_float), ("start_mass", C.c_float),
        ("pellet_mass", C.c_float), ("food_mass", C.c_float), ("virus_mass", C.c_float),
        ("split_min_mass", C.c_float), ("feed_min_mass", C.c_float), ("virus_pop_mass", C.c_float),
        ("eat_ratio", C.c_float), ("r2_per_mass", C.c_float), ("speed_k", C.c_float),
        ("target_reach", C.c_float), ("split_speed", C.c_float), ("impulse_decay", C.c_float),
        ("feed_speed", C.c_float), ("food_decay", C.c_float), ("food_spawn_gap", C.c_float),
        ("pop_speed", C.c_float), ("merge_delay", C.c_float), ("pop_pieces", C.c_int32),
        ("bot_roam_period", C.c_int32), ("bot_sense_radius", C.c_float), ("mass_decay", C.c_float),
    ]


class AgarLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "words", "pellet_x", "pellet_y", "virus_x", "virus_y", "dyn_begin", "cell_x", "cell_y",
        "cell_ix", "cell_iy", "cell_m", "cell_t", "food_x", "food_y", "food_vx", "food_vy",
        "agent_done", "tick", "episode_steps", "dyn_end", "num_players", "cells_total",
        "P_pad", "V_pad", "F_pad", "A_pad")]


# every symbol include/agar_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("agar_last_error", C.c_char_p, []),
    ("agar_abi_version", C.c_int, []),
    ("agar_config_default", C.c_int, [C.POINTER(AgarConfig)]),
    ("agar_layout_compute", C.c_int, [C.POINTER(AgarConfig), C.POINTER(AgarLayout)]),
    ("agar_obs_channels", C.c_int, [C.POINTER(AgarConfig)]),
    ("agar_create", C.c_int, [C.POINTER(AgarConfig), C.c_int, C.POINTER(_P)]),
    ("agar_destroy", C.c_int, [_P]),
    ("agar_get_config", C.c_int, [_P, C.POINTER(AgarConfig)]),
    ("agar_get_layout", C.c_int, [_P, C.POINTER(AgarLayout)]),
    ("agar_seed", C.c_int, [_P, C.c_uint64]),
    ("agar_reset", C.c_int, [_P, _P, _P, _P]),
    ("agar_step", C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    ("agar_set_action_table", C.c_int, [_P, _P, _P, C.c_int32]),
    ("agar_step_indexed", C.c_int, [_P, _P, _P, _P, _P, _P]),
    ("agar_observe", C.c_int, [_P, _P, _P]),
    ("agar_observe_records", C.c_int, [_P, _P, C.c_int32, _P, _P]),
    ("agar_get_state", C.c_int, [_P, _P, _P]),
    ("agar_set_state", C.c_int, [_P, _P, _P]),
    ("agar_events", C.c_int, [_P, _P, _P, _P]),
    ("agar_step_host", C.c_int, [_P, _P, _P, _P, _P, _P]),
    ("agar_host_obs_persistent", C.c_int, [_P, C.c_int]),
    ("agar_gae", C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, _P, _P, C.c_int32, C.c_int32, _P]),
    ("agar_ram_feature_size", C.c_int, [C.c_int32] * 5),
    ("agar_observe_ram", C.c_int, [_P] + [C.c_int32] * 5 + [_P, _P]),
    ("agar_screen_gray", C.c_int, [_P, C.c_int64, _P, C.c_int32, _P]),
    ("agar_launch_count", C.c_int64, [_P]),
    ("agar_host_copy_bytes", C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("agar_kernel_variant", C.c_char_p, [_P]),
    ("agar_step_smem_bytes", C.c_int64, [C.POINTER(AgarConfig)]),
]

_lib = None


class AgarError(RuntimeError):
    pass


def load(path):
    """Load one build of the CUDA library (the default one, or a specialised copy from build.build_specialised)."""
    if not os.path.exists(path):
        raise AgarError(
            f"{path} is missing: build it with `python -m agarai_b200.build` "
            "(nvcc, sm_100a). agarai_b200 has no CPU fallback.")
    L = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(L, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if L.agar_abi_version() != 1:
        raise AgarError("libagar_b200.so ABI version mismatch")
    return L


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        _lib = load(LIB_PATH)
    return _lib


def check(rc, L=None):
    """Raises AgarError with the message of the library `L` that returned `rc` (default: the default build)."""
    if rc != 0:
        raise AgarError((L or lib()).agar_last_error().decode())


def default_config(**overrides):
    cfg = AgarConfig()
    check(lib().agar_config_default(C.byref(cfg)))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown AgarConfig field {k!r}")
        setattr(cfg, k, v)
    return cfg
