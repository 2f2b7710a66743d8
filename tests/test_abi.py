"""The C-ABI shared library loads on a box without a GPU and exports every symbol that
include/agar_b200.h declares; the ctypes mirrors of the structs match the C layout.  No compute
call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "agar_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int)\s+(agar_\w+)\s*\(", txt, flags=re.M)
    assert len(names) >= 20
    return sorted(set(names))


@pytest.fixture(scope="module")
def built_lib():
    from agarai_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/agar_b200.h but not exported"


def test_ctypes_table_covers_the_header(built_lib):
    from agarai_b200 import _lib
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == declared_functions()
    L = _lib.lib()
    assert L.agar_abi_version() == 1


def test_struct_layouts_match_c(tmp_path, built_lib):
    from agarai_b200 import _lib
    src = tmp_path / "probe.c"
    cfg_fields = [n for n, _ in _lib.AgarConfig._fields_]
    lay_fields = [n for n, _ in _lib.AgarLayout._fields_]
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
            'printf("%zu %zu\\n", sizeof(AgarConfig), sizeof(AgarLayout));']
    body += [f'printf("%zu\\n", offsetof(AgarConfig, {f}));' for f in cfg_fields]
    body += [f'printf("%zu\\n", offsetof(AgarLayout, {f}));' for f in lay_fields]
    body += ['return 0;}']
    src.write_text("\n".join(body))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == C.sizeof(_lib.AgarConfig) and int(out[1]) == C.sizeof(_lib.AgarLayout)
    offs = [int(x) for x in out[2:]]
    for f, o in zip(cfg_fields, offs[:len(cfg_fields)]):
        assert getattr(_lib.AgarConfig, f).offset == o, f
    for f, o in zip(lay_fields, offs[len(cfg_fields):]):
        assert getattr(_lib.AgarLayout, f).offset == o, f


def test_config_layout_and_errors_without_gpu(built_lib, oracle):
    """Pure host entry points work without a device; device entry points fail loudly."""
    from agarai_b200 import _lib
    L = _lib.lib()
    cfg = _lib.default_config(num_arenas=3, num_agents=2, num_bots=3, num_pellets=333, num_viruses=7, max_foods=9)
    lay = _lib.AgarLayout()
    _lib.check(L.agar_layout_compute(C.byref(cfg), C.byref(lay)))
    olay = oracle.layout(oracle.default_config(num_arenas=3, num_agents=2, num_bots=3, num_pellets=333,
                                               num_viruses=7, max_foods=9))
    for name, _ in _lib.AgarLayout._fields_:  # library and oracle agree on the record format
        assert getattr(lay, name) == getattr(olay, name), name
    assert lay.words % 4 == 0 and lay.dyn_begin % 4 == 0 and lay.cells_total == 5 * 16
    assert L.agar_obs_channels(C.byref(cfg)) == 4
    bad = _lib.default_config(grid_size=30)
    assert L.agar_layout_compute(C.byref(bad), C.byref(lay)) != 0
    assert b"grid_size" in L.agar_last_error()
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert L.agar_create(C.byref(cfg), 0, C.byref(h)) != 0  # no device: an error, never a CPU fallback
        import agarai_b200 as ab
        with pytest.raises(ab.AgarError):
            ab.BatchedAgarioEnv(cfg)
        with pytest.raises(ab.AgarError):
            ab.gae(torch.zeros(4, 2), torch.zeros(5, 2), 0.9, 0.5)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under agarai_b200/ may reference it."""
    pkg = os.path.join(ROOT, "agarai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "agar_oracle" not in txt.replace("oracle/agar_oracle.c", ""), f


def test_per_config_specialised_build_exports_the_same_abi(built_lib):
    """build.build_specialised (nvcc, cached in-tree): the per-config copy of the library exports every declared
    symbol and computes the same layout as the default build (no GPU needed to check that)."""
    import shutil
    if not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        pytest.skip("nvcc is not on this machine")
    import ctypes as C
    from agarai_b200 import _lib, build, config_from_gym_kwargs
    cfg = config_from_gym_kwargs(num_envs=9, num_agents=2, num_bots=3, num_pellets=500, num_viruses=5, arena_size=200.0,
                                 grid_size=32, observe_viruses=True, difficulty="normal")
    path = build.build_specialised(cfg)
    assert os.path.exists(path) and os.path.dirname(path).endswith("_jit")
    L = _lib.load(path)
    for name in declared_functions():
        assert hasattr(L, name), name
    D = _lib.load(built_lib)  # the default build
    a, b = _lib.AgarLayout(), _lib.AgarLayout()
    assert L.agar_layout_compute(C.byref(cfg), C.byref(a)) == 0 and D.agar_layout_compute(C.byref(cfg), C.byref(b)) == 0
    assert bytes(a) == bytes(b)
    assert L.agar_step_smem_bytes(C.byref(cfg)) == D.agar_step_smem_bytes(C.byref(cfg)) > 0
