"""The C-ABI shared library: loads, exports every symbol include/cave_b200.h declares, and its
argument checking / sizing entry points behave (no compute calls here — those need a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cave_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "cave_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cave_[a-z_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert "cave_forward_backward" in names and "cave_pack" in names and len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cave_b200.h but not exported"
    from cave_b200 import _lib
    assert sorted(_lib.EXPORTS) == names


def test_version_limits_and_sizes(lib):
    from cave_b200 import _lib
    assert lib.cave_abi_version() == 4
    lim = _lib.Limits()
    assert lib.cave_get_limits(ctypes.byref(lim)) == 0
    assert lim.max_d >= 4950 and lim.max_m >= 5200
    n = ctypes.c_size_t()
    assert lib.cave_pack_bytes(4096, 1337, 1225, ctypes.byref(n)) == 0
    assert 4096 * 1337 * 8 < n.value < 4096 * 1337 * 1225 * 4 // 15       # ~5% of dense A (incl. the cached solver setup)
    opts = _lib.SolverOpts(0, 0, 0.0, 128, 20000, 0, 0)
    assert lib.cave_scratch_bytes(4096, 1337, 1225, _lib.F64, ctypes.byref(opts), ctypes.byref(n)) == 0
    assert n.value > 0
    small = n.value
    assert lib.cave_scratch_bytes(4096, 1337, 1225, _lib.F64, None, ctypes.byref(n)) == 0
    assert n.value > small
    # dense (tensor-core Gram) path: its workspace is part of the scratch only where the host gate enables it
    assert lib.cave_scratch_bytes(512, 1024, 1225, _lib.F64, None, ctypes.byref(n)) == 0          # auto: m_max <= d
    with_dense = n.value
    off = _lib.SolverOpts(0, 0, 0.0, 0, 0, 0, -1)
    assert lib.cave_scratch_bytes(512, 1024, 1225, _lib.F64, ctypes.byref(off), ctypes.byref(n)) == 0
    assert with_dense > n.value + 512 * 1024 * 1024 * 4
    on = _lib.SolverOpts(0, 0, 0.0, 128, 20000, 0, 1)
    assert lib.cave_scratch_bytes(4096, 1337, 1225, _lib.F64, ctypes.byref(on), ctypes.byref(n)) == 0
    assert n.value > small + (1 << 30)


def test_argument_errors_are_codes_with_messages(lib):
    n = ctypes.c_size_t()
    assert lib.cave_pack_bytes(0, 10, 10, ctypes.byref(n)) == -1
    assert b"positive" in lib.cave_last_error()
    assert lib.cave_pack_bytes(8, 10, 10 ** 6, ctypes.byref(n)) == -2
    assert b"exceeds" in lib.cave_last_error()
    assert lib.cave_pack(None, None, 8, 10, 10, None, 0, None) == -1
    assert lib.cave_forward_backward(None, None, None, 8, 10, 10, -1.0, 0, 0.2, 0, 0, 1, None, None, None, None, None,
                                     None, None, None, None, 0, None, 0, None) == -1


def test_graft_entry_build_check_passes():
    """The driver's "does it build" entry point: compiles (or finds fresh) the library and checks its ABI version
    against the header."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    ge = importlib.import_module("__graft_entry__")
    ge.build()


def test_launch_plan_rule_on_host(lib):
    """cave_plan_choice applies the device's selection rule to a host copy of the plan words: the smallest configuration
    whose shared memory holds the average working set with 5 % to spare and the largest hot set outright."""
    from cave_b200 import _lib

    def choice(n, avg8, avg4, hot8, hot4, d=1225, io=_lib.F32, comp=_lib.F64):
        words = (ctypes.c_uint64 * 8)(n, n * avg8, n * avg4, hot8, hot4, 0, 0, 0)
        t, c, sm = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        idx = lib.cave_plan_choice(words, d, io, comp, ctypes.byref(t), ctypes.byref(c), ctypes.byref(sm))
        return idx, t.value, c.value, sm.value

    assert choice(4096, 12000, 10000, 11000, 9000) == (0, 64, 8, 27520)            # TSP-20-sized
    assert choice(4096, 67000, 59000, 48500, 39300)[0] == 2                           # TSP-50, float64 factor
    assert choice(4096, 67000, 59000, 48500, 39300, comp=_lib.F32)[0] == 2            # ... float32 factor: 59 KB > 0.95 * 55 KB
    assert choice(4096, 40000, 30000, 30000, 25000)[0] == 1
    assert choice(4096, 40000, 30000, 60000, 25000)[0] == 2                           # one instance's hot set rules out 128 x 4
    assert choice(16, 600000, 600000, 20000, 20000) == (3, 256, 2, 112640)            # dense Gram: widest CTA
    assert choice(0, 0, 0, 0, 0)[0] == 3                                              # empty statistics: the default
    assert choice(4096, 26000, 20000, 20000, 15000, d=1225, io=_lib.F64)[0] == 1      # float64 I/O adds 4 d bytes: 30.9 KB > 26.7 KB
