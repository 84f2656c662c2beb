"""CPU: the C-ABI library loads and exports every symbol include/gspaln.h declares
(no compute calls: there is no GPU here and no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    txt = (ROOT / "include" / "gspaln.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gspaln_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_surface():
    from spaln_b200 import capi
    assert declared_functions() == sorted(capi.EXPORTS)


def test_library_exports_all_symbols():
    from spaln_b200 import capi
    if not capi.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(str(capi.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), name
    lib.gspaln_version.restype = ctypes.c_char_p
    assert b"gspaln" in lib.gspaln_version()


def test_struct_layouts_match_header():
    from spaln_b200 import capi
    # sizes implied by include/gspaln.h on LP64
    assert ctypes.sizeof(capi.GspalnParams) == 4 * (8 + 8 + 8 + 5) + 4 * 32 * 32
    assert ctypes.sizeof(capi.GspalnTask) == 8 + 4 * 8 + 4 * 12 + 8 + 8     # ... + int53, cip pointers
    assert ctypes.sizeof(capi.GspalnResult) == 32 + 16 + 8
    assert ctypes.sizeof(capi.GspalnHParams) == 4 * (10 + 8 + 8 + 4) + 4 * 32 * 32 + 4 * 3
    assert ctypes.sizeof(capi.GspalnHTask) == 8 + 3 * 8 + 4 * 14 + 8 + 8     # ... + int53, cip pointers
    assert capi.SGPT6_DTYPE.itemsize == 14


def test_no_cpu_fallback_without_device():
    """Engine creation must fail loudly when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import golden_io
    from spaln_b200 import Engine, EngineError
    prm, _ = golden_io.load("dna_A2_global")
    with pytest.raises(EngineError):
        Engine(prm)


def test_task_cells_host_helper():
    from spaln_b200 import capi
    lib = capi.load()
    t = capi.GspalnTask()
    t.a_left, t.a_right, t.b_left, t.b_right, t.lw, t.up = 0, 10, 0, 50, -5, 45
    # row m: columns max(m-5,0) < n <= min(m+46,50)
    want = sum(min(m + 46, 50) - max(m - 5, 0) for m in range(1, 11))
    assert lib.gspaln_task_cells(ctypes.byref(t)) == want


def test_task_cells_closed_form_equals_loop_count():
    """gspaln_task_cells / gspaln_h_task_cells (host code, no GPU): the closed-form band area must
    equal the reference's inner-loop trip count, rows m in (a_left, a_right], columns
    max(k m + lw, b_left) < n <= min(k m + up + 1, b_right) with k = 1 (DNA, src/fwd2s1.cc:252-276)
    and k = 3 (protein, src/fwd2h1.cc:324-331 counts n0 .. n9 inclusive of both ends minus one)"""
    import numpy as np
    from spaln_b200 import capi
    lib = capi.load()
    rng = np.random.default_rng(5)
    for _ in range(300):
        al = int(rng.integers(0, 20)); ar = al + int(rng.integers(0, 60))
        bl = int(rng.integers(0, 50)); br = bl + int(rng.integers(0, 400))
        lw = int(rng.integers(-80, 200)); up = lw + int(rng.integers(0, 300))
        t = capi.GspalnTask()
        t.a_left, t.a_right, t.b_left, t.b_right, t.lw, t.up = al, ar, bl, br, lw, up
        want = sum(max(0, min(m + up + 1, br) - max(m + lw, bl)) for m in range(al + 1, ar + 1))
        assert lib.gspaln_task_cells(ctypes.byref(t)) == want, (al, ar, bl, br, lw, up)
