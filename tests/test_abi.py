"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/dwgpu.h declares, and (without a GPU) creation fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_example
from dwsolver_b200 import engine as E


def test_library_exports_every_declared_symbol():
    lib = E.load_library()
    header = open(os.path.join(ROOT, "include", "dwgpu.h")).read()
    declared = set(re.findall(r"\b(dwgpu_[a-z_]+)\s*\(", header))
    assert declared == set(E.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dwgpu_version()


def test_struct_layouts_match_header():
    # sizes implied by include/dwgpu.h on LP64
    assert C.sizeof(E.BlockDesc) == 8 + 9 * 8 + 8 + 3 * 8
    assert C.sizeof(E.Proposals) == 7 * 8
    assert C.sizeof(E.Counters) == 9 * 8
    assert C.sizeof(E.Mailboxes) == 6 * 8 + 2 * 4
    assert C.sizeof(E.Packed) == 7 * 8
    o = E.Options()
    E.load_library().dwgpu_options_init(C.byref(o))
    assert (o.device, o.clamp_lo_cols, o.force_path, o.max_iter_mult) == (0, 1, 0, 50)
    assert (o.tol_dj, o.tol_bnd, o.tol_piv) == (1e-9, 1e-9, 1e-9)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.DwGpuError, match="CUDA"):
        E.Engine(load_example("book_lasdon"))


def test_bad_arguments_are_rejected():
    lib = E.load_library()
    h = C.c_void_p()
    assert lib.dwgpu_create(0, 1, None, None, C.byref(h)) == -1
    assert lib.dwgpu_create(1, 1, None, None, None) == -1
    assert b"" != lib.dwgpu_last_error()
