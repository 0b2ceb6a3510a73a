"""CPU: the C-ABI library loads, exports every symbol include/magic_sht.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "magic_sht.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(magic_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from magic_b200.lib import SYMBOLS, load_library
    lib = load_library()
    declared = header_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/magic_sht.h but not exported"
    assert sorted(SYMBOLS) == declared  # the Python binding list tracks the header


def test_no_cpu_fallback():
    import torch
    from magic_b200 import MagicError, Sht
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(MagicError, match="no CUDA device"):
        Sht(16)


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under magic_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "magic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "kernels_fft.cuh", (dirpath, f)


def test_params_struct_layout_matches_header():
    """ctypes mirror of magic_params has the header's field order."""
    from magic_b200.riter import Params
    src = open(os.path.join(ROOT, "include", "magic_sht.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} magic_params;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        stmt = re.sub(r"^(int|double)\s+", "", stmt)
        names += [n.strip() for n in stmt.split(",")]
    assert names == [n for n, _ in Params._fields_]


def test_log_step_entry_points_reject_null_handles_loudly():
    """Argument checks come before any device work: a null plan is an error with a message, never a crash or a silent no-op."""
    from ctypes import c_double, c_int, c_void_p
    from magic_b200.lib import load_library
    lib = load_library()
    null = c_void_p(None)
    calls = {
        "magic_rloop_diagnostics": (null, null, c_int(1), c_int(1), c_int(1), null),
        "magic_rloop_dtb": (null, null, null),
        "magic_rloop_to_next": (null, null),
        "magic_rloop_to": (null, null, c_double(1e-3), null),
        "magic_rloop_rms_keep": (null, null),
        "magic_rloop_rms": (null, null, c_double(1e-3), null),
    }
    for name, args in calls.items():
        assert getattr(lib, name)(*args) != 0, name
        assert b"null argument" in lib.magic_last_error(), (name, lib.magic_last_error())
