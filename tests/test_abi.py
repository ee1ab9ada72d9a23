"""The C-ABI library loads and exports every symbol include/trs.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "trs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trs_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_all_declared_symbols():
    from torchrecsys_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    declared = _declared()
    assert declared, "no functions parsed from trs.h"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in trs.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding list out of sync with trs.h"
    assert L.trs_abi_version() == 1


def test_struct_sizes_match_header():
    from torchrecsys_b200 import _lib
    assert ctypes.sizeof(_lib.Table) == 7 * 8
    assert ctypes.sizeof(_lib.Model) == 16 + (2 + _lib.MAX_META) * 56
    assert ctypes.sizeof(_lib.Epoch) == 5 * 8 + 8 + 8
    assert ctypes.sizeof(_lib.Optim) == 8 + 3 * 8 + 8


def test_argument_errors_are_reported_not_raised_across_the_abi():
    from torchrecsys_b200 import _lib
    L = _lib.lib()
    rc = L.trs_scores(None, None, None, None, ctypes.c_int64(4), None, None)
    assert rc == -1
    assert b"NULL" in L.trs_last_error()
