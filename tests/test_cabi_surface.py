"""CPU: the C-ABI library loads, exports every function include/luw_cuda.h declares (and nothing the header does not know), the ctypes table covers
the same set, and without a CUDA device the compute entry points fail loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "luw_cuda.h")


def declared():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return set(re.findall(r"\b(?:int|const char\*)\s+(luw_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_are_exported_and_bound():
    from latticeurbanwind_b200 import _cabi as A
    L = A.lib()
    names = declared()
    assert len(names) >= 40
    exported = {l.split()[-1] for l in subprocess.check_output(["nm", "-D", "--defined-only", A.LIB_PATH], text=True).splitlines() if " T " in l}
    missing = names - exported
    assert not missing, f"declared but not exported: {sorted(missing)}"
    extra = {s for s in exported if s.startswith("luw_")} - names
    assert not extra, f"exported but not declared in the header: {sorted(extra)}"
    assert set(A.EXPORTS) | {"luw_last_error_string"} == names, sorted(names ^ (set(A.EXPORTS) | {"luw_last_error_string"}))
    for n in names:
        assert getattr(L, n) is not None


def test_every_entry_point_cites_the_reference():
    """Each declaration (or the block comment in front of it) names the reference interface it replaces (FX/...:line)."""
    text = open(HEADER).read()
    assert text.count("FX/") >= 30


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    from latticeurbanwind_b200 import _cabi as A
    p = A.DomainParams(16, 8, 8, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1.0, 0, 1, 0.0, 0, 1, 0.0, 0)
    h = C.c_void_p()
    rc = A.lib().luw_domain_create(C.byref(p), C.byref(h))
    assert rc == A.ERR_NO_DEVICE and not h.value
    with pytest.raises(A.LuwError):
        A.check(rc)
    assert A.device_count() == 0


def test_inlet_searches_need_a_device_too():
    """SURVEY 8-f2 entry points: without a CUDA device the sample searches return an error (no host fallback inside the library)."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    from latticeurbanwind_b200 import _cabi as A
    L = A.lib()
    cell, pts = np.zeros(6, np.float32), np.zeros(6, np.float32)
    near, kept, used, mr, ex = np.zeros(2, np.uint32), np.zeros(128, np.uint32), np.zeros(2, np.uint32), np.zeros(2, np.float32), np.zeros(2, np.int32)
    assert L.luw_inlet_nearest(0, 2, cell.ctypes.data, 2, pts.ctypes.data, near.ctypes.data) != 0
    assert L.luw_last_error_string()
    assert L.luw_inlet_knn(0, 2, cell.ctypes.data, 2, pts.ctypes.data, kept.ctypes.data, used.ctypes.data, mr.ctypes.data, ex.ctypes.data) != 0
    with pytest.raises(A.LuwError):
        A.check(L.luw_inlet_knn(0, 2, cell.ctypes.data, 2, pts.ctypes.data, kept.ctypes.data, used.ctypes.data, mr.ctypes.data, ex.ctypes.data))


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 on its own (no C++ or torch types in the signatures)."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
