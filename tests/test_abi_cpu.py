"""The C-ABI shared library loads and exports every symbol include/mft_b200.h declares (no compute calls: CPU box)."""
import ctypes as C
import os
import re

import cases


def _mft():
    import mft_b200

    return mft_b200


def _declared():
    txt = open(os.path.join(cases.ROOT, "include", "mft_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mft_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    m = _mft()
    lib = C.CDLL(m._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mft_b200.h but not exported"
    assert set(m._lib.EXPORTS) == set(names)


def test_signatures_have_no_torch_types():
    txt = open(os.path.join(cases.ROOT, "include", "mft_b200.h")).read()
    assert "torch" not in txt.lower() and "at::" not in txt
    assert 'extern "C"' in txt


def test_no_cpu_fallback_without_device():
    m = _mft()
    lib = m._lib.load()
    if lib.mft_device_count() > 0:
        return
    ctx = C.c_void_p()
    rc = lib.mft_ctx_create(C.byref(ctx), 0, 100, 0, 4, 2, 20)
    assert rc == -3 and b"no CPU fallback" in lib.mft_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(cases.ROOT, "meshfreetrixi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mft_oracle" not in src and "oracle/" not in src, f
