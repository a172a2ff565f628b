"""The C-ABI shared library loads and exports every symbol include/o2a.h declares (no compute)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "o2a.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(o2a_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ortho2align_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/o2a.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert lib.o2a_abi_version() == 1


def test_ingest_library_exports_every_declared_symbol():
    from ortho2align_b200 import build, ingest
    build.build_ingest()
    lib = ingest.load_lib()
    with open(os.path.join(ROOT, "include", "o2a_ingest.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(o2a_ingest_[a-z_0-9]+)\s*\(", src)))
    assert set(names) == set(ingest.EXPORTS)
    for n in names:
        assert hasattr(lib, n)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ortho2align_b200 import _lib
    from ortho2align_b200.engine import Engine, O2AError
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.o2a_create(C.byref(h), 0) != 0
    with pytest.raises(O2AError):
        Engine(0)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ortho2align_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert "oracle" not in txt.replace("test oracle", "").replace("no reference oracle", "").replace("parity oracle", "").lower() \
                    or fn in ("synth.py",), f"{fn} mentions the oracle"
