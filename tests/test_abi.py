"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/ldd_b200.h declares; the package refuses to run without CUDA (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ldd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "lddecode_b200", "csrc"))
    import build as prod_build
    path = prod_build.build()
    from lddecode_b200 import _lib
    lib = _lib.load(path)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "ctypes signature missing for " + n
    assert lib.ldd_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lddecode_b200 import rfdecode
    with pytest.raises(RuntimeError):
        rfdecode.RFDecode(8 * 315 / 88, 'NTSC')


def test_product_does_not_import_oracle_or_emu():
    pkg = os.path.join(ROOT, "lddecode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) and "build" not in dirpath:
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libldd_emu" not in src, f
