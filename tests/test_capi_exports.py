"""The product library must load and export every symbol include/eicos_b200.h declares (no compute
calls here: this tier has no GPU), and the product package must not route through the oracle."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "eicos_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eicos_[a-z_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from eicos_b200.binding import EXPORTS, PRODUCT_LIB
    assert os.path.exists(PRODUCT_LIB), "run __graft_entry__.build() first (nvcc, sm_100a)"
    lib = ctypes.CDLL(PRODUCT_LIB)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(EXPORTS) == names


def test_library_contains_sm100a_kernels():
    import shutil
    import subprocess
    from eicos_b200.binding import PRODUCT_LIB
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "-lelf", PRODUCT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_product_never_touches_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "eicos_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "eicos_oracle" not in src, f
    hdr = open(os.path.join(ROOT, "include", "eicos_b200.h")).read()
    assert "oracle" not in hdr.lower()


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    from eicos_b200.binding import Library
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Library(str(tmp_path / "libeicos_b200.so"))
