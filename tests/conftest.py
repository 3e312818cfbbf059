import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FIXTURES = sorted(json.load(open(os.path.join(ROOT, "tests", "golden", "fixtures", "manifest.json"))).keys(), key=str.lower)
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libeicos_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulator of the tile program (test infrastructure; see tests/emu/Makefile)."""
    src = os.path.join(ROOT, "eicos_b200", "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(ROOT, "include", "eicos_b200.h")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    from eicos_b200.binding import Library
    return Library(EMU_LIB)


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; the gpu tests must exercise the CUDA path, so a missing .so or a
    missing device is a failure, not a skip."""
    import eicos_b200
    lib = eicos_b200.load()
    assert lib.L.eicos_device_count() >= 1, "no CUDA device visible"
    return lib


def relerr(a, b):
    import numpy as np
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.size == 0:
        return 0.0
    if a.ndim == 1:
        return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))
    return float(np.max(np.max(np.abs(a - b), axis=1) / np.maximum(1.0, np.max(np.abs(b), axis=1))))
