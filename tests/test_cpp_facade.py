"""include/eicos.hpp (the C++ facade a user of the reference switches to): EiCOS::Solver and
EiCOS::BatchSolver driven from C++ (tests/cpp/facade_driver.cpp) the way the reference's own tests
drive the solver - setup, solve, updateData, solve - and compared with the CPU oracle.  The CPU tier
links the driver against the kernel emulator (tests/emu); the gpu tier against the CUDA library."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, relerr

TOL = 1e-7
CPP = os.path.join(ROOT, "tests", "cpp")


def _build(target):
    subprocess.check_call(["make", "-C", CPP, "-s", f"_build/{target}"])
    return os.path.join(CPP, "_build", target)


def _write_script(path, P, updates=(), batch=0, stacks=None):
    """Binary script read by facade_driver.cpp (little-endian int32 / float64)."""
    f64 = lambda a: np.ascontiguousarray(a, np.float64).tobytes()
    i32 = lambda a: np.ascontiguousarray(a, np.int32).tobytes()
    n, m, p = int(P["n"]), int(P["m"]), int(P["p"])
    Gpr, Apr = np.asarray(P["Gpr"], float), np.asarray(P["Apr"], float)
    q = np.asarray(P["q"], np.int32) if P.get("q") is not None else np.zeros(0, np.int32)
    jc = lambda a: np.asarray(a, np.int32) if np.asarray(a).size == n + 1 else np.zeros(n + 1, np.int32)
    Gjc, Ajc = jc(P["Gjc"]), jc(P["Ajc"])
    with open(path, "wb") as f:
        f.write(struct.pack("<7i", n, m, p, int(P.get("l", 0)), q.size, Gpr.size, Apr.size))
        f.write(i32(q) + i32(Gjc) + i32(P["Gir"]) + i32(Ajc) + i32(P["Air"]))
        f.write(f64(Gpr) + f64(Apr) + f64(P["c"]) + f64(P["h"]) + f64(P["b"]))
        f.write(struct.pack("<i", len(updates)))
        for full, arrs in updates:  # arrs: (Gpr, Apr, c, h, b), None = absent
            mask = sum(1 << k for k, a in enumerate(arrs) if a is not None) | (32 if full else 0)
            f.write(struct.pack("<i", mask))
            for a in arrs:
                if a is not None:
                    f.write(f64(a))
        f.write(struct.pack("<i", batch))
        if batch:
            keys = ("Gs", "As", "cs", "hs", "bs")
            mask = sum(1 << k for k, key in enumerate(keys) if stacks.get(key) is not None)
            f.write(struct.pack("<i", mask))
            for key in keys:
                if stacks.get(key) is not None:
                    f.write(f64(stacks[key]))


def _run(binary, script):
    out = subprocess.run([binary, script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    solves, cur = [], None
    for line in out.stdout.splitlines():
        t = line.split()
        if t[0] in ("solve", "batch"):
            cur = {"kind": t[0], "exit": int(t[1]), "iter": int(t[2]), "pcost": float(t[3])}
            solves.append(cur)
        elif t[0] in "xyzs":
            cur[t[0]] = np.array([float(v) for v in t[1:]])
        elif t[0] in ("error-check", "multi-device"):
            solves.append({"kind": t[0], "result": t[1]})
    return solves


def _check_sequence(oracle_mod, binary, tmp_path):
    """The reference's updateData test (test/ecostester.cpp, update_data fixtures): solve, update with
    the second data set through both overloads, back again, then the pointer-overload quirk."""
    P1, P2 = oracle_mod.load_fixture("update_data_1"), oracle_mod.load_fixture("update_data_2")
    ups = []
    for full in (False, True):
        for Pn in (P2, P1):
            ups.append((full, (Pn["Gpr"], Pn["Apr"], Pn["c"], Pn["h"], Pn["b"])))
    ups.append((False, (None, None, P1["c"] * 1.5, P1["h"] * 3, None)))  # h without Gpr is ignored, c honoured
    from eicos_b200.workloads import perturbed, perturbed_matrices
    B = 9
    W, M = perturbed(P1, B, rel=0.03, seed=6), perturbed_matrices(P1, B, rel=0.02, seed=8)
    stacks = {"Gs": M["Gs"], "As": M["As"], "hs": W["hs"], "bs": W["bs"]}
    script = str(tmp_path / "script.bin")
    _write_script(script, P1, ups, B, stacks)
    got = _run(binary, script)
    O = oracle_mod.OracleSolver(P1)
    want = [(O.solve(), O.info(), O.solution())]
    for full, arrs in ups:
        O.update_data(*arrs, full=full)
        want.append((O.solve(), O.info(), O.solution()))
    singles = [g for g in got if g["kind"] == "solve"]
    assert len(singles) == len(want)
    for g, (code, info, (x, y, z, s)) in zip(singles, want):
        assert g["exit"] == code and g["iter"] == info["iter"]
        for k, ref in zip("xyzs", (x, y, z, s)):
            assert relerr(g[k], ref) <= TOL, k
        assert abs(g["pcost"] - info["pcost"]) <= TOL * max(1.0, abs(info["pcost"]))
    ref = oracle_mod.batch_run(P1, B, Gs=M["Gs"], As=M["As"], hs=W["hs"], bs=W["bs"], nthreads=2)
    batched = [g for g in got if g["kind"] == "batch"]
    assert len(batched) == B
    for k, g in enumerate(batched):
        assert g["exit"] == ref["exit"][k] and g["iter"] == ref["iter"][k]
        for key in "xyzs":
            assert relerr(g[key], ref[key][k]) <= TOL, key
    assert {"kind": "multi-device", "result": "ok"} in got  # BatchSolver(..., devices): bit-identical to one handle
    assert got[-1] == {"kind": "error-check", "result": "ok"}  # malformed pattern -> exception, no fallback


def _check_fixture(oracle_mod, binary, tmp_path, name):
    P = oracle_mod.load_fixture(name)
    script = str(tmp_path / f"{name}.bin")
    _write_script(script, P)
    g = _run(binary, script)[0]
    O = oracle_mod.OracleSolver(P)
    code = O.solve()
    assert g["exit"] == code and g["iter"] == O.info()["iter"]
    if code == 0:
        for k, ref in zip("xyzs", O.solution()):
            assert relerr(g[k], ref) <= TOL, k


def test_facade_sequence_emulated(oracle_mod, emu_lib, tmp_path):
    _check_sequence(oracle_mod, _build("facade_emu"), tmp_path)


@pytest.mark.parametrize("name", ["lp_afiro", "feas", "infeasible1", "issue98", "emptyProblem"])
def test_facade_fixtures_emulated(oracle_mod, emu_lib, tmp_path, name):
    _check_fixture(oracle_mod, _build("facade_emu"), tmp_path, name)


@pytest.mark.gpu
def test_facade_sequence_gpu(oracle_mod, gpu_lib, tmp_path):
    _check_sequence(oracle_mod, _build("facade_gpu"), tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lp_afiro", "infeasible1", "MPC02"])
def test_facade_fixtures_gpu(oracle_mod, gpu_lib, tmp_path, name):
    _check_fixture(oracle_mod, _build("facade_gpu"), tmp_path, name)


def test_eigen_typed_overloads_compile_and_run(emu_lib):
    """The Eigen-typed constructor / updateData / solution() of include/eicos.hpp (reference include/eicos.hpp:138-148)
    against tests/cpp/mock_eigen (Eigen itself is not in the image): a 2-variable problem with one LP row and one
    second-order cone, then updateData with the Eigen signature."""
    out = subprocess.run([_build("eigen_overloads_emu")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.splitlines()]
    assert [r[1] for r in rows] == ["0", "0"]
    assert np.allclose([float(v) for v in rows[0][3:]], [1.0, 2.0], atol=1e-6)
    assert np.allclose([float(v) for v in rows[1][3:]], [3.0, 5.0], atol=1e-6)
