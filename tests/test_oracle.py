"""Pins the CPU oracle (oracle/) against everything the reference's own tests hold for this path
(exit flags), independent HiGHS objectives, KKT optimality certificates and
closed-form SOCPs.  Ordering / iterate parity with Eigen itself cannot be pinned here (no Eigen)."""
import json
import os

import numpy as np
import pytest

from conftest import FIXTURES, ROOT, relerr

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_objectives.json")))

# unboundedMaxSqrt (3-cone SOCP, test/unboundedProblems/unboundedMaxSqrt.h:33 expects DINF) is the one expectation the
# double-precision restatement does not reproduce; test_unbounded_maxsqrt_is_a_rounding_knife_edge below establishes
# why, and tests/test_reference_tester.py carries the one xfail for it.
ROUNDING_KNIFE_EDGE = {"unboundedMaxSqrt"}


@pytest.mark.parametrize("name", [n for n in FIXTURES if n not in ROUNDING_KNIFE_EDGE])
def test_exit_flags_match_reference_tests(oracle_mod, name):
    P = oracle_mod.load_fixture(name)
    S = oracle_mod.OracleSolver(P)
    code = S.solve()
    assert code in P["expect"], (name, code)
    if name in GOLD:
        x = S.solution()[0]
        obj = float(P["c"] @ x)
        assert abs(obj - GOLD[name]) <= 1e-7 * max(1.0, abs(GOLD[name])), (name, obj, GOLD[name])
        assert abs(S.info()["pcost"] * 1.0 - S.info()["pcost"]) == 0.0


def _trace_column(trace, key):
    import re
    return [float(m.group(1)) for m in re.finditer(key + r"=([-+0-9.e]+)", trace)]


def test_unbounded_maxsqrt_is_a_rounding_knife_edge(oracle_mod, capsys):
    """The reference expects DINF for unboundedMaxSqrt.  The same restatement in 80-bit extended precision (11 more
    mantissa bits, oracle/_build/libeicos_oracle_ld.so) ends with DINF: dinfres drops below feastol = 1e-8 and the
    dual-infeasibility test of checkExitConditions (src/eicos.cpp:573-596) fires.  In double precision the iterates
    have separated from that trajectory by iteration 11 (the problem is unbounded: the KKT systems lose all accuracy
    as tau -> 0), dinfres hovers just above feastol for five iterations and the run ends in the safeguard
    `pres > 500 pres_prev` (:1010) with NUMERICS.  The margin by which the double run misses DINF is printed."""
    P = oracle_mod.load_fixture("unboundedMaxSqrt")
    assert 2 in P["expect"]
    code_ld, trace_ld = oracle_mod.solve_extended_precision("unboundedMaxSqrt", trace=True)
    assert code_ld == 2, "extended precision reaches the reference's DINF"
    code_d, trace_d = oracle_mod.solve_trace("unboundedMaxSqrt")
    assert code_d in (2, -2)
    dinf_d, dinf_ld = _trace_column(trace_d, "dinfres"), _trace_column(trace_ld, "dinfres")
    pres_d, prev_d = _trace_column(trace_d, "pres"), _trace_column(trace_d, "pres_prev")
    feastol = 1e-8
    assert min(v for v in dinf_ld if v > 0) < feastol
    with capsys.disabled():
        print("\n  unboundedMaxSqrt: extended precision -> exit %d after %d iterations (min dinfres %.3e < feastol)"
              % (code_ld, len(dinf_ld) - 1, min(v for v in dinf_ld if v > 0)))
        print("  double precision   -> exit %d after %d iterations; closest approach dinfres = %.4e = %.2f x feastol; "
              "last safeguard ratio pres / pres_prev = %.3g (limit 500)"
              % (code_d, len(dinf_d) - 1, min(v for v in dinf_d if v > 0), min(v for v in dinf_d if v > 0) / feastol,
                 pres_d[-1] / prev_d[-1]))
    if code_d == -2:
        assert min(v for v in dinf_d if v > 0) < 2 * feastol  # a near miss, not a different answer


def test_oracle_chaotic_case(oracle_mod):
    """The one deviation is rounding chaos, not logic: other pivot orders reach the reference's DINF."""
    import subprocess
    import sys
    codes = []
    for order in ("natural", "reverse", "rot1", "rot2", "rot3", "rot4", "rot5"):
        env = dict(os.environ, ORA_ORDERING=order, PYTHONPATH=ROOT)
        out = subprocess.check_output([sys.executable, "-c",
                                       "import oracle; S=oracle.OracleSolver(oracle.load_fixture('unboundedMaxSqrt')); print(S.solve())"],
                                      env=env, cwd=ROOT)
        codes.append(int(out.split()[-1]))
    assert set(codes) <= {2, -2} and codes.count(2) >= 2, codes


def _kkt_certificate(P, x, y, z, s, tol=2e-6):
    import scipy.sparse as sp
    n, m, p = P["n"], P["m"], P["p"]
    G = sp.csc_matrix((P["Gpr"], P["Gir"], P["Gjc"]), shape=(m, n))
    sc = 1.0 + max(np.abs(x).max(initial=0), np.abs(z).max(initial=0))
    assert np.abs(G @ x + s - P["h"]).max() <= tol * sc
    r = P["c"] + G.T @ z
    if p:
        A = sp.csc_matrix((P["Apr"], P["Air"], P["Ajc"]), shape=(p, n))
        assert np.abs(A @ x - P["b"]).max() <= tol * sc
        r = r + A.T @ y
    assert np.abs(r).max() <= tol * sc
    l = m - int(np.sum(P["q"]))
    assert s[:l].min(initial=0) >= -tol and z[:l].min(initial=0) >= -tol
    at = l
    for d in P["q"]:
        for v in (s, z):
            assert v[at] - np.linalg.norm(v[at + 1:at + d]) >= -tol * sc
        at += d
    assert abs(float(s @ z)) <= tol * sc * sc


def test_socp_closed_form(oracle_mod):
    """min t s.t. ||x - a|| <= t, x_1 = 1  (builder's own case): t* = |1 - a_1|, x* = (1, a_2, a_3)."""
    a = np.array([3.0, -2.0, 0.5])
    # variables (x1,x2,x3,t); cone (t; x - a) in Q^4  ->  G = -[e_t; I], h = (0; -a)
    Gd = np.zeros((4, 4))
    Gd[0, 3] = -1
    Gd[1:, :3] = -np.eye(3)
    from eicos_b200.workloads import _csc
    Gpr, Gjc, Gir = _csc(Gd)
    Apr, Ajc, Air = _csc(np.array([[1.0, 0, 0, 0]]))
    P = dict(n=4, m=4, p=1, l=0, ncones=1, q=np.array([4], np.int32), Gpr=Gpr, Gjc=Gjc, Gir=Gir,
             Apr=Apr, Ajc=Ajc, Air=Air, c=np.array([0, 0, 0, 1.0]), h=np.concatenate([[0], -a]), b=np.array([1.0]))
    S = oracle_mod.OracleSolver(P)
    assert S.solve() == 0
    x, y, z, s = S.solution()
    assert np.allclose(x, [1.0, -2.0, 0.5, 2.0], atol=1e-7)
    _kkt_certificate(P, x, y, z, s)


@pytest.mark.parametrize("T", [5, 20])
def test_soc_mpc_optimality_certificate(oracle_mod, T):
    from eicos_b200.workloads import soc_mpc
    P = soc_mpc(T=T)
    S = oracle_mod.OracleSolver(P)
    assert S.solve() == 0
    _kkt_certificate(P, *S.solution())


@pytest.mark.parametrize("name", ["lp_afiro", "lp_blend", "issue98", "MPC02", "update_data_1"])
def test_fixture_optimality_certificate(oracle_mod, name):
    P = oracle_mod.load_fixture(name)
    S = oracle_mod.OracleSolver(P)
    assert S.solve() in (0, 10)
    _kkt_certificate(P, *S.solution(), tol=1e-5)


def test_update_data_sequence(oracle_mod):
    """reference test/updateData/update_data.h:1657-1688: solve, replace G,A,c,h,b, solve again."""
    P1, P2 = oracle_mod.load_fixture("update_data_1"), oracle_mod.load_fixture("update_data_2")
    S = oracle_mod.OracleSolver(P1)
    assert S.solve() in (0, 10)
    S.update_data(P2["Gpr"], P2["Apr"], P2["c"], P2["h"], P2["b"])
    assert S.solve() in (0, 10)
    x_upd = S.solution()[0]
    F = oracle_mod.OracleSolver(P2)
    F.solve()
    assert relerr(x_upd, F.solution()[0]) <= 1e-9
    assert abs(float(P2["c"] @ x_upd) - GOLD["update_data_2"]) <= 1e-6
    # Eigen overload (src/eicos.cpp:2032-2051)
    S.update_data(P1["Gpr"], P1["Apr"], P1["c"], P1["h"], P1["b"], full=True)
    assert S.solve() in (0, 10)
    assert abs(float(P1["c"] @ S.solution()[0]) - GOLD["update_data_1"]) <= 1e-6


def test_pointer_update_quirk_h_follows_G(oracle_mod):
    """src/eicos.cpp:2059-2070: h is only read when Gpr is given, b only when Apr is."""
    P = oracle_mod.load_fixture("update_data_1")
    S = oracle_mod.OracleSolver(P)
    S.solve()
    x0 = S.solution()[0]
    S.update_data(None, None, None, P["h"] * 2.0, P["b"] * 2.0)  # ignored: no Gpr / Apr
    S.solve()
    assert relerr(S.solution()[0], x0) <= 1e-9


def test_run_cpp_sequence_on_mpc02(oracle_mod):
    """BASELINE.json configs[0] (src/run.cpp:34-52): setup, solve, updateData(same data), solve,
    assert optimal - on MPC02 because data_MPC01.hpp is missing from the checkout."""
    P = oracle_mod.load_fixture("MPC02")
    S = oracle_mod.OracleSolver(P)
    assert S.solve() == 0
    it0 = S.info()["iter"]
    S.update_data(P["Gpr"], P["Apr"], P["c"], P["h"], P["b"], full=True)
    assert S.solve() == 0 and S.info()["iter"] == it0


def test_sticky_infeasibility_flags_quirk(oracle_mod):
    """pinfres is never cleared (src/eicos.cpp:720-728): a reused Solver differs from a fresh one."""
    from eicos_b200.workloads import perturbed
    P = oracle_mod.load_fixture("MPC02")
    W = perturbed(P, 6, rel=0.05)
    fresh = oracle_mod.batch_run(P, 6, hs=W["hs"], bs=W["bs"], nthreads=1, reset_sticky=True)
    reused = oracle_mod.batch_run(P, 6, hs=W["hs"], bs=W["bs"], nthreads=1, reset_sticky=False)
    assert 1 in fresh["exit"]  # some instance is primal infeasible ...
    first = int(np.argmax(fresh["exit"] == 1))
    assert np.array_equal(fresh["iter"][:first + 1], reused["iter"][:first + 1])
    assert not np.array_equal(fresh["iter"], reused["iter"])  # ... and poisons the solves after it
