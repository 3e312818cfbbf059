"""Host symbolic analysis of the product (eicos_b200/csrc/symbolic.cpp, amd.cpp) against the oracle's
independent restatement: ordering, elimination tree, pattern of L and KKT pattern must be bit-equal."""
import numpy as np
import pytest

from conftest import FIXTURES


def _compare(oracle_mod, emu_lib, P):
    from eicos_b200.binding import BatchSolver
    O = oracle_mod.OracleSolver(P)
    assert O.factor_init() == 0
    so = O.symbolic()
    B = BatchSolver(P, lib=emu_lib, capacity=1)
    sb = B.symbolic()
    d = B.dims()
    assert (d["dim_K"], d["nnzK"], d["nnzL"]) == O.dims()
    for k in ("Kp", "Ki", "pinv", "parent", "Lp", "Li"):
        assert np.array_equal(so[k], sb[k]), k
    assert sorted(sb["pinv"].tolist()) == list(range(d["dim_K"]))
    return d


@pytest.mark.parametrize("name", FIXTURES)
def test_symbolic_matches_oracle(oracle_mod, emu_lib, name):
    P = oracle_mod.load_fixture(name)
    d = _compare(oracle_mod, emu_lib, P)
    n, p, m = P["n"], P["p"], P["m"]
    nc = int(np.asarray(P["q"]).size) if m else 0
    assert d["dim_K"] == n + (p if np.asarray(P["Apr"]).size else 0) + m + 2 * nc  # src/eicos.cpp:165


def _random_socp(rng, n, p, l, cones):
    from eicos_b200.workloads import _csc
    m = l + sum(cones)
    G = np.where(rng.random((m, n)) < 3.0 / n, rng.standard_normal((m, n)), 0.0)
    G[np.arange(m), rng.integers(0, n, m)] = rng.standard_normal(m) + 2.0
    A = np.where(rng.random((p, n)) < 3.0 / n, rng.standard_normal((p, n)), 0.0)
    A[np.arange(p), rng.permutation(n)[:p]] = 1.0
    x0 = rng.standard_normal(n)
    s0 = np.abs(rng.standard_normal(m)) + 0.5
    z0 = np.abs(rng.standard_normal(m)) + 0.5
    at = l
    for d in cones:
        s0[at] = np.linalg.norm(s0[at + 1:at + d]) + 1.0
        z0[at] = np.linalg.norm(z0[at + 1:at + d]) + 1.0
        at += d
    y0 = rng.standard_normal(p)
    Gpr, Gjc, Gir = _csc(G)
    Apr, Ajc, Air = _csc(A)
    return dict(n=n, m=m, p=p, l=l, ncones=len(cones), q=np.array(cones, np.int32),
                Gpr=Gpr, Gjc=Gjc, Gir=Gir, Apr=Apr, Ajc=Ajc, Air=Air,
                c=-(G.T @ z0) - A.T @ y0, h=G @ x0 + s0, b=A @ x0)


@pytest.mark.parametrize("seed", range(6))
def test_symbolic_random_patterns(oracle_mod, emu_lib, seed):
    rng = np.random.default_rng(seed)
    cones = [int(v) for v in rng.integers(3, 9, size=int(rng.integers(1, 8)))]
    P = _random_socp(rng, n=int(rng.integers(20, 120)), p=int(rng.integers(1, 15)), l=int(rng.integers(0, 40)), cones=cones)
    _compare(oracle_mod, emu_lib, P)


def test_dense_row_goes_last(oracle_mod, emu_lib):
    """MPC02's 997-entry row of A exceeds AMD's dense threshold max(16, 10 sqrt(N)) and is ordered last."""
    from eicos_b200.binding import BatchSolver
    P = oracle_mod.load_fixture("MPC02")
    B = BatchSolver(P, lib=emu_lib, capacity=1)
    pinv = B.symbolic()["pinv"]
    assert pinv[-1] == P["n"] + 498
