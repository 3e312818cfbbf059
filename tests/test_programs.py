"""The host-side program compiler (eicos_b200/csrc/streams.cpp): slot programs for the factorisation,
the triangular sweeps and the KKT mat-vecs.  Run through the CPU emulator, whose cp.async is an
immediate copy - so a load that the compiler schedules before its row has been produced reads stale
data and shows up as a parity failure against the oracle."""
import os

import numpy as np
import pytest

from conftest import relerr

TOL = 1e-7


def _budget(monkeypatch, sw, fa):
    monkeypatch.setenv("EICOS_MAX_SW_SLOTS", str(sw))
    monkeypatch.setenv("EICOS_MAX_FA_SLOTS", str(fa))


def test_mpc02_program_is_slot_resident(oracle_mod, emu_lib):
    """The benchmark pattern: every value of the factorisation and all but the dense row's operands of the
    sweeps live in slots; HBM reads are the algorithmic minimum."""
    from eicos_b200.binding import BatchSolver
    P = oracle_mod.load_fixture("MPC02")
    B = BatchSolver(P, lib=emu_lib, capacity=1)
    ps, d = B.program_stats(), B.dims()
    assert ps["fa_fast"] == 1 and ps["fa_home"] == 0 and ps["sw_direct"] == 0
    assert ps["sw_slots"] <= 32 and ps["fa_slots"] <= 32  # (the deeper-ring variants get twice the 16 slots)
    N, nnzL, nnzV = d["dim_K"], d["nnzL"], d["nnzV"]
    # HBM reads per run = the algorithmic minimum plus the re-reads of values that lost their slot
    # (the operands of the one dense row); sw_far counts the forward and the plain backward sweep
    assert ps["fa_loads"] == nnzV
    assert N + nnzL <= ps["fw_loads"] <= N + nnzL + ps["sw_far"]
    assert 3 * N + nnzL <= ps["bw_loads"] <= 3 * N + nnzL + ps["sw_far"]
    # (rows of L without entries are read from the right-hand side wherever they are used - build_forward - which
    #  trades the 3 497 copies of MPC02's forward sweep for ~5 000 re-reads)
    assert ps["sw_far"] <= 0.6 * nnzL


@pytest.mark.parametrize("name,sw,fa", [("update_data_1", 3, 2), ("update_data_1", 2, 2), ("lp_afiro", 3, 4),
                                        ("issue98", 2, 2), ("lp_blend", 4, 6), ("unboundedLP1", 2, 2)])
def test_parity_with_starved_slots(oracle_mod, emu_lib, monkeypatch, name, sw, fa):
    """Tiny slot budgets force evictions, re-reads of home rows through the ring (with the padding pops that
    keep them behind their writers), partial sums parked in their home rows and the general-form factor
    with home-row accumulators; results must not change.  (Two slots is the machine's minimum: one is held
    by tau in the computeResiduals program.)"""
    from eicos_b200.binding import BatchSolver, Solver
    _budget(monkeypatch, sw, fa)
    P = oracle_mod.load_fixture(name)
    B = BatchSolver(P, lib=emu_lib, capacity=1)
    ps = B.program_stats()
    assert ps["sw_slots"] <= sw and ps["fa_slots"] <= fa
    assert ps["sw_far"] + ps["sw_direct"] > 0 and ps["fa_home"] > 0
    O = oracle_mod.OracleSolver(P)
    co = O.solve()
    S = Solver(P, lib=emu_lib)
    assert S.solve() == co
    io, ie = O.info(), S.info()
    assert ie["iter"] == io["iter"] and ie["nitref1"] == io["nitref1"] and ie["nitref2"] == io["nitref2"]
    if co == 0:
        xo, yo, zo, so = O.solution()
        ye, ze, se = S.duals()
        for a, b in ((S.solution(), xo), (ye, yo), (ze, zo), (se, so)):
            assert relerr(a, b) <= TOL


def test_slot_budget_does_not_change_bits(oracle_mod, emu_lib, monkeypatch):
    """Slots only decide WHERE a value waits, not what is computed: starved and roomy programs give
    bit-identical solutions."""
    from eicos_b200.binding import BatchSolver
    from eicos_b200.workloads import perturbed
    P = oracle_mod.load_fixture("update_data_1")
    W = perturbed(P, 6, rel=0.05, seed=3)
    outs = []
    for sw, fa in ((24, 20), (2, 3)):
        _budget(monkeypatch, sw, fa)
        B = BatchSolver(P, lib=emu_lib, capacity=6)
        outs.append(B.solve(6, hs=W["hs"], bs=W["bs"]))
    for k in ("x", "y", "z", "s"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert np.array_equal(outs[0]["iter"], outs[1]["iter"])


def test_catastrophic_fill_is_refused(oracle_mod, emu_lib):
    """configs[3] at its literal size fills L to 1.57 M entries (8.3e8 Schur updates per factorisation):
    setup must fail with a clear message instead of compiling a multi-gigabyte factor program."""
    from eicos_b200.binding import BatchSolver
    from eicos_b200.workloads import synthetic_socp
    P = synthetic_socp()
    with pytest.raises(RuntimeError, match="fills too much"):
        BatchSolver(P, lib=emu_lib, capacity=1)
