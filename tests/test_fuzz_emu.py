"""Randomised parity (tools/fuzz_emu.py): random small SOCPs with irregular sparsity, random slot budgets
and worker counts, shared and per-instance matrices, kernel emulator against the oracle.  A bounded
slice of the sweep (the full one - 1500 problems, 9000 solves - is recorded in DESIGN.md section 6)."""
import os
import sys

from conftest import ROOT


def test_random_socps_match_the_oracle(oracle_mod, emu_lib):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_emu
    bad, edges, tally = fuzz_emu.run(40, first=2000, verbose=False)
    assert not bad, bad
    assert len(edges) <= 1, edges          # exactly-zero-pivot coin flips, see fuzz_emu.run
    assert tally.get(0, 0) >= 150          # the sweep really solves problems (optimal), it does not just agree on failures
