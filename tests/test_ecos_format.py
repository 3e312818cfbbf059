"""eicos_b200/ecos_format.py: fixtures in the reference's test-header format can be
read into problem dicts and written back; a written header compiles against the ECOS-name shim
(tests/cpp/shim/ecos.h, bound to the C ABI) and runs on the engine."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, EMU_LIB


@pytest.mark.parametrize("name", ["lp_afiro", "issue98", "update_data_1", "emptyProblem", "feas", "infeasible1"])
def test_header_round_trip(oracle_mod, name):
    from eicos_b200 import ecos_format as ef
    P = oracle_mod.load_fixture(name)
    Q = ef.load_problem(ef.write_header(P, name), prefix=name + "_", n=int(P["n"]), m=int(P["m"]), p=int(P["p"]))
    for k in ef.KEYS:
        a, b = np.asarray(P[k]), np.asarray(Q[k])
        assert a.size == b.size and np.array_equal(a.ravel(), b.ravel()), k


def test_reads_the_reference_spelling():
    """Array names with the ECOS aliases (Gx/Gp/Gi) and dimensions passed as literals, as in
    the reference's test/feasibilityProblems/feas.h (a two-row LP written here by hand)."""
    from eicos_b200 import ecos_format as ef
    txt = """
    static pfloat box_Gx[2] = {1, -1};
    static idxint box_Gp[2] = {0, 2};
    static idxint box_Gi[2] = {0, 1};
    static pfloat box_c[1] = {0};
    static pfloat box_h[2] = {1, 0};
    """
    P = ef.load_problem(txt, prefix="box_")
    assert (P["n"], P["m"], P["p"], P["l"], P["ncones"]) == (1, 2, 0, 2, 0)
    assert P["Gpr"].tolist() == [1.0, -1.0] and P["Gjc"].tolist() == [0, 2] and P["Gir"].dtype == np.int32
    with pytest.raises(ValueError):
        ef.load_problem(txt.replace("{0, 1}", "{0, 5}"), prefix="box_")  # row index out of range


def test_written_header_runs_on_the_engine(oracle_mod, emu_lib, tmp_path):
    from eicos_b200 import ecos_format as ef
    names = {"lp_afiro": "ECOS_OPTIMAL", "issue98": "ECOS_OPTIMAL", "infeasible1": "ECOS_PINF", "unboundedLP1": "ECOS_DINF"}
    main = ['#include "ecos.h"', '#include "minunit.h"']
    for name, expect in names.items():
        (tmp_path / f"{name}.h").write_text(ef.write_header(oracle_mod.load_fixture(name), name, expect))
        main.append(f'#include "{name}.h"')
    main.append("int main() {")
    main += [f'    test_{n}(); std::printf("PASS {n}\\n");' for n in names]
    main.append("    return 0;\n}")
    (tmp_path / "main.cpp").write_text("\n".join(main))
    exe = str(tmp_path / "tester")
    emu_dir = os.path.dirname(EMU_LIB)
    subprocess.check_call(["g++", "-O0", "-std=c++17", "-w", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "cpp", "shim"), "-I", str(tmp_path),
                           "-o", exe, str(tmp_path / "main.cpp"), "-L", emu_dir, "-leicos_emu", f"-Wl,-rpath,{emu_dir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split("\n")[:4] == [f"PASS {n}" for n in names]
