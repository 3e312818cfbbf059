#!/usr/bin/env python3
"""Convert the reference's ECOS-format test headers into compact .npz fixtures.

Run in the build container (where /root/reference exists):

    python tests/golden/make_fixtures.py

The GPU box has no /root/reference, so the parsed problem data (inputs only -
the reference's tests pin nothing but the exit flag) is committed
under tests/golden/fixtures/*.npz together with this script.  Expected exit
flags are the ones asserted by the reference's own tests (test/**/*.h,
`mu_assert` lines); the HiGHS objectives in golden_objectives.json were
computed independently with scipy (see make_objectives.py).
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("EICOS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from eicos_b200.ecos_format import KEYS, parse_header as parse, select as canon, finish as _finish  # noqa: E402


def finish(d, **dims):
    """ecos_format.finish with the scalar dtypes the committed .npz files were written with."""
    d = _finish(d, **dims)
    for k in ("n", "m", "p", "l", "ncones"):
        d[k] = np.int32(d[k])
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    T = os.path.join(REF, "test")
    manifest = {}

    def save(name, d, expect, src):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        manifest[name] = {"expect": expect, "source": src,
                          "n": int(d["n"]), "m": int(d["m"]), "p": int(d["p"]),
                          "l": int(d["m"] - d["q"].sum()), "ncones": int(d["q"].size)}

    # LPnetlib x10: expect OPTIMAL (test/LPnetlib/lp_*.h, mu_assert at the end of each)
    for nm in ["25fv47", "adlittle", "afiro", "agg", "agg2", "agg3", "bandm", "beaconfd", "blend", "bnl1"]:
        raw = parse(os.path.join(T, "LPnetlib", f"lp_{nm}.h"))
        save(f"lp_{nm}", finish(canon(raw, f"lp_{nm}_")), [0], f"test/LPnetlib/lp_{nm}.h")

    raw = parse(os.path.join(T, "MPC", "MPC02.h"))
    save("MPC02", finish(canon(raw, "MPC02_")), [0, 10], "test/MPC/MPC02.h:38")

    raw = parse(os.path.join(T, "updateData", "update_data.h"))
    for s in ("1", "2"):
        d = {}
        for k in KEYS:
            for nm in (f"udd_{k}{s}", f"udd_{k[0]}{s}{k[1:]}", f"udd_{k}"):
                if nm in raw:
                    d[k] = raw[nm]
                    break
        save(f"update_data_{s}", finish(d), [0, 10], "test/updateData/update_data.h:1657-1688")

    raw = parse(os.path.join(T, "cvxpyProblems", "githubIssue98.h"))
    save("issue98", finish(canon(raw, ""), n=5, m=11, p=0, l=6, ncones=1), [0], "test/cvxpyProblems/githubIssue98.h:26-43")

    raw = parse(os.path.join(T, "feasibilityProblems", "feas.h"))
    save("feas", finish(canon(raw, "feas_"), n=1, m=2, p=0, l=2, ncones=0), [0], "test/feasibilityProblems/feas.h:21-36")

    for nm, rel, exp in [("unboundedLP1", "unboundedProblems/unboundedLP1.h", [2]),
                         ("unboundedMaxSqrt", "unboundedProblems/unboundedMaxSqrt.h", [2]),
                         ("infeasible1", "infeasibleProblems/infeasible1.h", [1]),
                         ("infeasible2", "infeasibleProblems/infeasible2.h", [1]),
                         ("emptyProblem", "emptyProblem/emptyProblem.h", [0])]:
        raw = parse(os.path.join(T, rel))
        save(nm, finish(canon(raw, "")), exp, "test/" + rel)

    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)
    for k, v in sorted(manifest.items()):
        print(k, v)


if __name__ == "__main__":
    sys.exit(main())
