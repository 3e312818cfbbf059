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
import re
import sys

import numpy as np

REF = os.environ.get("EICOS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")

ARR = re.compile(r"(?:static\s+)?(idxint|pfloat)\s+(\w+)\s*\[\s*\d*\s*\]\s*=\s*\{([^}]*)\}\s*;", re.S)
SCL = re.compile(r"(?:static\s+)?(idxint|pfloat)\s+(\w+)\s*=\s*([-+0-9.eE]+)\s*;")

ALIAS = {"Gx": "Gpr", "Gp": "Gjc", "Gi": "Gir", "Ax": "Apr", "Ap": "Ajc", "Ai": "Air"}
KEYS = ["n", "m", "p", "l", "ncones", "q", "c", "h", "b", "Gpr", "Gjc", "Gir", "Apr", "Ajc", "Air"]


def parse(path):
    txt = open(path).read()
    out = {}
    for ty, name, body in ARR.findall(txt):
        vals = [v for v in re.split(r"[,\s]+", body.strip()) if v]
        out[name] = np.array([float(v) for v in vals], dtype=np.float64 if ty == "pfloat" else np.int32)
        if ty == "idxint":
            out[name] = out[name].astype(np.int32)
    for ty, name, val in SCL.findall(txt):
        out[name] = int(val) if ty == "idxint" else float(val)
    return out


def canon(raw, prefix, suffix=""):
    """Pick `<prefix><key><suffix>` entries (with ECOS aliases) into canonical keys."""
    d = {}
    for k in KEYS:
        for cand in [k] + [a for a, b in ALIAS.items() if b == k]:
            for nm in (prefix + cand + suffix, prefix + cand):
                if nm in raw:
                    d[k] = raw[nm]
                    break
            if k in d:
                break
    return d


def finish(d, n=None, m=None, p=None, l=None, ncones=None):
    for k, v in dict(n=n, m=m, p=p, l=l, ncones=ncones).items():
        if v is not None:
            d[k] = v
    d.setdefault("p", 0)
    d.setdefault("ncones", 0)
    for k in ("q", "Gjc", "Gir", "Ajc", "Air"):
        d[k] = np.asarray(d.get(k, np.zeros(0)), dtype=np.int32)
    for k in ("c", "h", "b", "Gpr", "Apr"):
        d[k] = np.asarray(d.get(k, np.zeros(0)), dtype=np.float64)
    for k in ("n", "m", "p", "l", "ncones"):
        d[k] = np.int32(d[k])
    assert d["c"].size == d["n"] and d["h"].size == d["m"] and d["b"].size == d["p"], (d["n"], d["m"], d["p"])
    if d["Gpr"].size:
        assert d["Gjc"].size == d["n"] + 1 and d["Gjc"][-1] == d["Gpr"].size == d["Gir"].size
    if d["Apr"].size:
        assert d["Ajc"].size == d["n"] + 1 and d["Ajc"][-1] == d["Apr"].size == d["Air"].size
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
