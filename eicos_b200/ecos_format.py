"""Reader / writer for the ECOS test-header fixture format the reference's tests are written in
(reference test/**/*.h: `static pfloat name[k] = {...};` / `static idxint name[k] = {...};` arrays
and scalar definitions, handed to `ECOS_setup`, see test/feasibilityProblems/feas.h:4-30).  Host-side
numpy only.  `load_problem` turns such a header into the problem dict every entry point of this
package takes, so that a new fixture in the reference's format can be dropped in; `write_header`
goes the other way and emits a header in the same style (data arrays plus a `test_<name>()` function
using the ECOS_* shim names), which compiles against the reference's test/ecos.h unchanged.
"""
import re

import numpy as np

_ARR = re.compile(r"(?:static\s+)?(idxint|pfloat)\s+(\w+)\s*\[\s*\d*\s*\]\s*=\s*\{([^}]*)\}\s*;", re.S)
_SCL = re.compile(r"(?:static\s+)?(idxint|pfloat)\s+(\w+)\s*=\s*([-+0-9.eE]+)\s*;")
# ECOS names for the CSC triplets
_ALIAS = {"Gx": "Gpr", "Gp": "Gjc", "Gi": "Gir", "Ax": "Apr", "Ap": "Ajc", "Ai": "Air"}
KEYS = ["n", "m", "p", "l", "ncones", "q", "c", "h", "b", "Gpr", "Gjc", "Gir", "Apr", "Ajc", "Air"]


def parse_header(path_or_text):
    """Every `idxint` / `pfloat` array and scalar definition of a header -> {name: ndarray | number}."""
    txt = path_or_text
    if "\n" not in txt and "{" not in txt:
        with open(path_or_text) as f:
            txt = f.read()
    out = {}
    for ty, name, body in _ARR.findall(txt):
        vals = [v for v in re.split(r"[,\s]+", body.strip()) if v]
        out[name] = np.array([float(v) for v in vals], dtype=np.float64)
        if ty == "idxint":
            out[name] = out[name].astype(np.int32)
    for ty, name, val in _SCL.findall(txt):
        out[name] = int(float(val)) if ty == "idxint" else float(val)
    return out


def select(raw, prefix="", suffix=""):
    """Pick `<prefix><key><suffix>` entries (ECOS aliases Gx/Gp/Gi/Ax/Ap/Ai accepted) into canonical keys."""
    d = {}
    for k in KEYS:
        for cand in [k] + [a for a, b in _ALIAS.items() if b == k]:
            for nm in (prefix + cand + suffix, prefix + cand):
                if nm in raw:
                    d[k] = raw[nm]
                    break
            if k in d:
                break
    return d


def finish(d, n=None, m=None, p=None, l=None, ncones=None):
    """Fill defaults, fix dtypes, derive missing dimensions from the arrays and check the CSC invariants."""
    d = dict(d)
    for k, v in dict(n=n, m=m, p=p, l=l, ncones=ncones).items():
        if v is not None:
            d[k] = v
    for k in ("q", "Gjc", "Gir", "Ajc", "Air"):
        d[k] = np.asarray(d.get(k, np.zeros(0)), dtype=np.int32)
    for k in ("c", "h", "b", "Gpr", "Apr"):
        d[k] = np.asarray(d.get(k, np.zeros(0)), dtype=np.float64)
    d.setdefault("n", d["c"].size)
    d.setdefault("m", d["h"].size)
    d.setdefault("p", d["b"].size)
    d.setdefault("ncones", d["q"].size)
    d.setdefault("l", int(d["m"]) - int(d["q"].sum()))
    for k in ("n", "m", "p", "l", "ncones"):
        d[k] = int(d[k])
    if not (d["c"].size == d["n"] and d["h"].size == d["m"] and d["b"].size == d["p"]):
        raise ValueError("vector sizes do not match n / m / p: %r" % ((d["c"].size, d["h"].size, d["b"].size, d["n"], d["m"], d["p"]),))
    for pr, jc, ir, rows in (("Gpr", "Gjc", "Gir", d["m"]), ("Apr", "Ajc", "Air", d["p"])):
        if d[pr].size:
            if not (d[jc].size == d["n"] + 1 and d[jc][-1] == d[pr].size == d[ir].size):
                raise ValueError(f"{pr}/{jc}/{ir}: not a CSC triplet with n+1 column pointers")
            if d[ir].min() < 0 or d[ir].max() >= rows:
                raise ValueError(f"{ir}: row index out of range")
    if d["l"] + int(d["q"].sum()) != d["m"]:
        raise ValueError("l + sum(q) != m")
    return d


def load_problem(path_or_text, prefix="", suffix="", **dims):
    """Header -> problem dict (keys KEYS).  prefix / suffix select one data set of the header
    (e.g. prefix="lp_afiro_", or prefix="udd_", suffix="1"); dims override scalars the header
    passes as literals to ECOS_setup (n=, m=, p=, l=, ncones=)."""
    return finish(select(parse_header(path_or_text), prefix, suffix), **dims)


def _carr(ty, name, a, per_line=8):
    a = np.asarray(a)
    if a.size == 0:
        return ""
    fmt = (lambda v: repr(float(v))) if ty == "pfloat" else (lambda v: str(int(v)))
    vals = [fmt(v) for v in a]
    lines = [", ".join(vals[i:i + per_line]) for i in range(0, len(vals), per_line)]
    return f"static {ty} {name}[{a.size}] = {{\n    " + ",\n    ".join(lines) + "};\n"


def write_header(P, name, expect="ECOS_OPTIMAL"):
    """Problem dict -> text of a test header in the reference's fixture style: data arrays `<name>_*`
    and `static char *test_<name>()` that sets up, solves, cleans up and asserts the exit flag."""
    P = finish(P)
    pre = name + "_"
    s = ['#include "ecos.h"', '#include "minunit.h"', ""]
    for ty, key in (("idxint", "q"), ("pfloat", "Gpr"), ("idxint", "Gjc"), ("idxint", "Gir"), ("pfloat", "Apr"),
                    ("idxint", "Ajc"), ("idxint", "Air"), ("pfloat", "c"), ("pfloat", "h"), ("pfloat", "b")):
        s.append(_carr(ty, pre + key, P[key]))
    ptr = lambda key: (pre + key) if np.asarray(P[key]).size else "NULL"
    hasG, hasA = P["Gpr"].size > 0, P["Apr"].size > 0
    s.append(f"""static char *test_{name}()
{{
    pwork *mywork;
    idxint exitflag = ECOS_FATAL;
    mywork = ECOS_setup({P['n']}, {P['m']}, {P['p']}, {P['l']}, {P['ncones']}, {ptr('q')}, 0,
                        {ptr('Gpr') if hasG else 'NULL'}, {ptr('Gjc') if hasG else 'NULL'}, {ptr('Gir') if hasG else 'NULL'},
                        {ptr('Apr') if hasA else 'NULL'}, {ptr('Ajc') if hasA else 'NULL'}, {ptr('Air') if hasA else 'NULL'},
                        {ptr('c')}, {ptr('h')}, {ptr('b')});
    if (mywork != NULL)
        exitflag = ECOS_solve(mywork);
    ECOS_cleanup(mywork, 0);
    mu_assert("{name}: unexpected exit flag", exitflag == {expect});
    return 0;
}}
""")
    return "\n".join(x for x in s if x is not None)
