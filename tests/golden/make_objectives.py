#!/usr/bin/env python3
"""Recompute golden_objectives.json with HiGHS (scipy) from the committed fixtures."""
import json
import os

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
man = json.load(open(os.path.join(HERE, "fixtures", "manifest.json")))
for name, meta in sorted(man.items()):
    if meta["ncones"] or meta["expect"] not in ([0], [0, 10]) or meta["n"] == 0:
        continue
    d = np.load(os.path.join(HERE, "fixtures", name + ".npz"))
    n, m, p = int(d["n"]), int(d["m"]), int(d["p"])
    G = sp.csc_matrix((d["Gpr"], d["Gir"], d["Gjc"]), shape=(m, n))
    A = sp.csc_matrix((d["Apr"], d["Air"], d["Ajc"]), shape=(p, n)) if p else None
    r = linprog(d["c"], A_ub=G, b_ub=d["h"], A_eq=A, b_eq=d["b"] if p else None, bounds=(None, None), method="highs")
    out[name] = float(r.fun)
    print(name, r.status, r.fun)
json.dump(out, open(os.path.join(HERE, "golden_objectives.recomputed.json"), "w"), indent=1)
