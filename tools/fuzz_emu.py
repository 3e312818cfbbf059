#!/usr/bin/env python3
"""Randomised parity sweep on the CPU: random small feasible SOCPs with irregular sparsity (cones of
dimension 1..7, optional equalities, no LP rows at all in some cases, a dense row or column now and
then) through the kernel emulator (tests/emu) against the oracle - shared-matrices and
per-instance-matrices handles, random slot budgets for the program compiler.  Test infrastructure.

    python tools/fuzz_emu.py [cases] [first_seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_mpc(seed):
    """Banded MPC-like LP (scalar or 2-state dynamics, box bounds, optional input-norm cones): narrow
    columns of L, so every value of the factor program stays in its slots."""
    from eicos_b200.workloads import _csc
    rng = np.random.default_rng(seed)
    nx, T = int(rng.integers(1, 3)), int(rng.integers(3, 25))
    nv = (nx + 1) * T  # per stage: u_k, x_{k+1}
    Ad = np.eye(nx) + 0.1 * np.triu(np.ones((nx, nx)), 1)
    Bd = np.full(nx, 0.1)
    A = np.zeros((nx * T, nv))
    b = np.zeros(nx * T)
    x0 = rng.uniform(-1, 1, nx)
    for k in range(T):
        r = slice(nx * k, nx * (k + 1))
        A[r, (nx + 1) * k] = -Bd
        A[r, (nx + 1) * k + 1:(nx + 1) * (k + 1)] = np.eye(nx)
        if k:
            A[r, (nx + 1) * (k - 1) + 1:(nx + 1) * k] = -Ad
        else:
            b[r] = Ad @ x0
    G = np.vstack([np.eye(nv), -np.eye(nv)])
    h = np.concatenate([np.full(nv, 3.0 + rng.random()), np.full(nv, 3.0 + rng.random())])
    c = rng.standard_normal(nv)
    Gpr, Gjc, Gir = _csc(G)
    Apr, Ajc, Air = _csc(A)
    return dict(n=nv, m=2 * nv, p=nx * T, l=2 * nv, ncones=0, q=np.zeros(0, np.int32), Gpr=Gpr, Gjc=Gjc, Gir=Gir,
                Apr=Apr, Ajc=Ajc, Air=Air, c=c, h=h, b=b)


def random_problem(seed):
    if seed % 7 == 3:
        return random_mpc(seed)
    rng = np.random.default_rng(seed)
    big = seed % 5 == 0  # every fifth problem is larger: longer programs, far gathers, FIFO wrap-arounds
    n = int(rng.integers(40, 160)) if big else int(rng.integers(1, 40))
    p = int(rng.integers(0, max(1, n // 2) + 1)) if rng.random() < 0.7 else 0
    ncones = int(rng.integers(0, 24 if big else 6)) if rng.random() < 0.7 else 0
    q = rng.integers(1, 8, size=ncones).astype(np.int32)
    l = int(rng.integers(0, 2 * n + 2)) if (ncones == 0 or rng.random() < 0.8) else 0
    if l + int(q.sum()) == 0:
        l = n + 1
    m = l + int(q.sum())
    dens = rng.choice([0.02, 0.05, 0.1]) if big else rng.choice([0.08, 0.2, 0.5])

    def sparse(rows):
        M = (rng.random((rows, n)) < dens) * rng.standard_normal((rows, n))
        if rows and rng.random() < 0.3:
            M[int(rng.integers(0, rows)), :] = rng.standard_normal(n)  # a dense row
        if rows and rng.random() < 0.3:
            M[:, int(rng.integers(0, n))] = rng.standard_normal(rows)  # a dense column
        return M
    G, A = sparse(m), sparse(p)
    for r in np.nonzero(~G.any(axis=1))[0]:  # every inequality row mentions a variable
        G[r, int(rng.integers(0, n))] = rng.standard_normal() + 2.0
    # bounded feasible set: box rows make sure every variable is constrained
    if l >= 2 * n:
        G[:n, :] = np.eye(n) * (1 + rng.random(n))[:, None] * np.eye(n)
        G[n:2 * n, :] = -np.eye(n)
    if p:  # full row rank
        for i in range(p):
            A[i, i % n] += 2.0 + rng.random()
    from eicos_b200.workloads import _csc
    x0, y0 = rng.standard_normal(n), rng.standard_normal(p)
    s0, z0 = np.abs(rng.standard_normal(m)) + 0.5, np.abs(rng.standard_normal(m)) + 0.5
    at = l
    for d in q:
        s0[at] = np.linalg.norm(s0[at + 1:at + d]) + 1.0
        z0[at] = np.linalg.norm(z0[at + 1:at + d]) + 1.0
        at += d
    Gpr, Gjc, Gir = _csc(G)
    Apr, Ajc, Air = _csc(A)
    return dict(n=n, m=m, p=p, l=l, ncones=ncones, q=q, Gpr=Gpr, Gjc=Gjc, Gir=Gir, Apr=Apr, Ajc=Ajc, Air=Air,
                c=-G.T @ z0 - A.T @ y0, h=G @ x0 + s0, b=A @ x0)


def run(cases, first=0, verbose=True):
    """Returns (mismatches, zero_pivot_edges, tally): runs whose exit flags / iteration counts / solutions
    differ; runs that differ only because ONE side stopped with FATAL (an exactly-zero pivot - both
    implementations do, rarely and on different instances: once eta^2 of a cone exceeds 2^53 delta the static
    regularisation is rounded away and the summation order decides whether the cancellation is exact);
    oracle exit-flag counts."""
    import oracle
    from conftest import EMU_LIB, relerr
    from eicos_b200.binding import BatchSolver, Library
    from eicos_b200.workloads import perturbed, perturbed_matrices
    lib = Library(EMU_LIB)
    bad, edges, tally = [], [], {}
    saved = {k: os.environ.get(k) for k in ("EICOS_MAX_SW_SLOTS", "EICOS_MAX_FA_SLOTS")}
    for seed in range(first, first + cases):
        P = random_problem(seed)
        rng = np.random.default_rng(10_000 + seed)
        os.environ["EICOS_MAX_SW_SLOTS"] = str(int(rng.integers(1, 24)))
        os.environ["EICOS_MAX_FA_SLOTS"] = str(int(rng.integers(1, 40)))
        B = 3
        W = perturbed(P, B, rel=0.02, seed=seed)
        M = perturbed_matrices(P, B, rel=0.01, seed=seed + 1, scale_spread=0.5 if seed % 2 else 0.0)
        for mode in ("shared", "pim"):
            tag = f"seed {seed} {mode} n={P['n']} m={P['m']} p={P['p']} l={P['l']} q={P['q'].tolist()}"
            try:
                if mode == "shared":
                    ref = oracle.batch_run(P, B, hs=W["hs"], bs=W["bs"], nthreads=1)
                    out = BatchSolver(P, lib=lib, capacity=2, workers=1 + seed % 4).solve(B, hs=W["hs"], bs=W["bs"])
                else:
                    hs = M["hs"] if M["hs"] is not None else W["hs"]
                    bs = M["bs"] if M["bs"] is not None else W["bs"]
                    ref = oracle.batch_run(P, B, Gs=M["Gs"], As=M["As"], hs=hs, bs=bs, nthreads=1)
                    out = BatchSolver(P, lib=lib, capacity=2, workers=1 + seed % 4, instance_matrices=True).solve(
                        B, hs=hs, bs=bs, Gs=M["Gs"], As=M["As"])
            except Exception as e:  # noqa: BLE001
                bad.append(f"{tag}: EXCEPTION {e}")
                continue
            for e in ref["exit"]:
                tally[int(e)] = tally.get(int(e), 0) + 1
            fatal = (ref["exit"] == -7) | (out["exit"] == -7)
            same = ~fatal
            msg = []
            if not np.array_equal(out["exit"][same], ref["exit"][same]):
                msg.append(f"exit {out['exit']} vs {ref['exit']}")
            if not np.array_equal(out["iter"][same], ref["iter"][same]):
                msg.append(f"iter {out['iter']} vs {ref['iter']}")
            ok = (ref["exit"] == 0) & (out["exit"] == 0)
            if ok.any():
                errs = {k: relerr(out[k][ok], ref[k][ok]) for k in "xyzs"}
                if P["n"] > P["m"] + P["p"]:
                    errs.pop("x")  # more variables than rows: the optimal x is not unique
                if max(errs.values()) > 1e-7:
                    msg.append("err " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
            if msg:
                bad.append(f"{tag}: " + "; ".join(msg))
            elif fatal.any() and not np.array_equal(out["exit"], ref["exit"]):
                edges.append(f"{tag}: exit {out['exit']} vs {ref['exit']}")
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    if verbose:
        for line in bad:
            print("MISMATCH", line)
        for line in edges:
            print("zero-pivot edge", line)
        print(f"{cases} problems x 2 modes x 3 instances: {len(bad)} mismatching runs, {len(edges)} zero-pivot edges; "
              f"oracle exit flags {tally}")
    return bad, edges, tally


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad, _, _ = run(cases, first)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
