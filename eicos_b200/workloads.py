"""Synthetic workloads for the parity tests and bench.py (host-side numpy only).

* `perturbed(P, batch, ...)` - BASELINE.json configs[2]: one fixture, `batch` instances with
  h and b perturbed entry-wise, h_b = h (1 + rel u), u ~ U(-1,1); zeros stay zero; G, A, c shared.
  MPC_REL is the perturbation bench.py uses on MPC02: b +-2 %, h +-0.2 % - every instance stays feasible (MPC02 has paired bounds
  39 <= x <= 39.3, so +-5 % on h makes two thirds of the instances primal infeasible; +-5 % on b about 2 %).
* `soc_mpc(...)` - a builder-defined SOC-bearing MPC problem (the reference's MPC01 data file is
  missing from the checkout): 2-D double integrator, horizon T, tracking cost
  through second-order cones.  NOT reference data.
"""
import numpy as np

MPC_REL = {"h": 0.002, "b": 0.02}


def perturbed(P, batch, rel=0.05, seed=1234, vary=("h", "b")):
    """rel: one float, or a dict per vector, e.g. {"h": 0.002, "b": 0.05}."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    out = {}
    for k in ("c", "h", "b"):
        r = rel.get(k, 0.0) if isinstance(rel, dict) else rel
        if k in vary and np.asarray(P[k]).size and r:
            base = np.asarray(P[k], dtype=np.float64)
            u = rng.uniform(-1.0, 1.0, size=(batch, base.size))
            out[k + "s"] = base[None, :] * (1.0 + r * u)
        else:
            out[k + "s"] = None
    return out


def perturbed_matrices(P, batch, rel=0.01, seed=77, scale_spread=0.0):
    """Per-instance G / A VALUES on the shared pattern (what updateData(Gpr, Apr, ...) feeds the
    reference one instance at a time): entry-wise G_b = G (1 + rel u), u ~ U(-1,1).  scale_spread > 0
    also multiplies every instance's G rows of the LP cone and A rows by 10^(spread v), v ~ U(-1,1)
    per row, so that the per-instance equilibration really differs between instances; the matching
    h / b stacks are returned scaled the same way (the feasible set is unchanged by a row scaling)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    G0, A0 = np.asarray(P["Gpr"], np.float64), np.asarray(P["Apr"], np.float64)
    Gs = G0[None, :] * (1.0 + rel * rng.uniform(-1.0, 1.0, size=(batch, G0.size)))
    As = A0[None, :] * (1.0 + rel * rng.uniform(-1.0, 1.0, size=(batch, A0.size))) if A0.size else None
    out = {"Gs": Gs, "As": As, "hs": None, "bs": None}
    if scale_spread:
        m, p, l = int(P["m"]), int(P["p"]), int(P["l"])
        rs = np.ones((batch, m))
        rs[:, :l] = 10.0 ** (scale_spread * rng.uniform(-1.0, 1.0, size=(batch, l)))
        out["Gs"] = Gs * rs[:, np.asarray(P["Gir"], np.int64)]
        out["hs"] = np.asarray(P["h"], np.float64)[None, :] * rs
        if A0.size:
            ra = 10.0 ** (scale_spread * rng.uniform(-1.0, 1.0, size=(batch, p)))
            out["As"] = As * ra[:, np.asarray(P["Air"], np.int64)]
            out["bs"] = np.asarray(P["b"], np.float64)[None, :] * ra
    return out


def _csc(M):
    """dense -> (pr, jc, ir) with rows ascending per column, explicit zeros dropped"""
    M = np.asarray(M, dtype=np.float64)
    jc, ir, pr = [0], [], []
    for j in range(M.shape[1]):
        nz = np.nonzero(M[:, j])[0]
        ir.extend(nz.tolist())
        pr.extend(M[nz, j].tolist())
        jc.append(len(ir))
    return np.array(pr), np.array(jc, np.int32), np.array(ir, np.int32)


def soc_mpc(T=20, dt=0.25, umax=0.5, vmax=3.0, x0=(4.0, -3.0, 0.5, 0.0), ref=(0.0, 0.0, 0.0, 0.0), rho=0.1):
    """min sum_k t_k + rho r_k   s.t.  x_{k+1} = Ad x_k + Bd u_k, x_0 given,
    |v_k| <= vmax (LP), r_k <= umax (LP), ||u_k|| <= r_k (SOC dim 3), ||x_{k+1} - ref|| <= t_k (SOC dim 5).
    Variables per stage k=0..T-1: u_k (2), r_k, t_k, x_{k+1} (4)  -> n = 8 T."""
    Ad = np.eye(4)
    Ad[0, 2] = Ad[1, 3] = dt
    Bd = np.zeros((4, 2))
    Bd[0, 0] = Bd[1, 1] = 0.5 * dt * dt
    Bd[2, 0] = Bd[3, 1] = dt
    nv = 8
    n = nv * T
    iu, ir_, it, ix = 0, 2, 3, 4
    c = np.zeros(n)
    A = np.zeros((4 * T, n))
    b = np.zeros(4 * T)
    for k in range(T):
        o = nv * k
        c[o + it] = 1.0
        c[o + ir_] = rho
        A[4 * k:4 * k + 4, o + ix:o + ix + 4] = np.eye(4)
        A[4 * k:4 * k + 4, o + iu:o + iu + 2] = -Bd
        if k == 0:
            b[0:4] = Ad @ np.asarray(x0, dtype=np.float64)
        else:
            A[4 * k:4 * k + 4, o - nv + ix:o - nv + ix + 4] = -Ad
    rows_lp = 5 * T
    m = rows_lp + 3 * T + 5 * T
    G = np.zeros((m, n))
    h = np.zeros(m)
    r = 0
    for k in range(T):
        o = nv * k
        for d in (2, 3):  # |v| <= vmax
            G[r, o + ix + d] = 1.0
            h[r] = vmax
            r += 1
            G[r, o + ix + d] = -1.0
            h[r] = vmax
            r += 1
        G[r, o + ir_] = 1.0  # r_k <= umax
        h[r] = umax
        r += 1
    q = []
    for k in range(T):  # (r_k; u_k) in Q^3
        o = nv * k
        G[r, o + ir_] = -1.0
        G[r + 1, o + iu] = -1.0
        G[r + 2, o + iu + 1] = -1.0
        r += 3
        q.append(3)
    for k in range(T):  # (t_k; x_{k+1} - ref) in Q^5
        o = nv * k
        G[r, o + it] = -1.0
        for d in range(4):
            G[r + 1 + d, o + ix + d] = -1.0
            h[r + 1 + d] = -ref[d]
        r += 5
        q.append(5)
    assert r == m
    Gpr, Gjc, Gir = _csc(G)
    Apr, Ajc, Air = _csc(A)
    return dict(n=n, m=m, p=4 * T, l=rows_lp, ncones=len(q), q=np.array(q, np.int32),
                Gpr=Gpr, Gjc=Gjc, Gir=Gir, Apr=Apr, Ajc=Ajc, Air=Air, c=c, h=h, b=b,
                meta=dict(T=T, Ad=Ad, nv=nv))


def soc_mpc_batch(P, batch, seed=7):
    """Per-instance initial state (enters b) and reference (enters h); everything else shared."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    T, Ad = P["meta"]["T"], P["meta"]["Ad"]
    bs = np.repeat(P["b"][None, :], batch, axis=0)
    hs = np.repeat(P["h"][None, :], batch, axis=0)
    x0 = rng.uniform(-5, 5, size=(batch, 4)) * np.array([1, 1, 0.3, 0.3])
    ref = rng.uniform(-1, 1, size=(batch, 4)) * np.array([1, 1, 0, 0])
    bs[:, 0:4] = x0 @ Ad.T
    base = 5 * T + 3 * T
    for k in range(T):
        hs[:, base + 5 * k + 1: base + 5 * k + 5] = -ref
    return dict(cs=None, hs=hs, bs=bs)


def synthetic_socp(n=2000, m=3000, ncones=500, p=200, seed=7):
    """BASELINE.json configs[3] in the concrete form SURVEY.md 8(d) proposes (the literal statement -
    500 cones uniform on 3..10 inside m = 3000 rows - is inconsistent): cone dims q_i = 3 + floor(8 u_i^2),
    resampled until sum q <= m - 100; l = m - sum q LP rows; G has 4 entries per row, three inside a
    +-16 column window around row * n / m plus one uniform column (bounded fill); A has 3 per row by
    the same rule; values N(0,1).  Feasible by construction: x0 ~ N(0,1), s0 and z0 strictly interior,
    y0 ~ N(0,1), h = G x0 + s0, b = A x0, c = -G' z0 - A' y0.  NOT reference data."""
    rng = np.random.default_rng(seed)
    while True:
        q = (3 + np.floor(8.0 * rng.random(ncones) ** 2)).astype(np.int32)
        if q.sum() <= m - 100:
            break
    l = int(m - q.sum())

    def banded(rows, per_row):
        M = {}
        for r in range(rows):
            centre = r * n // rows
            cols = set()
            while len(cols) < per_row - 1:
                cols.add(int(np.clip(centre + rng.integers(-16, 17), 0, n - 1)))
            while len(cols) < per_row:
                cols.add(int(rng.integers(0, n)))
            for c in cols:
                M[(r, c)] = rng.standard_normal()
        return M

    def to_csc(M, rows):
        by_col = [[] for _ in range(n)]
        for (r, c), v in M.items():
            by_col[c].append((r, v))
        pr, ir, jc = [], [], [0]
        for c in range(n):
            for r, v in sorted(by_col[c]):
                ir.append(r)
                pr.append(v)
            jc.append(len(ir))
        return np.array(pr), np.array(jc, np.int32), np.array(ir, np.int32)

    Gm, Am = banded(m, 4), banded(p, 3)
    Gpr, Gjc, Gir = to_csc(Gm, m)
    Apr, Ajc, Air = to_csc(Am, p)

    def matvec(pr, jc, ir, rows, x):
        out = np.zeros(rows)
        for c in range(n):
            sl = slice(jc[c], jc[c + 1])
            out[ir[sl]] += pr[sl] * x[c]
        return out

    def rmatvec(pr, jc, ir, y):
        return np.array([np.dot(pr[jc[c]:jc[c + 1]], y[ir[jc[c]:jc[c + 1]]]) for c in range(n)])

    x0, y0 = rng.standard_normal(n), rng.standard_normal(p)
    s0, z0 = np.abs(rng.standard_normal(m)) + 0.5, np.abs(rng.standard_normal(m)) + 0.5
    at = l
    for d in q:
        s0[at] = np.linalg.norm(s0[at + 1:at + d]) + 1.0
        z0[at] = np.linalg.norm(z0[at + 1:at + d]) + 1.0
        at += d
    return dict(n=n, m=m, p=p, l=l, ncones=int(ncones), q=q,
                Gpr=Gpr, Gjc=Gjc, Gir=Gir, Apr=Apr, Ajc=Ajc, Air=Air,
                c=-rmatvec(Gpr, Gjc, Gir, z0) - rmatvec(Apr, Ajc, Air, y0),
                h=matvec(Gpr, Gjc, Gir, m, x0) + s0, b=matvec(Apr, Ajc, Air, p, x0))
