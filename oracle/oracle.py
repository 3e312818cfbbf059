"""ctypes binding of oracle/eicos_oracle.cpp (see eicos_oracle.h for what each call restates)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libeicos_oracle.so")
LIB_LD = os.path.join(HERE, "_build", "libeicos_oracle_ld.so")  # the same restatement in 80-bit extended precision
FIXTURE_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "fixtures")

_lib = None


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("eicos_oracle.cpp", "eicos_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB) or not os.path.exists(LIB_LD) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


class Info(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("pcost", "dcost", "pres", "dres", "pinfres", "dinfres", "gap", "relgap",
                 "sigma", "mu", "step", "step_aff", "kapovert")] + \
               [(k, C.c_int) for k in
                ("pinf", "dinf", "has_pinfres", "has_dinfres", "has_relgap",
                 "iter", "iter_max", "nitref1", "nitref2", "nitref3")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.ora_setup.restype = C.c_void_p
        L.ora_setup.argtypes = [C.c_int] * 5 + [_ip, _dp, _ip, _ip, _dp, _ip, _ip, _dp, _dp, _dp]
        for f in (L.ora_update_data, L.ora_update_data_full):
            f.restype = None
            f.argtypes = [C.c_void_p] + [_dp] * 5
        L.ora_solve.restype = C.c_int
        L.ora_solve.argtypes = [C.c_void_p]
        L.ora_debug_line_search.restype = C.c_double
        L.ora_debug_line_search.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double]
        L.ora_misaligned_cones.restype = C.c_longlong
        L.ora_misaligned_cones.argtypes = [C.c_void_p]
        L.ora_get_solution.restype = None
        L.ora_get_solution.argtypes = [C.c_void_p] + [_dp] * 4
        L.ora_get_info.restype = None
        L.ora_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.ora_cleanup.restype = None
        L.ora_cleanup.argtypes = [C.c_void_p]
        L.ora_dims.restype = None
        L.ora_dims.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.ora_get_symbolic.restype = None
        L.ora_get_symbolic.argtypes = [C.c_void_p] + [_ip] * 6
        L.ora_debug_factor_init.restype = C.c_int
        L.ora_debug_factor_init.argtypes = [C.c_void_p]
        L.ora_debug_get_factor.restype = None
        L.ora_debug_get_factor.argtypes = [C.c_void_p, _dp, _dp]
        L.ora_debug_get_K.restype = None
        L.ora_debug_get_K.argtypes = [C.c_void_p, _dp]
        L.ora_debug_ldl_solve.restype = None
        L.ora_debug_ldl_solve.argtypes = [C.c_void_p, _dp, _dp]
        L.ora_debug_solve_kkt.restype = C.c_int
        L.ora_debug_solve_kkt.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int]
        L.ora_debug_get_equil.restype = None
        L.ora_debug_get_equil.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ora_debug_get_data.restype = None
        L.ora_debug_get_data.argtypes = [C.c_void_p] + [_dp] * 5
        L.ora_batch_run.restype = C.c_double
        L.ora_batch_run.argtypes = ([C.c_int] * 5 + [_ip, _dp, _ip, _ip, _dp, _ip, _ip, _dp, _dp, _dp] +
                                    [C.c_int] + [_dp] * 5 + [C.c_int, C.c_int] + [_ip, _ip] + [_dp] * 5)
        L.ora_debug_set_iter_max.restype = None
        L.ora_debug_set_iter_max.argtypes = [C.c_int]
        _lib = L
    return _lib


def debug_set_iter_max(iter_max):
    """Test hook: cap the iterations of every later solve (0 restores the reference's 100)."""
    lib().ora_debug_set_iter_max(int(iter_max))


def load_fixture(name):
    """Problem data parsed from the reference's test headers (tests/golden/make_fixtures.py)."""
    d = np.load(os.path.join(FIXTURE_DIR, name + ".npz"))
    out = {k: d[k] for k in d.files}
    for k in ("n", "m", "p", "l", "ncones"):
        out[k] = int(out[k])
    with open(os.path.join(FIXTURE_DIR, "manifest.json")) as f:
        out["expect"] = json.load(f)[name]["expect"]
    return out


def _problem_args(P):
    """Keep-alive list + ctypes args in ECOS_setup order; empty arrays become NULL like the fixtures do."""
    def arr(k, dt):
        a = np.ascontiguousarray(P.get(k, np.zeros(0)), dtype=dt)
        return a if a.size else None
    q = arr("q", np.int32)
    Gpr, Gjc, Gir = arr("Gpr", np.float64), arr("Gjc", np.int32), arr("Gir", np.int32)
    Apr, Ajc, Air = arr("Apr", np.float64), arr("Ajc", np.int32), arr("Air", np.int32)
    c, h, b = arr("c", np.float64), arr("h", np.float64), arr("b", np.float64)
    keep = [q, Gpr, Gjc, Gir, Apr, Ajc, Air, c, h, b]
    args = [int(P["n"]), int(P["m"]), int(P["p"]), int(P.get("l", 0)), int(0 if q is None else q.size),
            _i(q), _d(Gpr), _i(Gjc), _i(Gir), _d(Apr), _i(Ajc), _i(Air), _d(c), _d(h), _d(b)]
    return keep, args


def solve_extended_precision(P, trace=False):
    """Exit flag of the long-double build of the oracle on problem P (and, with trace=True, the per-iteration lines
    it prints under ORA_VERBOSE: run in a subprocess so that the output can be captured)."""
    import sys
    code = ("import ctypes as C, sys; sys.path.insert(0, %r); import oracle; from oracle.oracle import _problem_args, LIB_LD, _dp, _ip;"
            "P = oracle.load_fixture(%r) if isinstance(%r, str) else None;"
            "L = C.CDLL(LIB_LD); L.ora_setup.restype = C.c_void_p;"
            "L.ora_setup.argtypes = [C.c_int] * 5 + [_ip, _dp, _ip, _ip, _dp, _ip, _ip, _dp, _dp, _dp];"
            "L.ora_solve.restype = C.c_int; L.ora_solve.argtypes = [C.c_void_p];"
            "keep, args = _problem_args(P); h = L.ora_setup(*args); print('EXIT', L.ora_solve(h))") % (os.path.dirname(HERE), P, P)
    env = dict(os.environ)
    if trace:
        env["ORA_VERBOSE"] = "1"
    out = subprocess.check_output([sys.executable, "-c", code], env=env, text=True)
    exitflag = int([l for l in out.splitlines() if l.startswith("EXIT")][-1].split()[1])
    return (exitflag, out) if trace else exitflag


def solve_trace(name):
    """Exit flag and ORA_VERBOSE trace of the double-precision oracle on fixture `name` (subprocess)."""
    import sys
    code = ("import sys; sys.path.insert(0, %r); import oracle; S = oracle.OracleSolver(oracle.load_fixture(%r)); print('EXIT', S.solve())"
            % (os.path.dirname(HERE), name))
    out = subprocess.check_output([sys.executable, "-c", code], env=dict(os.environ, ORA_VERBOSE="1"), text=True)
    return int([l for l in out.splitlines() if l.startswith("EXIT")][-1].split()[1]), out


class OracleSolver:
    """Mirrors EiCOS::Solver's pointer interface (include/eicos.hpp:151-163 of the reference)."""

    def __init__(self, P):
        self.P = P
        self._keep, args = _problem_args(P)
        self.n, self.m, self.p = args[0], args[1], args[2]
        if args[6] is None:  # no G => reference leaves n_ineq = 0
            self.m = 0
        if args[9] is None:
            self.p = 0
        self.h = lib().ora_setup(*args)

    def close(self):
        if self.h:
            lib().ora_cleanup(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self):
        return lib().ora_solve(self.h)

    def line_search(self, lam, ds, dz, tau, dtau, kap, dkap):
        a = [np.ascontiguousarray(v, np.float64) for v in (lam, ds, dz)]
        return float(lib().ora_debug_line_search(self.h, *[v.ctypes.data_as(_dp) for v in a], tau, dtau, kap, dkap))

    def misaligned_cones(self):
        """How often lineSearch skipped the cone-offset advance with cones still to come (src/eicos.cpp:1423-1424)."""
        return int(lib().ora_misaligned_cones(self.h))

    def update_data(self, Gpr=None, Apr=None, c=None, h=None, b=None, full=False):
        a = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (Gpr, Apr, c, h, b)]
        (lib().ora_update_data_full if full else lib().ora_update_data)(self.h, *[_d(v) for v in a])

    def solution(self):
        x, y, z, s = (np.zeros(k) for k in (self.n, self.p, self.m, self.m))
        lib().ora_get_solution(self.h, _d(x), _d(y), _d(z), _d(s))
        return x, y, z, s

    def info(self):
        i = Info()
        lib().ora_get_info(self.h, C.byref(i))
        return i.asdict()

    def dims(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().ora_dims(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def symbolic(self):
        N, nnzK, nnzL = self.dims()
        pinv, parent = np.zeros(N, np.int32), np.zeros(N, np.int32)
        Lp, Li = np.zeros(N + 1, np.int32), np.zeros(nnzL, np.int32)
        Kp, Ki = np.zeros(N + 1, np.int32), np.zeros(nnzK, np.int32)
        lib().ora_get_symbolic(self.h, _i(pinv), _i(parent), _i(Lp), _i(Li), _i(Kp), _i(Ki))
        return dict(pinv=pinv, parent=parent, Lp=Lp, Li=Li, Kp=Kp, Ki=Ki)

    def factor_init(self):
        return lib().ora_debug_factor_init(self.h)

    def factor(self):
        N, _, nnzL = self.dims()
        Lx, D = np.zeros(nnzL), np.zeros(N)
        lib().ora_debug_get_factor(self.h, _d(Lx), _d(D))
        return Lx, D

    def K_values(self):
        _, nnzK, _ = self.dims()
        Kx = np.zeros(nnzK)
        lib().ora_debug_get_K(self.h, _d(Kx))
        return Kx

    def ldl_solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros_like(rhs)
        lib().ora_debug_ldl_solve(self.h, _d(rhs), _d(x))
        return x

    def solve_kkt(self, rhs, initialize):
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        dx, dy, dz = np.zeros(self.n), np.zeros(self.p), np.zeros(self.m)
        k = lib().ora_debug_solve_kkt(self.h, _d(rhs), _d(dx), _d(dy), _d(dz), int(initialize))
        return k, dx, dy, dz

    def equil(self):
        xe, Ae, Ge = np.zeros(self.n), np.zeros(self.p), np.zeros(self.m)
        lib().ora_debug_get_equil(self.h, _d(xe), _d(Ae), _d(Ge))
        return xe, Ae, Ge


def batch_run(P, batch, Gs=None, As=None, cs=None, hs=None, bs=None, nthreads=1, want_solution=True, reset_sticky=True):
    """CPU baseline driver: one solver per thread, updateData + solve per instance."""
    keep, args = _problem_args(P)
    n, m, p = args[0], args[1], args[2]
    st = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (Gs, As, cs, hs, bs)]
    ex, it = np.zeros(batch, np.int32), np.zeros(batch, np.int32)
    pc = np.zeros(batch)
    if want_solution:
        xs, ys, zs, ss = np.zeros((batch, n)), np.zeros((batch, p)), np.zeros((batch, m)), np.zeros((batch, m))
    else:
        xs = ys = zs = ss = None
    secs = lib().ora_batch_run(*args, batch, *[_d(v) for v in st], int(nthreads), int(reset_sticky),
                               _i(ex), _i(it), _d(xs), _d(ys), _d(zs), _d(ss), _d(pc))
    return dict(seconds=secs, exit=ex, iter=it, pcost=pc, x=xs, y=ys, z=zs, s=ss)
