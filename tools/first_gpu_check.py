import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import OracleSolver, load_fixture, batch_run
import eicos_b200
from eicos_b200.workloads import perturbed
names = sorted(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/fixtures/manifest.json'))).keys(), key=str.lower)
bad = 0
for nm in names:
    P = load_fixture(nm)
    O = OracleSolver(P); co = O.solve(); io = O.info(); xo,yo,zo,so = O.solution()
    S = eicos_b200.Solver(P); ce = S.solve(); ie = S.info(); xe = S.solution(); ye,ze,se = S.duals()
    def rel(a,b): return float(np.max(np.abs(a-b))/max(1,np.max(np.abs(b)))) if a.size else 0.0
    ok = (co==ce and io['iter']==ie['iter'] and max(rel(xe,xo),rel(ye,yo),rel(ze,zo),rel(se,so))<=1e-7)
    bad += (not ok)
    print(f"{'OK ' if ok else 'BAD'} {nm:18s} oracle exit={co:3d} it={io['iter']:3d} | gpu exit={ce:3d} it={ie['iter']:3d} ir={ie['nitref1']},{ie['nitref2']},{ie['nitref3']} vs {io['nitref1']},{io['nitref2']},{io['nitref3']} dx={rel(xe,xo):.1e} dy={rel(ye,yo):.1e} dz={rel(ze,zo):.1e} ds={rel(se,so):.1e}", flush=True)
print("single-instance mismatches:", bad)
P = load_fixture('MPC02')
for B, workers in ((256,4),(4096,4),(4096,8),(4096,2),(16384,4)):
    W = perturbed(P, B)
    bs = eicos_b200.BatchSolver(P, capacity=B, workers=workers)
    bs.set_timing(True)
    t0=time.time(); out = bs.solve(B, hs=W['hs'], bs=W['bs'], want=("x",), want_info=False); t1=time.time()
    t0=time.time(); out = bs.solve(B, hs=W['hs'], bs=W['bs'], want=("x",), want_info=False); t1=time.time()
    st = bs.stats(); d = bs.dims()
    print(f"MPC02 B={B} workers={workers} wall={t1-t0:.3f}s solves/s={B/(t1-t0):.0f} exits={np.unique(out['exit'], return_counts=True)} stats={st}", flush=True)
    if B == 256:
        ref = batch_run(P, B, hs=W['hs'], bs=W['bs'], nthreads=8)
        print(" parity exit", np.array_equal(ref['exit'], out['exit']), "x relerr", np.max(np.abs(ref['x']-out['x'])/np.maximum(1,np.abs(ref['x']).max(axis=1,keepdims=True))), "cpu secs", ref['seconds'])
        print(d)
    bs.close()
