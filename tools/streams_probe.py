#!/usr/bin/env python3
"""Probe: S handles (one stream and one host thread each) on ONE device, each solving a contiguous slice of the batch."""
import os, sys, time, threading, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import eicos_b200
from bench import make_problem

def run(batch, S, steps=2, warmup=1):
    P, name, gen = make_problem("mpc02")
    W = gen(batch, 1234)
    dev = torch.device("cuda", 0)
    n = P["n"]
    hs = torch.from_numpy(np.ascontiguousarray(W["hs"])).to(dev)
    bs = torch.from_numpy(np.ascontiguousarray(W["bs"])).to(dev)
    x_d = torch.empty((batch, n), dtype=torch.float64, device=dev)
    exit_d = torch.empty((batch,), dtype=torch.int32, device=dev)
    iter_d = torch.empty((batch,), dtype=torch.int32, device=dev)
    bounds = [batch * k // S for k in range(S + 1)]
    solvers = [eicos_b200.BatchSolver(P, device=0, capacity=bounds[k + 1] - bounds[k]) for k in range(S)]
    torch.cuda.synchronize()
    def part(k):
        lo, hi = bounds[k], bounds[k + 1]
        solvers[k].solve_device(hi - lo, d_Gs=0, d_As=0, d_cs=0, d_hs=hs[lo:hi].data_ptr(), d_bs=bs[lo:hi].data_ptr(),
                                d_x=x_d[lo:hi].data_ptr(), d_exit=exit_d[lo:hi].data_ptr(), d_iter=iter_d[lo:hi].data_ptr())
    def step():
        th = [threading.Thread(target=part, args=(k,)) for k in range(S)]
        for t in th: t.start()
        for t in th: t.join()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    ex = exit_d.cpu().numpy(); it = iter_d.cpu().numpy()
    print(json.dumps({"batch": batch, "streams": S, "ms_per_step": dt * 1e3, "solves_per_s": batch / dt,
                      "exit0": int((ex == 0).sum()), "iter_sum": int(it.sum()), "x_sum": float(x_d.sum().item())}), flush=True)
    for s in solvers:
        s.close() if hasattr(s, "close") else None
    del solvers
    torch.cuda.empty_cache()

if __name__ == "__main__":
    batch = int(sys.argv[1]); 
    for S in [int(a) for a in sys.argv[2:]]:
        run(batch, S)
