#!/usr/bin/env python3
"""Probe: S handles on ONE device, each looping K passes over its slice of the batch in its own host thread, thread k
started k * delay ms late (no barrier between passes): the slow tail of one slice overlaps the bulk of the others."""
import os, sys, time, threading, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import eicos_b200
from bench import make_problem

def run(batch, S, delays, steps=3, warmup=1):
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
    def part(k, reps, delay):
        if delay > 0:
            time.sleep(delay * 1e-3)
        lo, hi = bounds[k], bounds[k + 1]
        for _ in range(reps):
            solvers[k].solve_device(hi - lo, d_Gs=0, d_As=0, d_cs=0, d_hs=hs[lo:hi].data_ptr(), d_bs=bs[lo:hi].data_ptr(),
                                    d_x=x_d[lo:hi].data_ptr(), d_exit=exit_d[lo:hi].data_ptr(), d_iter=iter_d[lo:hi].data_ptr())
    def go(reps, delay):
        th = [threading.Thread(target=part, args=(k, reps, k * delay)) for k in range(S)]
        for t in th: t.start()
        for t in th: t.join()
    go(warmup, 0)
    for delay in delays:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        go(steps, delay)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0)
        print(json.dumps({"batch": batch, "handles": S, "delay_ms": delay, "steps": steps, "ms_per_step": dt * 1e3 / steps,
                          "solves_per_s": batch * steps / dt, "iter_sum": int(iter_d.sum().item())}), flush=True)
    for s in solvers:
        s.close()

if __name__ == "__main__":
    batch = int(sys.argv[1]); S = int(sys.argv[2]); steps = int(sys.argv[3])
    run(batch, S, [float(a) for a in sys.argv[4:]], steps=steps)
