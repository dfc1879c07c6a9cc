#!/usr/bin/env python
"""Host-time breakdown of the e2e loop of bench.py (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from weldformfem_b200 import cases
from weldformfem_b200.domain import Domain_d
from bench import linear_velocity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 215
case = cases.c3_hexes(n)
d = Domain_d()
case.apply(d, init=False)
if len(sys.argv) > 2 and sys.argv[2] == "torch":
    import torch
    st = torch.cuda.Stream()
    d.set_stream(st.cuda_stream)
d.init(case.timestep)
nn, ne, _ = d.counts()
d.set("v", linear_velocity(case, nn))
d.step(50); d.synchronize()
bcn, bcd, bcv = case.bc_arrays()
vals = np.ascontiguousarray(bcv[bcd == 2])
for mode in ("sync", "async", "step-only", "step+bc", "step+mon"):
    T = {"bc": 0.0, "step": 0.0, "mon": 0.0, "wait": 0.0}
    d.synchronize()
    t0 = time.perf_counter()
    N = 30
    for i in range(N):
        a = time.perf_counter()
        if mode in ("sync", "async", "step+bc"):
            d.set_bc_values(2, vals)
        b = time.perf_counter()
        d.step(1)
        c = time.perf_counter()
        if mode == "sync":
            d.energies(); d.nonfinite_flag()
        elif mode in ("async", "step+mon"):
            d.monitor_async()
        e = time.perf_counter()
        if mode in ("async", "step+mon") and i >= 1:
            d.monitor_wait()
        f = time.perf_counter()
        T["bc"] += b - a; T["step"] += c - b; T["mon"] += e - c; T["wait"] += f - e
    if mode in ("async", "step+mon"):
        d.monitor_wait()
    d.synchronize()
    tot = time.perf_counter() - t0
    print(mode, "ms/step %.3f" % (1e3 * tot / N), {k: round(1e3 * v / N, 3) for k, v in T.items()})
