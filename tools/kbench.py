#!/usr/bin/env python
"""Per-kernel timing of the fused step for kernel variants (run under gpurun).
usage: tools/kbench.py [--n 215] [--preload 100] [--steps 20] --configs e2=0,n2=0 e2=1,n2=1 ..."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from weldformfem_b200 import cases
from weldformfem_b200.domain import Domain_d
from bench import linear_velocity

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=215)
ap.add_argument("--kind", default="hex")
ap.add_argument("--preload", type=int, default=100)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--strict", action="store_true")
ap.add_argument("--configs", nargs="*", default=["e2=0,n2=0"])
a = ap.parse_args()
case = {"hex": cases.c3_hexes, "tet": cases.c2_tets, "quad": cases.c4_axisymm_quads}[a.kind](a.n)
base = None
for cfg in a.configs:
    kv = dict(x.split("=") for x in cfg.split(",") if x)
    if "plan" in kv:
        os.environ["WF_BRICK_PLAN"] = kv["plan"]
    else:
        os.environ.pop("WF_BRICK_PLAN", None)
    d = Domain_d(strict=a.strict, elem_order=int(kv["order"]) if "order" in kv else None)
    case.apply(d)
    for name, idx in (("e1", 0), ("n1", 1), ("e2", 2), ("n2", 3)):
        if name in kv:
            d.set_variant(idx, int(kv[name]))
    nn, ne, _ = d.counts()
    d.set("v", linear_velocity(case, nn))
    d.step(a.preload)
    import torch
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.step(a.steps); d.synchronize(); wall_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    d.step_timed(3)
    ms = d.step_timed(a.steps)
    tot = sum(ms)
    rate = ne * a.steps / (tot * 1e-3)
    st = {nm: d.get(nm) for nm in ("x", "m_tau", "pl_strain", "v")}
    diff = {}
    if base is None:
        base = st
    else:
        for nm in st:
            diff[nm] = float(np.abs(st[nm] - base[nm]).max() / max(np.abs(base[nm]).max(), 1e-300))
    print(json.dumps({"cfg": cfg, "preload": a.preload, "ms_per_step": {k: round(v / a.steps, 4) for k, v in zip(["pred", "E1", "N1", "E2", "N2"], ms)},
                      "total_ms": round(tot / a.steps, 4), "batch_ms_per_step": round(wall_ms, 5), "rate": rate, 
                      "plastic": float((st["pl_strain"] > 0).mean()), "diff_vs_first": diff}), flush=True)
    d.close()
