#!/usr/bin/env python
"""Measurement of the 'next' rows of SURVEY.md §8f on one GPU: the explicit step with penalty contact (two rigid planes,
friction), with thermal coupling, and with a Johnson-Cook flow stress, on tetrahedral boxes — element-steps/s on the GPU
next to the reference's CPU path (oracle/_ref, all host threads) on a bounded sample of the same workload.

    python tools/contact_bench.py [--n 60] [--cpu-n 20] [--steps 200]  ->  one JSON line per workload
"""
import argparse
import dataclasses
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from weldformfem_b200 import cases  # noqa: E402


def workloads(n):
    R = dataclasses.replace
    dense = lambda c: R(c, planes=tuple(dict(p, dens=20) for p in c.planes))   # 2 x 800 facets, Contact_Compression_tetra.json
    return {
        "tet_plain": R(cases.c2_tets(n), top_vel=-200.0),
        "tet_contact": dense(cases.contact_tets(n)),
        "tet_contact_thermal": cases.with_thermal(dense(cases.contact_tets(n)), heat_cond=25000.0, T_die=200.0),
        "tet_johnson_cook": cases.with_johnson_cook(R(cases.c2_tets(n), top_vel=-200.0)),
    }


def gpu_rate(case, steps):
    import torch
    from weldformfem_b200.domain import Domain_d
    d = Domain_d(device=0, strict=False)
    case.apply(d)
    d.step(20)
    d.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.Stream()
    d.set_stream(s.cuda_stream)
    ev0.record(s)
    d.step(steps)
    ev1.record(s)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out = {"ms_per_step": ms / steps, "value": case.n_elems * steps / (ms * 1e-3), "n_elems": case.n_elems,
           "nonfinite": d.nonfinite_flag()}
    if case.contact is not None:
        out["nodes_in_contact"] = int((d.get("m_mesh_in_contact") >= 0).sum())
        out["external_nodes"] = int(d.get("ext_nodes").sum())
    if case.thermal is not None:
        out["T_max"] = float(d.get("T").max())
    d.close()
    return out


def cpu_rate(case, steps):
    from oracle import refdrv
    cls = refdrv.RefDomain if refdrv.have_ref() else refdrv.OracleDomain
    cls.set_threads(os.cpu_count() or 1)
    d = cls()
    case.apply(d)
    d.step(2)
    t = d.time_steps(steps)
    return {"value": case.n_elems * steps / t, "cores": os.cpu_count(), "kind": "reference" if cls is refdrv.RefDomain else "port",
            "sample": f"{case.name}: {case.n_elems} elements x {steps} steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=60)
    ap.add_argument("--cpu-n", type=int, default=16)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    big, small = workloads(a.n), workloads(a.cpu_n)
    for key in big:
        line = {"workload": key, "metric": "element-steps/s", "gpu": gpu_rate(big[key], a.steps)}
        if not a.no_cpu:
            line["cpu_baseline"] = cpu_rate(small[key], 5)
            line["gpu_over_cpu"] = line["gpu"]["value"] / line["cpu_baseline"]["value"]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
