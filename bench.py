#!/usr/bin/env python
"""bench.py — element-steps/s of the explicit Chung-Hulbert step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one explicit time step over the whole mesh (fused rows 1-22 of
Solver_explicit.C:524-978).  Workload at N=1: BASELINE.json configs[2], the synthetic structured
hexa cube (n=215 -> 9 938 375 elements) with reduced integration + viscous hourglass 0.06 and
Hollomon J2 plasticity; the state is pre-loaded by stepping from a uniform-compression velocity
field until the mesh is plastic (plastic fraction reported in `config`).  N>1: the same per-GPU
block on every rank (weak scaling), element-block partition with nodal halo sums.

Prints ONE JSON line (rank 0).  See the module docstring of each section for what is timed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

from weldformfem_b200 import cases

# algorithmic bytes per element-step of the 4-pass schedule (SURVEY.md §8d / DESIGN.md §4)
ALG_BYTES = {"hex": 1096.0, "tet": 537.0, "quad": 688.0}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def linear_velocity(case, nn):
    """Uniform-compression velocity field v_z(z) = top_vel * z / H on the AddBoxLength lattice."""
    d = case.dim
    n1 = [q + 1 for q in case.n]
    v = np.zeros((nn, d))
    layer = np.arange(nn) // (n1[0] * (n1[1] if d == 3 else 1))
    v[:, d - 1] = case.top_vel * layer / case.n[d - 1]
    return v.reshape(-1)


def build_case(n, kind):
    if kind == "hex":
        return cases.c3_hexes(n)
    if kind == "tet":
        return cases.c2_tets(n)
    return cases.c4_axisymm_quads(n)


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_run(kind, n, steps, warmup, threads=None):
    from oracle import refdrv
    if refdrv.have_ref():
        cls, tag = refdrv.RefDomain, "reference"
    else:
        refdrv.build("port")
        cls, tag = refdrv.OracleDomain, "port"
    case = build_case(n, kind)
    dom = cls()
    ncores = threads or os.cpu_count() or 1
    cls.set_threads(ncores)
    case.apply(dom)
    nn = dom.info()["n_nodes"]
    v = linear_velocity(case, nn)
    dom.set("v", v)
    if warmup:
        dom.step(warmup)
    t = dom.time_steps(steps)
    rate = case.n_elems * steps / t
    return {"value": rate, "unit": "element-steps/s", "cores": ncores, "kind": tag,
            "sample": f"{case.name}: {case.n_elems} elements x {steps} steps after {warmup} warm-up, "
                      f"{'oracle/_ref (unmodified reference, g++ -O2 -fopenmp)' if tag == 'reference' else 'oracle port (plain C, gcc -O2 -fopenmp)'}",
            "seconds": t, "ms_per_step": 1e3 * t / steps, "n_elems": case.n_elems}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = args.kind
    n = args.cpu_n
    r = cpu_run(kind, n, max(1, args.steps), max(0, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": "element-steps/s", "value": r["value"], "unit": "element-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{kind} box compression (CPU sample n={n}, {r['n_elems']} elements; rate is "
                                   f"size-independent, SURVEY.md §6)", "sample": r["sample"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from weldformfem_b200.domain import Domain_d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream()
    kind = args.kind
    case = build_case(args.n, kind)
    dom = Domain_d(device=local, strict=args.strict)
    case.apply(dom, init=False)
    dom.set_stream(stream.cuda_stream)
    dom.init(case.timestep)
    nn, ne, _ = dom.counts()
    dom.set("v", linear_velocity(case, nn))
    # pre-load: evolve until plastic (untimed workload construction)
    t0 = time.time()
    if args.preload:
        dom.step(args.preload)
        dom.synchronize()
    preload_s = time.time() - t0
    pl = dom.get("pl_strain")
    plastic_frac = float((pl > 0).mean())
    eps1 = (case.sy0 / case.K) ** (1.0 / case.m) - case.sy0 / case.E
    harden_frac = float((pl > eps1).mean())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    for _ in range(args.warmup):
        dom.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        dom.step(args.steps)
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    flag = dom.nonfinite_flag()

    # e2e: every step goes through the public C-ABI call with host buffers: the prescribed-velocity table
    # is re-uploaded from pinned host memory (H2D) and the step monitor (kinetic energy + non-finite flag)
    # is read back (D2H) inside the timed region.
    e2e_steps = max(3, min(args.steps, 20))
    bcn, bcd, bcv = case.bc_arrays()
    bc_host = torch.from_numpy(bcv.copy()).pin_memory()
    bc_dev = torch.empty_like(bc_host, device="cuda")
    barrier()
    t0 = time.perf_counter()
    ek = 0.0
    for _ in range(e2e_steps):
        with torch.cuda.stream(stream):
            bc_dev.copy_(bc_host, non_blocking=True)
        dom.step(1)
        ek, _ = dom.energies()
        if dom.nonfinite_flag():
            flag = True
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    total_elems = ne * world
    value = total_elems * args.steps / (ms * 1e-3)
    e2e_value = total_elems * e2e_steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    alg = ALG_BYTES[kind]
    achieved = alg * ne * args.steps / (ms * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cpu = cpu_run(kind, args.cpu_n, 10, 2)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the checker is optional for the bench line
            cpu = {"value": None, "unit": "element-steps/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    line = {
        "metric": "element-steps/s", "value": value, "unit": "element-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[2]: synthetic structured {kind} box compression n={args.n} "
                               f"({ne} elements, {nn} nodes per GPU), Hollomon J2, viscous hourglass {case.hexa_hg}",
                   "numerics": "strict" if args.strict else "fast", "preload_steps": args.preload,
                   "plastic_fraction": plastic_frac, "hardening_fraction": harden_frac,
                   "l2_policy": "working set (>4 GB per step) exceeds the 126 MB L2; no flush needed",
                   "nonfinite": bool(flag), "preload_seconds": preload_s, "kinetic_energy": ek},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "bytes_per_element_step": alg,
                     "kernels": "whole step (E1+N1+E2+N2), CUDA events on the launch stream"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "element-steps/s", "h2d_bytes_per_step": int(bc_host.numel() * 8),
                "d2h_bytes_per_step": 20, "steps": e2e_steps},
        "gpu_launches": 4 * args.steps + 1,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kind", default="hex", choices=["hex", "tet", "quad"])
    ap.add_argument("--n", type=int, default=215)
    ap.add_argument("--cpu-n", type=int, default=64)
    ap.add_argument("--preload", type=int, default=1200)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
