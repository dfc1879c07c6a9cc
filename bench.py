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
import dataclasses
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


def linear_velocity(case, nodes):
    """Uniform-compression velocity field v_z(z) = top_vel * z / H on the AddBoxLength lattice, for the given
    (global) node ids — or for nodes 0..n-1 when an int is passed."""
    d = case.dim
    n1 = [q + 1 for q in case.n]
    ids = np.arange(nodes) if np.isscalar(nodes) else np.asarray(nodes)
    v = np.zeros((ids.size, d))
    layer = ids // (n1[0] * (n1[1] if d == 3 else 1))
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
WORKLOAD = {"hex": "configs[2]: synthetic structured hexa cube compression, reduced integration + viscous hourglass",
            "tet": "configs[1]-shaped: constant-stress tetra box compression (6 tets per cell), J2 plasticity",
            "quad": "configs[3]: 2D axisymmetric quad upsetting with hourglass"}

# algorithmic bytes per element of each pass, [E1, N1, E2, N2]
PASS_BYTES = {"hex": (64.0, 88.0, 456.0, 488.0), "tet": (28.0, 45.0, 297.0, 167.0), "quad": (40.0, 72.0, 320.0, 256.0)}


def ncu_traffic(kind):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kind)
    except Exception:
        return None


def ours(args):
    import torch
    import torch.distributed as dist
    from weldformfem_b200.domain import Domain_d
    from weldformfem_b200.distributed import RankDomain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    kind = args.kind
    case = build_case(args.n, kind)
    if world > 1 and args.cube:
        # configs[4]: one n^3 block (e.g. --cube 431 = 80 062 991 hexes) domain-decomposed across the ranks
        case = build_case(args.cube, kind)
        case = dataclasses.replace(case, top_vel=case.top_vel * args.cube / args.n)
        dom = RankDomain(rank, world, device=local, strict=args.strict, halo=args.halo)
    elif world > 1:
        # weak scaling: every rank owns one n^(d-1) x n slab of an n^(d-1) x (n * world) box
        nn_ = list(case.n)
        nn_[-1] *= world
        # same strain rate as the one-GPU workload: the moving plane is `world` times further from the clamped one
        case = dataclasses.replace(case, n=tuple(nn_), name=case.name + f"_x{world}", top_vel=case.top_vel * world)
        dom = RankDomain(rank, world, device=local, strict=args.strict, halo=args.halo)
    else:
        dom = Domain_d(device=local, strict=args.strict)
    case.apply(dom, init=False)
    dom.set_stream(stream.cuda_stream)
    if world > 1:
        dom.connect()
        dist.barrier()
    dom.init(case.timestep)
    nn, ne, _ = dom.counts()
    node_ids = dom.node_l2g if world > 1 else np.arange(nn)
    dom.set("v", linear_velocity(case, node_ids))
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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: K fused steps, state resident in HBM, CUDA events on the launch stream ------------
    for _ in range(args.warmup):
        dom.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    dom.step(args.steps)
    ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    flag = dom.nonfinite_flag()
    if world > 1:
        dom.halo_status()

    # ---- per-kernel device times (CUDA events around every launch), single GPU ----------------------
    kms = None
    if world == 1:
        dom.step_timed(3)
        kms = [t / args.steps for t in dom.step_timed(args.steps)]

    # ---- e2e: the call sequence a host solver loop makes through the C ABI with HOST buffers ----------
    # every step: new prescribed-velocity values for the moving plane (H2D from host memory via
    # wf_set_bc_values), one step, and the step monitor (kinetic energy + non-finite flag, D2H) read back.
    e2e_steps = max(3, min(args.steps, 50))
    bcn, bcd, bcv = case.bc_arrays()
    d_last = case.dim - 1
    vals_last = np.ascontiguousarray(bcv[bcd == d_last])
    nrows = len(np.unique(bcn if world == 1 else np.intersect1d(bcn, node_ids)))
    barrier()
    t0 = time.perf_counter()
    ek = 0.0
    for i in range(e2e_steps):
        dom.set_bc_values(d_last, vals_last)     # H2D: this step's prescribed velocities, from host memory
        dom.step(1)
        dom.monitor_async()                       # D2H: kinetic energy + non-finite flag of this step (pinned)
        if i >= 1:                                # read the previous step's monitor while this one runs
            ek, bad = dom.monitor_wait()
            flag = flag or bad
    ek, bad = dom.monitor_wait()
    flag = flag or bad
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    total_elems = ne * world if world == 1 else case.n_elems
    value = total_elems * args.steps / (ms * 1e-3)
    e2e_value = total_elems * e2e_steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    alg = ALG_BYTES[kind]
    step_gbs = alg * total_elems / world * args.steps / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_src}
    if kms is not None:
        names = ["E1 element volume", "N1 nodal sums", "E2 main element pass", "N2 assembly+integration"]
        pb = PASS_BYTES[kind]
        dom_i = int(np.argmax(kms[1:5]))
        ach = pb[dom_i] * ne / (kms[1 + dom_i] * 1e-3) / 1e9
        tr = ncu_traffic(kind)
        roof.update({"achieved": ach, "frac": ach / peak, "kernel": names[dom_i],
                     "algorithmic_bytes_per_launch": pb[dom_i] * ne, "kernel_ms": kms[1 + dom_i],
                     "traffic": (tr or {}).get("bytes_per_launch") if tr and tr.get("n_elems") == ne else None,
                     "traffic_source": (tr or {}).get("source"),
                     "passes": {nm: {"ms": kms[1 + i], "GB/s": pb[i] * ne / (kms[1 + i] * 1e-3) / 1e9,
                                     "frac": pb[i] * ne / (kms[1 + i] * 1e-3) / 1e9 / peak}
                                for i, nm in enumerate(names)}})
        # measured DRAM bytes (ncu, profiles/traffic.json) over the live kernel time: what the memory system really
        # moved.  The tile-reduced force path moves FEWER bytes than the SURVEY 8(d) model, so `frac` (algorithmic
        # bytes / time / peak, the contract's definition) can exceed the fraction of peak the DRAM actually ran at.
        if tr and tr.get("n_elems") == ne and tr.get("passes"):
            tot_b = 0.0
            for i, nm in enumerate(names):
                q = tr["passes"].get(nm)
                if not q:
                    continue
                b = q["dram_bytes_read"] + q["dram_bytes_write"]
                tot_b += b
                roof["passes"][nm].update({"dram_bytes": b, "dram_GB/s": b / (kms[1 + i] * 1e-3) / 1e9,
                                           "dram_frac": b / (kms[1 + i] * 1e-3) / 1e9 / peak})
            roof["dram"] = {"bytes_per_step": tot_b, "bytes_per_element_step": tot_b / ne,
                            "GB/s": tot_b / (sum(kms[1:5]) * 1e-3) / 1e9, "frac": tot_b / (sum(kms[1:5]) * 1e-3) / 1e9 / peak,
                            "note": "ncu dram__bytes_read+write per launch (profiles/r01c_*, r01d_*) / CUDA-event time of this run"}
    if kms is None:  # N > 1: no per-kernel timing hook; the whole fused step per GPU
        roof.update({"achieved": step_gbs, "frac": step_gbs / peak, "kernel": "whole step (E1+N1+E2+N2 + halo kernels)"})
    roof["whole_step"] = {"bytes_per_element_step": alg, "achieved": step_gbs, "frac": step_gbs / peak,
                          "note": "per GPU; all four passes, CUDA events on the launch stream"}
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cpu = cpu_run(kind, args.cpu_n, 10, 2)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the checker is optional for the bench line
            cpu = {"value": None, "unit": "element-steps/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    # N = 1: E1, N1, E2, N2.  N > 1: + halo send / wait / finish after E1, send after E2, wait, node pass of the shared nodes
    launches_per_step = 4 if world == 1 else 10
    line = {
        "metric": "element-steps/s", "value": value, "unit": "element-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if (world > 1 and args.cube) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (WORKLOAD[kind] if not (world > 1 and args.cube) else "configs[4]: synthetic hexa/tet block domain-decomposed across the GPUs") + f": {ne} elements / {nn} nodes per GPU "
                               f"(global box {'x'.join(str(q) for q in case.n)}), Hollomon J2" +
                               (f", viscous hourglass {case.hexa_hg}" if kind == "hex" else ""),
                   "why_this_config": "BASELINE.json quotes its target (>=60 % of HBM roofline) on the 10M-element hexa "
                                      "compression step = configs[2], the largest single-GPU configuration; "
                                      "--kind tet --n 26 and --kind quad --n 1000 run configs[1] and configs[3]",
                   "numerics": "strict" if args.strict else "fast", "preload_steps": args.preload,
                   "plastic_fraction": plastic_frac, "hardening_fraction": harden_frac,
                   "l2_policy": "inputs larger than L2: every step streams >4 GB per GPU through the 126 MB L2, no flush needed",
                   "halo": (args.halo if world > 1 else None),
                   "nonfinite": bool(flag), "preload_seconds": preload_s, "kinetic_energy": ek},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "element-steps/s", "h2d_bytes_per_step": int(3 * nrows * 8),
                "d2h_bytes_per_step": 16, "steps": e2e_steps,
                "what": "per step: wf_set_bc_values (host -> device) + wf_step(1) + wf_monitor_async; the monitor (kinetic energy, non-finite flag) of step i is read on the host (wf_monitor_wait) while step i+1 runs"},
        "gpu_launches": launches_per_step * args.steps,
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
    ap.add_argument("--n", "--size", dest="n", type=int, default=215,
                    help="elements per side (use --size under torchrun, whose own parser grabs --n as a prefix of --nnodes)")
    ap.add_argument("--cpu-n", type=int, default=64)
    ap.add_argument("--preload", type=int, default=1200)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--cube", type=int, default=0, help="N>1: partition one cube of this many elements per side (configs[4]: 431)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
