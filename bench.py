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

# algorithmic bytes per element-step of the 4-pass schedule as SURVEY.md §8(d) models it: every element node writes
# and reads its own force record (8 k d bytes each way)
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
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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
    hg_note = ("; the hexa viscous hourglass is harness code restated from f90_ver/src/Mechanical.f90:241-344 (absent from the C++ at this commit)"
               if kind == "hex" and tag == "reference" else "")
    return {"value": rate, "unit": "element-steps/s", "cores": ncores, "kind": tag,
            "sample": f"{case.name}: {case.n_elems} elements x {steps} steps after {warmup} warm-up, elastic regime (no "
                      f"plastic pre-load), "
                      f"{'oracle/_ref (unmodified reference, g++ -O2 -fopenmp)' if tag == 'reference' else 'oracle port (plain C, gcc -O2 -fopenmp)'}"
                      + hg_note,
            "same_config": False,
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
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config")},
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
# ... and of the schedule that SHIPS (DESIGN.md §3 derives every term): the fast 3D flavours hand forces on as one
# partial per (32-element tile, unique node) instead of one record per element node, and the hexa passes read packed
# per-element index records instead of connectivity + scatter offsets.  Minimal DRAM bytes: every array element once
# per pass that uses it, index tables included.  2D keeps the node-ordered buffer, i.e. the §8(d) model.
PASS_BYTES_SHIPPED = {"hex": (61.5, 88.0, 299.0, 275.0), "tet": (28.0, 45.0, 225.0, 93.0), "quad": (40.0, 72.0, 320.0, 256.0)}
ALG_BYTES_SHIPPED = {k: sum(v) for k, v in PASS_BYTES_SHIPPED.items()}


def ncu_traffic(kind):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kind)
    except Exception:
        return None


def _events(torch, stream):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed_steps(torch, dom, stream, steps, sync):
    """device time (ms) of `steps` fused steps: CUDA events on the launch stream, synchronised on both sides"""
    sync()
    ev0, ev1 = _events(torch, stream)
    ev0.record(stream)
    dom.step(steps)
    ev1.record(stream)
    sync()
    return ev0.elapsed_time(ev1)


def side_config(torch, stream, local, kind, n, label, preload, top_vel, min_seconds=0.3):
    """One of the other BASELINE.json configurations on one GPU (reported in `other_configs`): state resident, a short
    plastic pre-load, then >= min_seconds of fused steps timed with CUDA events."""
    from weldformfem_b200.domain import Domain_d
    case = dataclasses.replace(build_case(n, kind), top_vel=top_vel)   # fast enough to be plastic after the pre-load
    dom = Domain_d(device=local)
    case.apply(dom, init=False)
    dom.set_stream(stream.cuda_stream)
    dom.init(case.timestep)
    nn, ne, _ = dom.counts()
    dom.set("v", linear_velocity(case, nn))
    dom.step(preload)
    sync = torch.cuda.synchronize
    ms10 = timed_steps(torch, dom, stream, 10, sync)
    steps = int(max(20, min(20000, min_seconds * 1e3 / max(ms10 / 10, 1e-4))))
    timed_steps(torch, dom, stream, min(steps, 50), sync)
    ms = timed_steps(torch, dom, stream, steps, sync)
    ksteps = min(steps, 500)
    kms = [t / ksteps for t in dom.step_timed(ksteps)]
    plastic = float((dom.get("pl_strain") > 0).mean())
    bad = dom.nonfinite_flag()
    dom.close()
    peak, _ = measured_peak()
    rate = ne * steps / (ms * 1e-3)
    out = {"config": label, "kind": kind, "n_elems": int(ne), "n_nodes": int(nn), "steps": steps, "ms_per_step": ms / steps,
           "value": rate, "unit": "element-steps/s", "preload_steps": preload, "plastic_fraction": plastic, "nonfinite": bool(bad),
           "frac_shipped": rate * ALG_BYTES_SHIPPED[kind] / 1e9 / peak, "frac_model8d": rate * ALG_BYTES[kind] / 1e9 / peak,
           "l2_policy": ("working set %.0f MB: L2-resident (126 MB), launch-latency-bound" % (ne * ALG_BYTES_SHIPPED[kind] / 1e6)
                         if ne * ALG_BYTES_SHIPPED[kind] < 126e6 else "inputs larger than L2")}
    if kms:
        out["pass_ms"] = dict(zip(["E1", "N1", "E2", "N2"], [round(t, 5) for t in kms[1:5]]))
    return out


def ipc_parity_check(torch, dist, rank, world, local, stream, halo):
    """Untimed self-check of the REAL multi-process path (CUDA IPC + NVLink peer stores): a 48^3 hexa block stepped 20
    times by the `world` ranks against the one-GPU engine run by every rank on the whole block; worst relative error
    over the nodal and element state of all ranks."""
    from weldformfem_b200.distributed import RankDomain
    from weldformfem_b200.domain import Domain_d
    case = dataclasses.replace(cases.c3_hexes(48), top_vel=-200.0)
    one = Domain_d(device=local)
    case.apply(one)
    one.step(20)
    dom = RankDomain(rank, world, device=local, halo=halo)
    case.apply(dom, init=False)
    dom.set_stream(stream.cuda_stream)
    dom.connect()
    dist.barrier()
    dom.init(case.timestep)
    dom.step(20)
    dom.synchronize()
    ids = dom.node_l2g
    eb, ee = dom.elem_range()
    worst = 0.0
    for nm, per, nodal in (("x", 3, True), ("v", 3, True), ("u", 3, True), ("prev_a", 3, True), ("m_tau", 6, False),
                           ("pl_strain", 1, False), ("p", 1, False), ("sigma_y", 1, False)):
        g = one.get(nm).reshape(-1, per)
        w = g[ids] if nodal else g[eb:ee]
        got = dom.get(nm).reshape(-1, per)
        scale = max(float(np.abs(g).max()), 1e-300)
        worst = max(worst, float(np.abs(got - w).max()) / scale)
    plastic = float((one.get("pl_strain") > 0).mean())
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    dom.close()
    one.close()
    return {"workload": "48^3 hexes (110 592), 20 steps, hourglass 0.06, %d ranks over %s halo vs the one-GPU engine" % (world, halo),
            "max_rel_err": float(t.item()), "tolerance": 1e-11, "ok": bool(t.item() <= 1e-11), "plastic_fraction": plastic}


def strong_config(torch, dist, rank, world, local, stream, halo, kind, cube, one_gpu_rate=None, min_seconds=0.3):
    """One n^3 block partitioned over all ranks (strong scaling / configs[4]): total element-steps/s, max over ranks."""
    from weldformfem_b200.distributed import RankDomain
    case = build_case(cube, kind)
    dom = RankDomain(rank, world, device=local, halo=halo)
    case.apply(dom, init=False)
    dom.set_stream(stream.cuda_stream)
    dom.connect()
    dist.barrier()
    dom.init(case.timestep)
    nn, ne, _ = dom.counts()
    dom.set("v", linear_velocity(case, dom.node_l2g))
    dom.step(200)

    def sync():
        dist.barrier()
        torch.cuda.synchronize()
    ms10 = timed_steps(torch, dom, stream, 10, sync)
    t = torch.tensor([ms10], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    steps = int(max(20, min(5000, min_seconds * 1e3 / max(float(t.item()) / 10, 1e-4))))
    ms = timed_steps(torch, dom, stream, steps, sync)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    bad = dom.nonfinite_flag()
    dom.halo_status()
    dom.close()
    dist.barrier()
    peak, _ = measured_peak()
    rate = case.n_elems * steps / (ms * 1e-3)
    return {"kind": kind, "cube": cube, "n_elems_total": int(case.n_elems), "n_elems_per_gpu": int(ne), "n_gpus": world,
            "steps": steps, "ms_per_step": ms / steps, "value": rate, "unit": "element-steps/s", "nonfinite": bool(bad),
            "frac_shipped_per_gpu": rate / world * ALG_BYTES_SHIPPED[kind] / 1e9 / peak}


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

    # ---- value: exactly K fused steps, state resident in HBM, CUDA events on the launch stream ------------
    # The clock sampler covers warm-up, the K-step region AND a sustained region of >= 0.5 s that follows it (the
    # K-step region of the default run is tens of milliseconds: too short for nvidia-smi to sample).
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        dom.step(1)
    ms = max_over_ranks(timed_steps(torch, dom, stream, args.steps, barrier))
    sus_steps = int(max(args.steps, min(20000, 0.5 * 1e3 / max(ms / args.steps, 1e-4))))
    if args.no_sustained:
        sus_steps, sus_ms = args.steps, ms
    else:
        sus_ms = max_over_ranks(timed_steps(torch, dom, stream, sus_steps, barrier))
    clocks = sampler.stop() if rank == 0 else None
    flag = dom.nonfinite_flag()
    if world > 1:
        dom.halo_status()

    # ---- per-kernel device times (CUDA events around every launch), single GPU ----------------------
    kms, timeline = None, None
    if world == 1:
        dom.step_timed(3)
        kms = [t / args.steps for t in dom.step_timed(args.steps)]
    elif args.halo == "peer":
        # where a distributed step spends its time: an event after each of its launches, max over ranks per slot.  The
        # folded waits sit at the head of "shared-node sums" and "N2 shared nodes": a late neighbour shows up there.
        dom.step_timed(3)
        barrier()   # the ranks enter the measured batch together: host-side skew would be booked on the first flag wait
        kt = torch.tensor(dom.step_timed(args.steps), device="cuda", dtype=torch.float64) / args.steps
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
        kt = [float(v) for v in kt.tolist()]
        timeline = {"unit": "ms per step, max over ranks", "E1": kt[1], "N1 (+ volume-partial send)": kt[2],
                    "shared-node sums (+ wait for exchange 1)": kt[5], "E2": kt[3],
                    "N2 of the nodes not shared (+ force-partial send)": kt[4],
                    "N2 of the shared nodes (+ wait for exchange 2)": kt[6], "halo kernels of their own": kt[7],
                    "sum": sum(kt[1:8])}

    # ---- e2e: the call sequence a host solver loop makes through the C ABI with HOST buffers ----------
    # every step: new prescribed-velocity values for the moving plane (H2D from host memory via wf_set_bc_values: pinned
    # staging, copy stream, overlapped with the running step), one step that leaves the engine in predicted state
    # (wf_step_open: single-step calls keep the fused schedule of a batch), and the step monitor (kinetic energy +
    # non-finite flag, D2H) read back one step behind.
    e2e_steps = int(max(3, min(20000, max(args.steps, 0.3 * 1e3 / max(ms / args.steps, 1e-4)))))
    bcn, bcd, bcv = case.bc_arrays()
    d_last = case.dim - 1
    vals_last = np.ascontiguousarray(bcv[bcd == d_last])
    nrows = len(np.unique(bcn if world == 1 else np.intersect1d(bcn, node_ids)))
    open_ok = not args.strict and (world == 1 or args.halo == "peer")   # the host-driven (NCCL) transport steps phase by phase
    step1 = dom.step_open if open_ok else dom.step
    for i in range(3):                                # warm-up of the loop itself (second BC buffer, pinned ring)
        dom.set_bc_values(d_last, vals_last)
        step1(1)
        dom.monitor_async()
        dom.monitor_wait()
    barrier()
    t0 = time.perf_counter()
    ek = 0.0
    for i in range(e2e_steps):
        dom.set_bc_values(d_last, vals_last)     # H2D: this step's prescribed velocities, from host memory
        step1(1)
        dom.monitor_async()                       # D2H: kinetic energy + non-finite flag of this step (pinned)
        if i >= 1:                                # read the previous step's monitor while this one runs
            ek, bad = dom.monitor_wait()
            flag = flag or bad
    ek, bad = dom.monitor_wait()
    flag = flag or bad
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if open_ok:
        dom.step_close()
    total_elems = ne * world if world == 1 else case.n_elems
    value = total_elems * args.steps / (ms * 1e-3)
    e2e_value = total_elems * e2e_steps / e2e_s
    dom.close()

    # ---- the other configurations / scaling records (untimed with respect to `value`) -------------------------------
    other, strong, parity = None, None, None
    if world == 1 and not args.no_other and kind == "hex" and not args.strict:
        other = []
        for k2, n2, label, pre, tv in (("tet", 26, "configs[1] size: 105 456 constant-stress tets (structured 6-tet split of a 26^3 box)", 400, -10.0),
                                       ("tet", 118, "9.86 M tets (118^3 x 6)", 1200, -25.0),
                                       ("quad", 1000, "configs[3]: 1 M axisymmetric quads with hourglass", 5000, -8.0)):
            try:
                other.append(side_config(torch, stream, local, k2, n2, label, pre, tv))
            except Exception as ex:
                other.append({"config": label, "error": str(ex)[:200]})
    if world > 1 and not args.cube and not args.no_other and not args.strict:
        try:
            parity = ipc_parity_check(torch, dist, rank, world, local, stream, args.halo)
        except Exception as ex:
            parity = {"ok": False, "error": str(ex)[:300]}
        strong = []
        for k2, cube in (("hex", 215), ("hex", 431), ("tet", 118)):
            try:
                strong.append(strong_config(torch, dist, rank, world, local, stream, args.halo, k2, cube))
            except Exception as ex:
                strong.append({"kind": k2, "cube": cube, "error": str(ex)[:300]})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    alg, alg_s = ALG_BYTES[kind], ALG_BYTES_SHIPPED[kind]
    per_gpu_rate = total_elems / world * args.steps / (ms * 1e-3)
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_src,
            "definition": "achieved = algorithmic bytes of the SHIPPED schedule (DESIGN.md 3: tile partials instead of one "
                          "force record per element node) / CUDA-event time; frac_model8d uses the SURVEY 8(d) byte model "
                          "(bytes the shipped schedule no longer moves are charged: it can exceed 1)"}
    if kms is not None:
        names = ["E1 element volume", "N1 nodal sums", "E2 main element pass", "N2 assembly+integration"]
        pb, ps = PASS_BYTES[kind], PASS_BYTES_SHIPPED[kind]
        dom_i = int(np.argmax(kms[1:5]))
        ach = ps[dom_i] * ne / (kms[1 + dom_i] * 1e-3) / 1e9
        tr = ncu_traffic(kind)
        roof.update({"achieved": ach, "frac": ach / peak, "frac_shipped": ach / peak,
                     "frac_model8d": pb[dom_i] * ne / (kms[1 + dom_i] * 1e-3) / 1e9 / peak, "kernel": names[dom_i],
                     "algorithmic_bytes_per_launch": ps[dom_i] * ne, "kernel_ms": kms[1 + dom_i],
                     "traffic": (tr or {}).get("bytes_per_launch") if tr and tr.get("n_elems") == ne else None,
                     "traffic_source": (tr or {}).get("source"),
                     "passes": {nm: {"ms": kms[1 + i], "GB/s": ps[i] * ne / (kms[1 + i] * 1e-3) / 1e9,
                                     "frac_shipped": ps[i] * ne / (kms[1 + i] * 1e-3) / 1e9 / peak,
                                     "frac_model8d": pb[i] * ne / (kms[1 + i] * 1e-3) / 1e9 / peak}
                                for i, nm in enumerate(names)}})
        # measured DRAM bytes (ncu, profiles/traffic.json) over the live kernel time: what the memory system really moved
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
                            "note": "ncu dram__bytes_read+write per launch (profiles/traffic.json) / CUDA-event time of this run"}
    if kms is None:  # N > 1: no per-kernel timing hook; the whole fused step per GPU
        g = per_gpu_rate * alg_s / 1e9
        roof.update({"achieved": g, "frac": g / peak, "frac_shipped": g / peak, "frac_model8d": per_gpu_rate * alg / 1e9 / peak,
                     "kernel": "whole step (E1+N1+E2+N2 + halo kernels)"})
    roof["whole_step"] = {"bytes_per_element_step_shipped": alg_s, "bytes_per_element_step_model8d": alg,
                          "achieved": per_gpu_rate * alg_s / 1e9, "frac_shipped": per_gpu_rate * alg_s / 1e9 / peak,
                          "frac_model8d": per_gpu_rate * alg / 1e9 / peak,
                          "note": "per GPU; all four passes, CUDA events on the launch stream"}
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cpu = cpu_run(kind, args.cpu_n, 10, 2)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config")}
        except Exception as ex:  # the checker is optional for the bench line
            cpu = {"value": None, "unit": "element-steps/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    # N = 1: E1, N1, E2, N2.  N > 1 (peer transport, sends / waits folded into the node passes): + the nodal-sum finish of
    # the shared nodes and the node pass of the shared nodes = 6; WF_HALO_FOLD=0 keeps send / wait kernels of their own = 10
    folded = os.environ.get("WF_HALO_FOLD", "1") != "0" and args.halo == "peer"
    launches_per_step = 4 if world == 1 else (6 if folded else 10)
    line = {
        "metric": "element-steps/s", "value": value, "unit": "element-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if (world > 1 and args.cube) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (WORKLOAD[kind] if not (world > 1 and args.cube) else "configs[4]: synthetic hexa/tet block domain-decomposed across the GPUs") + f": {ne} elements / {nn} nodes per GPU "
                               f"(global box {'x'.join(str(q) for q in case.n)}), Hollomon J2" +
                               (f", viscous hourglass {case.hexa_hg}" if kind == "hex" else ""),
                   "why_this_config": "BASELINE.json quotes its target (>=60 % of HBM roofline) on the 10M-element hexa "
                                      "compression step = configs[2], the largest single-GPU configuration; configs[1], "
                                      "configs[3] and the 9.9 M-tet block are in other_configs, configs[4] in strong (N > 1)",
                   "numerics": "strict" if args.strict else "fast", "preload_steps": args.preload,
                   "plastic_fraction": plastic_frac, "hardening_fraction": harden_frac,
                   "l2_policy": "inputs larger than L2: every step streams >4 GB per GPU through the 126 MB L2, no flush needed",
                   "halo": (args.halo if world > 1 else None),
                   "nonfinite": bool(flag), "preload_seconds": preload_s, "kinetic_energy": ek},
        "sustained": {"steps": sus_steps, "ms_per_step": sus_ms / sus_steps, "value": total_elems * sus_steps / (sus_ms * 1e-3),
                      "seconds": sus_ms * 1e-3, "note": "same loop run for >= 0.5 s right after the K-step region (clock samples cover both)"},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "element-steps/s", "h2d_bytes_per_step": int(3 * nrows * 8),
                "d2h_bytes_per_step": 8 * 256 + 8, "steps": e2e_steps, "seconds": e2e_s, "ratio_to_value": e2e_value / value,
                "what": "per step: wf_set_bc_values (pinned host -> device on a copy stream) + " + ("wf_step_open(1)" if open_ok else "wf_step(1)") +
                        " + wf_monitor_async; the monitor (256 partial kinetic-energy sums + non-finite flag) of step i is read on the host (wf_monitor_wait) while step i+1 runs"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }
    if timeline is not None:
        line["timeline"] = timeline
    if other is not None:
        line["other_configs"] = other
    if parity is not None:
        line["parity_check"] = parity
    if strong is not None:
        line["strong"] = strong
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
    ap.add_argument("--cpu-n", type=int, default=128,
                    help="elements per side of the CPU sample (128 -> 2 097 152 hexes, SURVEY.md 8(d): n = 100-128)")
    ap.add_argument("--preload", type=int, default=1200)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip other_configs (N=1) / parity_check + strong (N>1)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 0.5 s sustained region (profiling runs)")
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
