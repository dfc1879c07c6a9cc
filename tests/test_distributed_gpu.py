"""GPU tests of the partitioned (multi-GPU) step.  All ranks run inside this process through
wf_step_all (LocalCluster); when the box has fewer GPUs than ranks, several ranks share cuda:0 —
the halo exchange (stores into the neighbour's receive region + flag, one-CTA wait kernel) is the
same code either way.  The partitioned result must agree with the one-GPU engine up to the association
of the shared-node sums, and with the CPU oracle within the tolerances of BASELINE.json."""
import dataclasses
import os

import numpy as np
import pytest

from parity_util import compare, relerr
from weldformfem_b200 import cases

pytestmark = pytest.mark.gpu
R = dataclasses.replace

os.environ.setdefault("WF_HALO_TIMEOUT_S", "10")

NAMES = "x v u prev_a m_mdiag vol p pl_strain sigma_y m_sigma m_tau".split()

CASES = {
    "hex": R(cases.c3_hexes(8), n=(6, 5, 9), top_vel=-200.0),
    "tet": R(cases.c2_tets(6), n=(4, 4, 7), top_vel=-200.0),
    "tet_anp_nodal": R(cases.c2_tets(6, press=3), n=(4, 4, 7), top_vel=-200.0),
    "psquad": R(cases.plane_strain_quads(12), n=(10, 9), top_vel=-50.0),
    "pstri": R(cases.plane_strain_tris(12), n=(10, 9), top_vel=-50.0),
    # axisymmetric: the axis constraint uses the rank-local minimum of x_r (wf_set_axis_xmin; every slab touches the axis)
    "axiquad": R(cases.c4_axisymm_quads(12), n=(10, 9), top_vel=-50.0),
}


def devices_for(nranks):
    import torch
    n = torch.cuda.device_count()
    return [p % n for p in range(nranks)]


def run_cluster(case, nranks, nsteps, strict=False):
    from weldformfem_b200.distributed import LocalCluster
    cl = LocalCluster(nranks, devices_for(nranks), strict=strict)
    case.apply(cl)
    if nsteps:
        cl.step(nsteps)
    return cl


def run_single(case, nsteps, strict=False):
    from weldformfem_b200.domain import Domain_d
    e = Domain_d(strict=strict)
    case.apply(e)
    if nsteps:
        e.step(nsteps)
    return e


@pytest.mark.parametrize("key", sorted(CASES))
@pytest.mark.parametrize("nranks", [2, 3])
def test_partitioned_matches_single_gpu(key, nranks):
    case = CASES[key]
    cl = run_cluster(case, nranks, 40)
    one = run_single(case, 40)
    assert (one.get("pl_strain") > 0).mean() > 0.3
    worst = {nm: relerr(cl.get(nm), one.get(nm)) for nm in NAMES}
    assert max(worst.values()) < 1e-11, worst
    # interior state is untouched by the partition: elements away from the cuts are bit-identical after 1 step
    cl.close(); one.close()


@pytest.mark.parametrize("key", ["hex", "tet"])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_partitioned_matches_oracle(key, strict, oracle_port):
    case = CASES[key]
    ref = oracle_port()
    case.apply(ref)
    cl = run_cluster(case, 2, 1, strict)
    ref.step(1)
    compare(cl, ref, NAMES + ["m_fi"], 1e-10, f"{key} partitioned, 1 step")
    cl.step(99); ref.step(99)
    compare(cl, ref, NAMES, 1e-7, f"{key} partitioned, 100 steps")
    assert not cl.nonfinite_flag()
    cl.close()


def test_shared_node_copies_are_bit_identical():
    """Every sharer sums the partials in ascending rank order, so all copies of a shared node carry the
    same bits (assemble_global raises otherwise) — including nodes shared by more than two ranks."""
    case = R(cases.c2_tets(6), n=(3, 3, 3), top_vel=-200.0)   # 162 tets over 4 ranks: ragged cuts, 3- and 4-way sharing
    cl = run_cluster(case, 4, 25)
    multi = np.zeros(cl.n_nodes, dtype=int)
    for r in cl.ranks:
        multi[r.node_l2g] += 1
    assert multi.max() >= 3
    for nm in ("x", "v", "u", "prev_a", "m_mdiag", "m_fi"):
        cl.get(nm)
    one = run_single(case, 25)
    assert relerr(cl.get("x"), one.get("x")) < 1e-12
    cl.close(); one.close()


def test_partitioned_is_deterministic():
    case = CASES["hex"]
    a = run_cluster(case, 2, 30)
    b = run_cluster(case, 2, 30)
    for nm in ("x", "m_tau", "pl_strain"):
        assert np.array_equal(a.get(nm), b.get(nm))
    # the python restatement of the comm-block layout used by the gloo tests matches the library
    from weldformfem_b200.distributed import slot_table
    for r in a.ranks:
        tab = slot_table(r.partition.neigh_ranks, r.partition.halo_offset)
        assert tab == {q: r.slot_offsets(i)[:2] for i, q in enumerate(r.neigh)}
    a.close(); b.close()


def test_global_mesh_entry_point_matches_box(oracle_port):
    """wf_partition_build on explicit connectivity == wf_partition_build_box."""
    from weldformfem_b200.distributed import LocalCluster
    case = CASES["hex"]
    a = run_cluster(case, 2, 5)
    o = oracle_port()
    case.apply(o)
    x0 = o.get("x") - o.get("u")
    cl = LocalCluster(2, devices_for(2))
    cl.set_mesh(3, 8, x0, o.get("m_elnod"))
    cl.set_material(case.E, case.nu, case.rho0, case.model, case.sy0, case.K, case.m)
    cl.set_stab(**case.stab)
    cl.set_options(case.press, case.av[0], case.av[1], case.hexa_hg)
    cl.add_bcs(*case.bc_arrays())
    cl.allocate_bcs()
    cl.init(case.timestep)
    cl.step(5)
    for nm in ("x", "m_tau"):
        assert np.array_equal(cl.get(nm), a.get(nm))
    a.close(); cl.close()


def test_halo_timeout_is_sticky_and_surfaces_everywhere(monkeypatch):
    """A neighbour that never sends: the wait kernel gives up after WF_HALO_TIMEOUT_S, and from then on every entry point
    that reads the engine fails loudly (wf_get_array, wf_synchronize, wf_nonfinite_flag, wf_halo_status) instead of
    handing back shared-node state summed from stale partials."""
    from weldformfem_b200.distributed import LocalCluster
    from weldformfem_b200.domain import WfError
    monkeypatch.setenv("WF_HALO_TIMEOUT_S", "0.3")
    cl = LocalCluster(2, devices_for(2))
    CASES["hex"].apply(cl)              # init runs on both ranks: the exchanges of wf_init complete
    cl.step(2)
    cl.synchronize()
    r0 = cl.ranks[0]
    r0.step(1)                          # rank 1 does not step: its partials never arrive
    with pytest.raises(WfError, match="halo exchange timed out"):
        r0.synchronize()
    with pytest.raises(WfError, match="halo exchange timed out"):
        r0.get("x")
    with pytest.raises(WfError, match="halo exchange timed out"):
        r0.nonfinite_flag()
    with pytest.raises(WfError, match="halo exchange timed out"):
        r0.halo_status()
    r0.step(3)                          # sticky: later waits return at once instead of spinning 3 x timeout
    import time
    t0 = time.time()
    with pytest.raises(WfError):
        r0.synchronize()
    assert time.time() - t0 < 0.25
    cl.close()


def test_open_stepping_on_a_partitioned_mesh():
    """wf_step_open + wf_set_bc_values per step on every rank (what bench.py's e2e loop does at N > 1): prescribed values
    are given for the GLOBAL BC list on every rank, the engine keeps the rows of its own nodes; compared with the one-GPU
    engine stepping closed with the same values."""
    from weldformfem_b200.distributed import LocalCluster
    case = CASES["hex"]
    cl = LocalCluster(2, devices_for(2))
    case.apply(cl)
    one = run_single(case, 0)
    _, dims, vals = case.bc_arrays()
    vz = vals[dims == 2]
    for i in range(7):
        newv = vz * (1.0 + 0.1 * np.cos(i))
        one.set_bc_values(2, newv)
        one.step(1)
        for r in cl.ranks:
            r.set_bc_values(2, newv)
        for r in cl.ranks:
            r.step_open(1)
            r.monitor_async()
        ek = sum(r.monitor_wait()[0] for r in cl.ranks)
    for r in cl.ranks:
        r.step_close()
    worst = {nm: relerr(cl.get(nm), one.get(nm)) for nm in NAMES}
    assert max(worst.values()) < 1e-11, worst
    # kinetic energy of the last step: shared nodes are counted by every sharer, so the sum over ranks is >= the global value
    assert ek >= one.energies()[0] * (1 - 1e-12)
    cl.close(); one.close()


def test_axisymmetric_partition_needs_the_axis_on_every_rank():
    """The axis constraint uses the rank-local minimum of x_r (wf_set_axis_xmin): a rank whose block does not touch the
    axis is refused, and so is an axisymmetric partition that was never told the global minimum."""
    import ctypes as C
    from weldformfem_b200.distributed import Partition, RankDomain
    from weldformfem_b200.domain import WfError
    # a strip of 4 quads along x_r, numbered along x_r: the second half of the elements lies away from the axis
    x = np.array([[i * 0.1, j * 0.1] for j in range(2) for i in range(5)], np.float64)
    el = np.array([[i, i + 1, i + 6, i + 5] for i in range(4)], np.uint32)
    ok = RankDomain(0, 2)
    ok.setAxiSymm()
    ok.set_mesh(2, 4, x, el)            # rank 0 owns the elements at the axis
    assert ok.counts()[1] == 2
    far = RankDomain(1, 2)
    far.setAxiSymm()
    with pytest.raises(WfError, match="no node on the axis"):
        far.set_mesh(2, 4, x, el)
    raw = RankDomain(0, 2)
    raw.setAxiSymm()
    raw._create(2, 4)
    part = Partition(2, 0, mesh=(4, x.shape[0], el))
    xl = np.ascontiguousarray(x[part.node_l2g])
    rc = raw._lib.wf_set_mesh_partition(raw._h, part._h, xl.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc != 0 and b"wf_set_axis_xmin" in raw._lib.wf_last_error(raw._h)
    for d in (ok, far, raw):
        d.close()
