"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on identical inputs.

Tolerances are the ones BASELINE.json states: nodal force / velocity / displacement and element
stress / plastic strain within 1e-10 relative after one step and 1e-6 after 1000 steps; integer
artefacts bit-exact.  The strict flavour is additionally held to 1e-12 (it differs from the CPU path
only by libm pow/log rounding)."""
import dataclasses

import numpy as np
import pytest

from parity_util import STATE, compare, relerr, run_pair
from weldformfem_b200 import cases

pytestmark = pytest.mark.gpu

TOL_1STEP = 1e-10
TOL_1000 = 1e-6
TOL_STRICT = 1e-12

R = dataclasses.replace


def report(key, worst):
    """Append the measured worst relative errors to gpurun_out/parity_report.jsonl (evidence for DESIGN.md)."""
    import json, os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"case": key, "max_rel_err": max(worst.values()), "per_array": worst}) + "\n")


SMALL = {
    "hex": R(cases.c3_hexes(8), top_vel=-200.0),
    "hex_nohg": R(cases.c3_hexes(6, hexa_hg=0.0), top_vel=-200.0),
    "tet": R(cases.c2_tets(6), top_vel=-200.0),
    "tet_anp_nodal": R(cases.c2_tets(6, press=3), top_vel=-200.0),
    "axiquad": R(cases.c4_axisymm_quads(12), top_vel=-50.0),
    "psquad": R(cases.plane_strain_quads(12), top_vel=-50.0),
    "pstri": R(cases.plane_strain_tris(12), top_vel=-50.0),
    "hex_stab": R(cases.c3_hexes(6), top_vel=-200.0, av=(1.0, 0.2),
                  stab=dict(alpha_free=0.3, hg_coeff_free=0.2, av_coeff_div=0.15, av_coeff_bulk=0.15, log_factor=0.8,
                            pspg_scale=0.2, p_pspg_bulkfac=0.05, J_min=0.1)),
    "tet_stab": R(cases.c2_tets(5), top_vel=-200.0, av=(1.0, 0.2),
                  stab=dict(alpha_free=0.3, hg_coeff_free=0.2, av_coeff_div=0.15, av_coeff_bulk=0.15, log_factor=0.8,
                            pspg_scale=0.2, p_pspg_bulkfac=0.05, J_min=0.1)),
}


SMALL.update({
    # Johnson-Cook / GMT flow stress (Material.cuh:377-483) at a uniform temperature, SURVEY 8f-3
    "hex_jc": cases.with_johnson_cook(R(cases.c3_hexes(6), top_vel=-200.0)),
    "hex_gmt": cases.with_gmt(R(cases.c3_hexes(6), top_vel=-200.0)),
    "tet_jc": cases.with_johnson_cook(R(cases.c2_tets(5), top_vel=-200.0)),
    "axiquad_gmt": cases.with_gmt(R(cases.c4_axisymm_quads(12), top_vel=-50.0)),
})


def _names(case):
    n = list(STATE)
    if case.dim == 2 and not case.tritet:
        n.append("m_hg_q")
    return n


@pytest.mark.parametrize("key", sorted(SMALL))
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_one_step(key, strict, oracle_port):
    case = SMALL[key]
    eng, ref = run_pair(case, oracle_port, 1, strict)
    w = compare(eng, ref, _names(case), TOL_STRICT if strict else TOL_1STEP, f"{key} 1 step")
    report(f"1_step_{key}_{'strict' if strict else 'fast'}", w)
    for nm in ("m_elnod", "m_nodel", "m_nodel_loc", "m_nodel_offset", "m_nodel_count"):
        assert np.array_equal(eng.get(nm), ref.get(nm)), nm


@pytest.mark.parametrize("key", sorted(SMALL))
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_100_steps_plastic(key, strict, oracle_port):
    case = SMALL[key]
    eng, ref = run_pair(case, oracle_port, 100, strict)
    assert (ref.get("pl_strain") > 0).mean() > 0.5, "case should be mostly plastic"
    compare(eng, ref, _names(case), 1e-9 if strict else 1e-7, f"{key} 100 steps")
    assert not eng.nonfinite_flag()


def test_anp_as_shipped(oracle_port):
    """m_press_algorithm 1 accumulates pressure (Mechanical.C:1243-1247); reproduce it for a few steps."""
    case = R(cases.c2_tets(4, press=1), top_vel=-200.0)
    eng, ref = run_pair(case, oracle_port, 5, True)
    compare(eng, ref, STATE, 1e-11, "ANP as shipped")


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_c1_one_hexa_126_steps(strict, oracle_port):
    """configs[0]: 1-element reduced-integration hexa compression with hourglass 0.06, 126 steps."""
    case = cases.c1_one_hex()
    eng, ref = run_pair(case, oracle_port, 126, strict)
    # the acceleration is a small difference of large forces (|a| ~ 3 vs |f| ~ 5e6): looser bound
    compare(eng, ref, [n for n in STATE if n not in ("a", "prev_a")], 1e-11 if strict else 1e-9, "C1")
    compare(eng, ref, ["a", "prev_a"], 1e-7, "C1 accelerations")
    u = eng.get("u").reshape(-1, 3)
    # validation/1elem_3d_red_int_f_0.06.txt:5-14 (older pressure law, ~1 % pin): u_x(node 1) = 2.991992e-04
    assert abs(u[1, 0] - 2.991992e-04) / 2.991992e-04 < 0.02
    assert abs(u[4, 2] + 1.008e-3) < 1e-12


@pytest.mark.parametrize("key,strict", [("hex", True), ("hex", False), ("tet", False), ("axiquad", False)])
def test_1000_steps(key, strict, oracle_port):
    case = {"hex": R(cases.c3_hexes(10), top_vel=-40.0), "tet": R(cases.c2_tets(6), top_vel=-40.0),
            "axiquad": SMALL["axiquad"]}[key]
    eng, ref = run_pair(case, oracle_port, 1000, strict)
    assert (ref.get("pl_strain") > 0).mean() > 0.5
    w = compare(eng, ref, _names(case), TOL_1000, f"{key} 1000 steps")
    report(f"1000_steps_{key}_{'strict' if strict else 'fast'}", w)


def test_tracking_eps_and_sigma(oracle_port):
    case = SMALL["hex"]
    eng, ref = run_pair(case, oracle_port, 20, False, tracking=dict(eps=True, sigma=True))
    compare(eng, ref, ["m_eps", "m_sigma"], 1e-9, "tracking")


@pytest.mark.parametrize("key", ["hex", "tet", "axiquad", "pstri"])
def test_unfused_sequence_matches_oracle(key, oracle_port):
    """Drive one step through the 1:1 entry points in the order of Solver_explicit.C:524-978 and
    compare every intermediate the reference keeps."""
    case = SMALL[key]
    eng, ref = run_pair(case, oracle_port, 3, True)
    seq = [("UpdatePrediction", 0), ("ImposeBCVAllDim", 0), ("calcElemJAndDerivatives", 0)]
    if case.dim == 2 and case.domtype == cases.AXISYMM:
        seq.append(("Calc_Element_Radius", 0))
    seq += [("CalcElemVol", 0), ("CalcNodalVol", 0), ("CalcNodalMassFromVol", 0), ("calcElemStrainRates", 0),
            ("calcElemPressure", 0), ("CalcStressStrain", case.timestep), ("calcArtificialViscosity", 0),
            ("calcElemForces", 0), ("calcElemHourglassForces", 0), ("assemblyForces", 0), ("calcAccel", 0),
            ("ImposeBCAAllDim", 0), ("UpdateCorrectionAccVel", 0), ("ImposeBCVAllDim", 0), ("AxisConstraint", 0),
            ("UpdateCorrectionPos", 0)]
    for fn, arg in seq:
        eng.call(fn, arg)
        ref.call(fn, arg)
    names = _names(case) + ["m_detJ", "m_dH_detJ_dx", "m_dH_detJ_dy", "m_str_rate", "m_rot_rate", "m_f_elem",
                            "m_f_elem_hg", "u_dt"]
    if case.dim == 3:
        names.append("m_dH_detJ_dz")
    compare(eng, ref, names, TOL_STRICT, f"{key} unfused")
    # and the fused path lands on the same state
    eng2, _ = run_pair(case, oracle_port, 0, True)
    eng2.step(4)
    for nm in ("x", "v", "u", "prev_a", "m_tau", "pl_strain"):
        assert relerr(eng2.get(nm), eng.get(nm)) <= 1e-15, nm


def test_against_compiled_reference(oracle_ref):
    """Same comparison against the UNMODIFIED reference build (oracle/_ref), when it travelled to this box."""
    oracle_ref.set_threads(1)
    for key in ("hex", "tet", "axiquad"):
        case = SMALL[key]
        eng, ref = run_pair(case, oracle_ref, 20, True)
        compare(eng, ref, _names(case), 1e-11, f"{key} vs reference build")


def test_set_get_roundtrip_and_restart(oracle_port):
    """wf_get_array / wf_set_array as checkpoint: a restarted engine continues bit-identically."""
    case = SMALL["hex"]
    eng, _ = run_pair(case, oracle_port, 30, False)
    snap = {nm: eng.get(nm) for nm in ("x", "v", "u", "u_dt", "prev_a", "m_tau", "pl_strain", "p", "sigma_y")}
    eng.step(10)
    from weldformfem_b200.domain import Domain_d
    e2 = Domain_d(strict=False)
    case.apply(e2)
    for nm, arr in snap.items():
        e2.set(nm, arr)
    e2.step(10)
    for nm in snap:
        assert np.array_equal(e2.get(nm), eng.get(nm)), nm


def test_deterministic_repeat():
    from weldformfem_b200.domain import Domain_d
    case = R(cases.c3_hexes(24), top_vel=-100.0)
    outs = []
    for _ in range(2):
        e = Domain_d()
        case.apply(e)
        e.step(50)
        outs.append((e.get("x"), e.get("m_tau"), e.get("m_fi")))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_time_dependent_bc_values_and_async_monitor(oracle_port):
    """wf_set_bc_values == rewriting bcz_val between steps (Domain_d.h:901); wf_monitor_async/wait returns the
    kinetic energy of computeEnergies (Mechanical.C:2145) for the step it was enqueued after."""
    case = SMALL["hex"]
    eng, ref = run_pair(case, oracle_port, 5, False)
    _, dims, vals = case.bc_arrays()
    vz = vals[dims == 2]
    for i in range(6):
        newv = vz * (1.0 + 0.1 * (i + 1))
        eng.set_bc_values(2, newv)
        ref.set("bcz_val", newv)
        eng.step(1)
        ref.step(1)
        eng.monitor_async()
        if i >= 1:
            ek_prev, bad = eng.monitor_wait()
            assert not bad and abs(ek_prev - ek_ref_prev) <= 1e-9 * abs(ek_ref_prev)
        ek_ref_prev = ref.energies()[0]
    ek, bad = eng.monitor_wait()
    assert not bad and abs(ek - ek_ref_prev) <= 1e-9 * abs(ek_ref_prev)
    compare(eng, ref, STATE, 1e-9, "time-dependent BC values")


@pytest.mark.parametrize("key", ["hex", "tet", "axiquad"])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_device_diagnostics(key, strict, oracle_port):
    """SURVEY §8f-1 on the device: calcMinEdgeLength (Domain_d.C:2224), p_node (Mechanical.C:1187), max |v| and the
    variable step dt = cfl * min_length / (cs + max|v|) (Solver_explicit.C:579-598), then stepping with the new dt."""
    case = SMALL[key]
    eng, ref = run_pair(case, oracle_port, 40, strict)
    ref.call("calcMinEdgeLength")
    ml, mh = eng.calcMinEdgeLength()
    c = ref.consts()
    # heights of a hexa's first four (coplanar) nodes are pure cancellation noise (Domain_d.C:2243-2247 reads every 3D
    # element as a tetrahedron): compare on the scale of the edge length
    assert abs(ml - c["min_length"]) <= 1e-12 * c["min_length"]
    assert abs(mh - c["min_height"]) <= 1e-9 * c["min_length"]
    assert np.allclose(eng.get("m_elem_length"), ref.get("m_elem_length"), rtol=1e-9, atol=1e-9 * c["min_length"])
    ref.call("calcNodalPressureFromElemental")
    assert relerr(eng.get("p_node"), ref.get("p_node")) <= (1e-13 if strict else 1e-10)
    v = ref.get("v").reshape(-1, case.dim)
    vmax = float(np.sqrt((v * v).sum(axis=1)).max())
    assert abs(eng.max_velocity() - vmax) <= 1e-10 * vmax
    cs = np.sqrt(case.bulk / ref.get("rho")[0])
    dt_ref = 0.25 * c["min_length"] / (cs + vmax)
    dt = eng.cfl_dt(0.25)
    assert abs(dt - dt_ref) <= 1e-10 * dt_ref
    eng.set_dt(dt)
    ref.call("SetDT", dt)
    eng.step(10)
    ref.step(10)
    compare(eng, ref, _names(case), 1e-9 if strict else 1e-8, f"{key} after a time-step change")


@pytest.mark.parametrize("key,shuffle_elems", [("tet", True), ("tet", False), ("hex", False), ("hex", True)])
def test_renumbered_mesh_through_set_mesh(key, shuffle_elems, oracle_port):
    """Arbitrary numbering (the CreateFromLSDyna path): node ids permuted, elements optionally permuted.  The
    tile-reduced force path must cope with scattered node ids (tets: any order; hexes: shuffled elements break the
    conflict-free-rounds condition and the engine falls back to the node-ordered buffer) — either way the fast engine
    agrees with the oracle run on the SAME renumbered mesh."""
    from weldformfem_b200.domain import Domain_d
    case = SMALL[key]
    o = oracle_port()
    case.apply(o)
    x0 = (o.get("x") - o.get("u")).reshape(-1, 3)
    el = o.get("m_elnod").reshape(-1, case.nodxelem).astype(np.int64)
    rng = np.random.default_rng(1234)
    nperm = rng.permutation(len(x0))            # new id of old node i
    x1 = np.empty_like(x0)
    x1[nperm] = x0
    el1 = nperm[el]
    if shuffle_elems:
        el1 = el1[rng.permutation(len(el1))]
        if key == "hex":   # rotate every other hexahedron about its zeta axis: corners now collide inside a tile
            el1[1::2] = el1[1::2][:, [1, 2, 3, 0, 5, 6, 7, 4]]
    nodes, dims, vals = case.bc_arrays()
    nodes1 = nperm[nodes]

    def setup(dom):
        dom.set_mesh(3, case.nodxelem, x1.ravel(), el1.ravel().astype(np.int32))
        dom.set_material(case.E, case.nu, case.rho0, case.model, case.sy0, case.K, case.m)
        dom.set_stab(**case.stab)
        dom.set_options(case.press, case.av[0], case.av[1], case.hexa_hg)
        if hasattr(dom, "add_bcs"):
            dom.add_bcs(nodes1.astype(np.int32), dims, vals)
        else:
            for nd, dd, val in zip(nodes1, dims, vals):
                dom.add_bc(int(nd), int(dd), float(val))
        dom.allocate_bcs()
        dom.init(case.timestep)
        return dom

    ref = setup(oracle_port())
    eng = setup(Domain_d(strict=False))
    ref.step(60)
    eng.step(60)
    compare(eng, ref, _names(case), 1e-8, f"{key} renumbered (elements shuffled: {shuffle_elems})")
    # and the renumbered run is the original run, renumbered (the physics does not depend on numbering)
    base, _ = run_pair(case, oracle_port, 60, False)
    xb = base.get("x").reshape(-1, 3)
    assert relerr(eng.get("x").reshape(-1, 3)[nperm], xb) <= 1e-8


@pytest.mark.parametrize("key", ["tet", "hex"])
def test_remesh_handoff_new_engine_with_mapped_fields(key, oracle_port):
    """SURVEY §8f-4, second half: after a remesh the reference swaps in a new mesh and the fields mapped onto it
    (ReMesher::WriteDomain, ReMesher.C:339: u v prev_a, T; pl_strain p sigma_y tau rho vol_0) and goes on stepping.
    Here the hand-off is a fresh engine + wf_set_array of the same fields.  The 'remesh' is a renumbering of nodes and
    elements, for which the mapping is exact, so the continued run must follow the uninterrupted one."""
    from weldformfem_b200.domain import Domain_d
    case = SMALL[key]
    a, _ = run_pair(case, oracle_port, 0, False)
    a.step(40)
    k = case.nodxelem
    el = a.get("m_elnod").reshape(-1, k).astype(np.int64)
    rng = np.random.default_rng(99)
    nperm = rng.permutation(case.n_nodes)        # new id of old node i
    eperm = rng.permutation(len(el))             # new element j = old element eperm[j]

    def nodal(arr, c):
        out = np.empty_like(arr.reshape(-1, c))
        out[nperm] = arr.reshape(-1, c)
        return out.ravel()

    def elem(arr, c):
        return arr.reshape(-1, c)[eperm].ravel()

    b = Domain_d(strict=False)
    b.set_mesh(3, k, nodal(a.get("x"), 3), nperm[el][eperm].ravel().astype(np.int32))   # the CURRENT configuration
    b.set_material(case.E, case.nu, case.rho0, case.model, case.sy0, case.K, case.m)
    b.set_stab(**case.stab)
    b.set_options(case.press, case.av[0], case.av[1], case.hexa_hg)
    nodes, dims, vals = case.bc_arrays()
    b.add_bcs(nperm[nodes].astype(np.int32), dims, vals)
    b.allocate_bcs()
    b.init(case.timestep)
    b.set_time(*a.time())
    for nm, c in (("u", 3), ("v", 3), ("prev_a", 3)):
        b.set(nm, nodal(a.get(nm), c))
    for nm, c in (("m_tau", 6), ("pl_strain", 1), ("p", 1), ("sigma_y", 1), ("rho", 1), ("vol_0", 1)):
        b.set(nm, elem(a.get(nm), c))
    a.step(30)
    b.step(30)
    assert b.time()[1] == a.time()[1] == 70 and abs(b.time()[0] - a.time()[0]) <= 1e-15
    for nm, c in (("x", 3), ("v", 3), ("u", 3)):
        assert relerr(b.get(nm).reshape(-1, c)[nperm], a.get(nm).reshape(-1, c)) <= 1e-9, nm
    for nm, c in (("m_tau", 6), ("pl_strain", 1), ("p", 1)):
        assert relerr(b.get(nm).reshape(-1, c), a.get(nm).reshape(-1, c)[eperm]) <= 1e-8, nm


# ---- internal element order (Morton bricks) is invisible at the ABI ------------------------------
@pytest.mark.parametrize("key", ["hex", "tet", "axiquad", "pstri", "hex_stab"])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_elem_order_is_invisible(key, strict, oracle_port):
    """The engine keeps its element arrays in Morton order; every array crosses the ABI in the caller's numbering and
    every nodal sum keeps the caller's (= reference's) element order, so the strict flavour is BIT-identical with and
    without the reordering, and the fast flavour differs only by the association of its per-tile partial sums."""
    from weldformfem_b200.domain import Domain_d
    case = SMALL[key]
    engs = []
    for mode in (0, 1):
        e = Domain_d(strict=strict, elem_order=mode)
        case.apply(e)
        e.step(20)
        engs.append(e)
    perm0, perm1 = engs[0].get("elem_perm"), engs[1].get("elem_perm")
    assert np.array_equal(perm0, np.arange(perm0.size))
    assert np.array_equal(np.sort(perm1), np.arange(perm1.size)) and not np.array_equal(perm1, perm0)
    for nm in _names(case):
        a, b = engs[0].get(nm), engs[1].get(nm)
        if strict:
            assert np.array_equal(a, b), nm
        else:
            assert relerr(b, a) < 1e-12, (nm, relerr(b, a))
    ref = oracle_port()
    case.apply(ref)
    ref.step(20)
    compare(engs[1], ref, _names(case), 1e-11 if strict else 1e-9, f"{key} reordered vs oracle")
    # element arrays written through the ABI land on the caller's element
    tau = np.arange(perm1.size * 6, dtype=np.float64).reshape(-1, 6)
    engs[1].set("m_tau", tau.ravel())
    assert np.array_equal(engs[1].get("m_tau").reshape(-1, 6), tau)
    pl = np.arange(perm1.size, dtype=np.float64)
    engs[1].set("pl_strain", pl)
    assert np.array_equal(engs[1].get("pl_strain"), pl)
    for e in engs:
        e.close()


@pytest.mark.parametrize("dims", [(13, 11, 9), (8, 4, 4), (5, 3, 2), (17, 6, 7)])
def test_brick_plan_vs_compact_thread_slots(dims, oracle_port, monkeypatch):
    """Hexa boxes whose edge counts are no multiples of the 8x4x4 brick: the brick passes give every CTA the cells of ONE
    brick group (wf_host_brick_plan; idle threads where a cell does not exist) instead of 128 consecutive elements.  Same
    results as the compact thread slots up to the association of the per-tile partial sums, and as the oracle."""
    from weldformfem_b200.domain import Domain_d
    case = R(cases.Case("hex_%dx%dx%d" % dims, 3, dims, 1.0e-3, cfl=0.3, hexa_hg=0.06), top_vel=-200.0)
    engs = []
    for plan in ("1", "0"):
        monkeypatch.setenv("WF_BRICK_PLAN", plan)
        e = Domain_d()
        case.apply(e)
        n_cta, is_plan = e.brick_info()
        groups = -(-dims[0] // 8) * -(-dims[1] // 4) * -(-dims[2] // 4)
        ne = dims[0] * dims[1] * dims[2]
        if plan == "1":
            assert (n_cta, is_plan) == (groups, 1)
        else:   # compact slots: chunks of 128 consecutive elements, or (ragged chunks that do not fit) the generic tile kernel
            assert is_plan == 0 and n_cta in (0, -(-ne // 128))
        e.step(20)
        engs.append(e)
    for nm in _names(case):
        a, b = engs[0].get(nm), engs[1].get(nm)
        assert relerr(a, b) < 1e-12, (nm, relerr(a, b))
    ref = oracle_port()
    case.apply(ref)
    ref.step(20)
    compare(engs[0], ref, _names(case), 1e-9, "hexa brick plan vs oracle")
    for e in engs:
        e.close()


@pytest.mark.parametrize("key", ["hex", "tet", "psquad"])
def test_open_stepping_with_per_step_bc_values_and_monitor(key, oracle_port):
    """wf_step_open: a host loop that talks to the engine every step (new prescribed velocities in, kinetic energy out)
    keeps the fused schedule — the call's last node pass already runs the next predictor.  Prescribed values change
    between calls: their u_dt must use the OLD value, their velocity the NEW one (UpdatePrediction then ImposeBCV,
    Solver_explicit.C:524-540).  Compared with the oracle stepping one step at a time with the same values."""
    case = SMALL[key]
    eng, ref = run_pair(case, oracle_port, 3, False)
    _, dims, vals = case.bc_arrays()
    d = case.dim - 1
    name = "bcz_val" if d == 2 else "bcy_val"
    vd = vals[dims == d]
    ek_ref = []
    for i in range(8):
        newv = vd * (1.0 + 0.15 * np.sin(1.0 + i))
        eng.set_bc_values(d, newv)
        ref.set(name, newv)
        eng.step_open(1)
        ref.step(1)
        eng.monitor_async()
        ek_ref.append(ref.energies()[0])
        if i >= 1:
            ek, bad = eng.monitor_wait()
            assert not bad and abs(ek - ek_ref[i - 1]) <= 1e-9 * abs(ek_ref[i - 1])
    ek, bad = eng.monitor_wait()
    assert not bad and abs(ek - ek_ref[-1]) <= 1e-9 * abs(ek_ref[-1])
    with pytest.raises(Exception, match="mid-batch"):
        eng.get("v")
    eng.step_close()
    compare(eng, ref, [n for n in STATE], 1e-9, f"{key} open stepping")
    with pytest.raises(Exception):
        eng.get("u_dt")                      # the fused schedule does not keep the last increment
    # a plain step closes too, and u_dt is back
    eng2, ref2 = run_pair(case, oracle_port, 2, False)
    eng2.step_open(3)
    eng2.step(1)
    ref2.step(4)
    compare(eng2, ref2, STATE + ["u_dt"], 1e-9, f"{key} open then closed")


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_many_cta_waves_against_the_compiled_reference(strict, oracle_ref):
    """262 144 hexes (64^3: 2 048 CTAs of the element passes = several waves on 148 SMs, ragged Morton bricks at the mesh
    faces, tail tiles) with hourglass 0.06, 50 steps into the plastic range, against the reference compiled from its own
    sources (oracle/_ref, all host threads) — the brick form of E1 / E2, the tile partials and the node passes on a mesh
    that does not fit one wave."""
    case = R(cases.c3_hexes(64), top_vel=-200.0)
    oracle_ref.set_threads(0)
    eng, ref = run_pair(case, oracle_ref, 0, strict)
    ref.step(1)
    eng.step(1)
    compare(eng, ref, STATE, TOL_STRICT if strict else TOL_1STEP, "64^3 hexes, 1 step")
    ref.step(49)
    eng.step(49)
    assert (ref.get("pl_strain") > 0).mean() > 0.05
    w = compare(eng, ref, STATE, 1e-9 if strict else 1e-7, "64^3 hexes, 50 steps")
    report(f"hex64_50_steps_{'strict' if strict else 'fast'}", w)
    assert not eng.nonfinite_flag()


def test_full_size_properties_10m_hexes():
    """configs[2] at its full size (215^3 = 9 938 375 hexes, ~77 600 CTAs per element pass), where the oracle cannot
    follow: size-independent properties of the step.
      1. the caller's element order and the engine's brick order give the same state (1e-12: association of the tile sums);
      2. the step is deterministic: a second run in brick order is BIT-identical (gathers only, no atomics);
      3. a rigid translation produces no internal force and moves every node by v dt.
    (Mirror symmetries are not a property of this algorithm: the reference's tensor product in the Jaumann terms,
    Tensor3.C:290-304, treats the axes differently, and the engine reproduces that.)"""
    from weldformfem_b200.domain import Domain_d
    n = 215
    case = R(cases.c3_hexes(n), top_vel=-200.0)
    states = []
    for mode in (1, 0, 1):
        e = Domain_d(elem_order=mode)
        case.apply(e)
        e.step(20)
        states.append({nm: e.get(nm) for nm in ("u", "v", "m_tau", "pl_strain", "p")})
        assert not e.nonfinite_flag()
        e.close()
    a, b, c = states
    for nm in a:
        assert relerr(a[nm], b[nm]) < 1e-12, (nm, relerr(a[nm], b[nm]))
        assert np.array_equal(a[nm], c[nm]), nm
    assert (a["pl_strain"] > 0).mean() > 0.01
    del states, a, b, c
    # rigid translation, no boundary conditions
    e = Domain_d()
    e.box((0.0, 0.0, 0.0), [n * case.h * (1 + 1e-6)] * 3, 0.5 * case.h, False)
    e.set_material(case.E, case.nu, case.rho0, case.model, case.sy0, case.K, case.m)
    e.set_stab()
    e.set_options(0, 0.0, 0.0, case.hexa_hg)
    e.allocate_bcs()
    e.init(case.timestep)
    nn = e.counts()[0]
    x0 = e.get("x")
    v0 = np.tile(np.array([3.0, -2.0, 5.0]), nn)
    e.set("v", v0)
    e.step(3)
    f = e.get("m_fi")
    scale = case.E * case.h * case.h                 # force scale of a unit strain on one element face
    assert np.abs(f).max() < 1e-9 * scale, np.abs(f).max()
    assert np.abs(e.get("x") - x0 - 3 * case.timestep * v0).max() < 1e-12
    assert (e.get("pl_strain") == 0).all()
    e.close()
