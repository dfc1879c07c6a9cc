"""GPU parity of the penalty contact with rigid tool surfaces: the CUDA engine through the C ABI
against the CPU oracle (bit-identical to the compiled reference, tests/test_oracle.py) on the same inputs.

Reference: Domain_d::SearchExtNodes / CalcExtFaceAreas (Domain_d.C:110-315), CalcContactForces (Contact.C:31-336),
TriMesh_d (Mesh.h / Mesh.C), loop positions Solver_explicit.C:445-450, 769-770, 981-1005.
Tolerances as BASELINE.json: 1e-10 relative after one step (1e-12 for the strict flavour), looser after many steps
because contact makes the response non-smooth (a node entering / leaving a facet one step earlier changes forces)."""
import dataclasses

import numpy as np
import pytest

from parity_util import STATE, compare, relerr, run_pair
from weldformfem_b200 import cases

pytestmark = pytest.mark.gpu
R = dataclasses.replace

STAB = dict(alpha_free=0.3, alpha_contact=0.6, hg_coeff_free=0.2, hg_coeff_contact=0.1, av_coeff_div=0.15,
            av_coeff_bulk=0.15, log_factor=0.8, pspg_scale=0.2, p_pspg_bulkfac=0.05, J_min=0.1)
CASES = {
    "tet": cases.contact_tets(6),
    "tet_stab": cases.contact_tets(5, stab=STAB),
    "tet_frictionless": cases.contact_tets(5, mu=(0.0, 0.0), two_planes=False),
    "quad": cases.contact_quads(12),
    "axiquad": cases.contact_quads(10, domtype=cases.AXISYMM),
}
CONTACT = "contforce ut_prev node_area m_elem_area".split()
TRIMESH = "trimesh.node trimesh.node_v trimesh.normal trimesh.pplane".split()


def _names(case):
    n = list(STATE) + CONTACT
    if case.dim == 2:
        n.append("m_hg_q")
    return n


@pytest.mark.parametrize("key", sorted(CASES))
def test_setup_artefacts(key, oracle_port):
    """ext_nodes / m_mesh_in_contact bit-exact; nodal areas, rigid-surface arrays, m_elem_length after setup."""
    case = CASES[key]
    eng, ref = run_pair(case, oracle_port, 0, True)
    assert np.array_equal(eng.get("ext_nodes"), ref.get("ext_nodes"))
    assert np.array_equal(eng.get("m_mesh_in_contact"), ref.get("m_mesh_in_contact"))
    assert eng.trimesh_counts() == ref.trimesh_counts()
    for nm in ["node_area", "m_elem_area"] + TRIMESH:
        assert np.array_equal(eng.get(nm), ref.get(nm)), nm
    assert relerr(eng.get("m_elem_length"), ref.get("m_elem_length")) < 1e-12


@pytest.mark.parametrize("key", sorted(CASES))
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_contact_steps(key, strict, oracle_port):
    case = CASES[key]
    eng, ref = run_pair(case, oracle_port, 0, strict)
    tol1 = 1e-12 if strict else 1e-10
    first = None
    for s in range(1, 41):                      # the tool touches the body after a few steps
        eng.step(1)
        ref.step(1)
        mic = ref.get("m_mesh_in_contact")
        assert np.array_equal(eng.get("m_mesh_in_contact"), mic), f"step {s}"
        if first is None and (mic >= 0).any():
            first = s
            compare(eng, ref, _names(case) + TRIMESH, tol1 * 10, f"{key} first contact step {s}")
        if s == 1:
            compare(eng, ref, _names(case) + TRIMESH, tol1, f"{key} step 1")
    assert first is not None, "the tool never reached the body"
    compare(eng, ref, _names(case) + TRIMESH, 1e-9 if strict else 1e-8, f"{key} 40 steps")
    eng.step(60)
    ref.step(60)
    assert (ref.get("m_mesh_in_contact") >= 0).sum() > 0 and (ref.get("pl_strain") > 0).mean() > 0.3
    assert np.array_equal(eng.get("m_mesh_in_contact"), ref.get("m_mesh_in_contact"))
    compare(eng, ref, _names(case) + TRIMESH, 1e-8 if strict else 1e-6, f"{key} 100 steps")
    assert not eng.nonfinite_flag()


def test_contact_friction_slides_and_sticks(oracle_port):
    """The case must exercise both friction branches (Contact.C:282-298), including the reference's reset of the
    accumulated slip, or the parity above proves little."""
    ref = oracle_port()
    CASES["quad"].apply(ref)
    seen_reset = seen_stick = False
    for _ in range(60):
        ref.step(1)
        hit = ref.get("m_mesh_in_contact") >= 0
        ut = ref.get("ut_prev").reshape(-1, 2)
        if hit.any():
            seen_reset |= bool((np.abs(ut[hit]).sum(axis=1) == 0).any())
            seen_stick |= bool((np.abs(ut[hit]).sum(axis=1) > 0).any())
    assert seen_reset and seen_stick


def test_contact_unfused_sequence(oracle_port):
    """One step through the 1:1 entry points in the order of Solver_explicit.C:445-1005 with contact on."""
    case = CASES["tet"]
    eng, ref = run_pair(case, oracle_port, 12, True)
    assert (ref.get("m_mesh_in_contact") >= 0).any()
    seq = [("CalcExtFaceAreas", 0), ("UpdatePrediction", 0), ("ImposeBCVAllDim", 0), ("calcElemJAndDerivatives", 0),
           ("CalcElemVol", 0), ("CalcNodalVol", 0), ("CalcNodalMassFromVol", 0), ("calcElemStrainRates", 0),
           ("calcElemPressure", 0), ("CalcStressStrain", case.timestep), ("calcArtificialViscosity", 0),
           ("calcElemForces", 0), ("calcElemHourglassForces", 0), ("CalcContactForces", 0), ("assemblyForces", 0),
           ("calcAccel", 0), ("ImposeBCAAllDim", 0), ("UpdateCorrectionAccVel", 0), ("ImposeBCVAllDim", 0),
           ("UpdateCorrectionPos", 0), ("MoveTriMesh", 0)]
    for fn, arg in seq:
        eng.call(fn, arg)
        ref.call(fn, arg)
    compare(eng, ref, _names(case) + TRIMESH, 1e-11, "contact unfused")


def test_contact_against_compiled_reference(oracle_ref):
    oracle_ref.set_threads(1)
    for key in ("tet_stab", "quad"):
        case = CASES[key]
        eng, ref = run_pair(case, oracle_ref, 30, True)
        assert np.array_equal(eng.get("m_mesh_in_contact"), ref.get("m_mesh_in_contact"))
        compare(eng, ref, _names(case) + TRIMESH, 1e-9, f"{key} vs reference build")


def test_contact_refused_where_the_reference_has_no_face_tables():
    from weldformfem_b200.domain import Domain_d, WfError
    e = Domain_d()
    R(cases.c3_hexes(4)).apply(e, init=False)
    with pytest.raises(WfError):
        e.SearchExtNodes()


def test_large_contact_case_runs_and_matches_small_step_count(oracle_port):
    """~100k tets (configs[1] size) with two 20x20 rigid planes like Contact_Compression_tetra.json: 30 steps, compared
    with the oracle on the nodal and element state."""
    case = cases.contact_tets(26)
    case = R(case, planes=tuple(dict(p, dens=20) for p in case.planes))
    eng, ref = run_pair(case, oracle_port, 30, False)
    assert (ref.get("m_mesh_in_contact") >= 0).sum() > 500
    assert np.array_equal(eng.get("m_mesh_in_contact"), ref.get("m_mesh_in_contact"))
    compare(eng, ref, _names(case), 1e-8, "contact 105k tets, 30 steps")


# ---- thermal coupling: Thermal.C fused into the element / node passes -------------------------------
THERMAL = {
    "hex_th": cases.with_thermal(R(cases.c3_hexes(6), top_vel=-200.0)),
    "tet_th": cases.with_thermal(R(cases.c2_tets(5), top_vel=-200.0)),
    "axiquad_th": cases.with_thermal(R(cases.c4_axisymm_quads(12), top_vel=-50.0)),
    "hex_jc_th": cases.with_thermal(cases.with_johnson_cook(R(cases.c3_hexes(5), top_vel=-200.0)), T0=400.0),
    "contact_tet_th": cases.with_thermal(cases.contact_tets(5), heat_cond=25000.0, T_die=200.0),
    "contact_quad_th": cases.with_thermal(cases.contact_quads(10), heat_cond=25000.0, T_die=200.0),
}


@pytest.mark.parametrize("key", sorted(THERMAL))
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_thermal_coupling(key, strict, oracle_port):
    """Conduction + plastic heating + thermal expansion (+ the Johnson-Cook T[e] read, + contact heat exchange): one step
    to 1e-10 / 1e-12, 100 steps within the many-step tolerance."""
    case = THERMAL[key]
    eng, ref = run_pair(case, oracle_port, 0, strict)
    # a temperature gradient from the start, so that conduction is more than cancellation noise in step 1
    x = ref.get("x").reshape(-1, case.dim)
    T0 = case.thermal["T0"] * (1.0 + 0.5 * x[:, -1] / x[:, -1].max() + 0.2 * x[:, 0] / x[:, 0].max())
    eng.set("T", T0)
    ref.set("T", T0)
    eng.step(1)
    ref.step(1)
    names = list(STATE) + ["T", "m_dTedt", "m_q_plheat"]
    if case.contact is not None:
        names += CONTACT + ["q_cont_conv"]
    compare(eng, ref, names, 1e-12 if strict else 1e-10, f"{key} 1 step")
    eng.step(99)
    ref.step(99)
    T = ref.get("T")
    assert T.max() > 1.5 * case.thermal["T0"] and (ref.get("m_q_plheat") > 0).any(), "case should heat up"
    if case.contact is not None:
        assert (ref.get("q_cont_conv") != 0).any()
    compare(eng, ref, names, 1e-8 if strict else 1e-6, f"{key} 100 steps")
    assert not eng.nonfinite_flag()
