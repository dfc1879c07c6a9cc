"""The plain-C oracle (oracle/wf_oracle.c) is pinned: bit-for-bit against the committed fixtures that
were dumped from the unmodified reference build, and — where oracle/_ref exists — against that build
run live on larger cases.  Also the approximate pins of the reference's validation/ printouts."""
import dataclasses
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases_golden import CONTACT_ARRAYS, FLOAT_ARRAYS, GOLDEN, INT_ARRAYS, THERMAL_ARRAYS  # noqa: E402

from weldformfem_b200 import cases  # noqa: E402


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_matches_reference_fixtures_bit_for_bit(name, oracle_port):
    case, steps = GOLDEN[name]
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert int(g["steps"][0]) == steps
    d = oracle_port()
    case.apply(d)
    for nm in INT_ARRAYS:
        assert np.array_equal(d.get(nm), g[nm]), nm
    assert np.array_equal(d.get("x"), g["x0"])
    extra = (CONTACT_ARRAYS if case.contact is not None else []) + (THERMAL_ARRAYS if case.thermal is not None else [])
    for nm in extra:
        assert np.array_equal(d.get(nm), g["s0_" + nm]), ("setup", nm)
    d.step(1)
    for nm in FLOAT_ARRAYS:
        assert np.array_equal(d.get(nm), g["s1_" + nm]), ("step 1", nm)
    d.step(steps - 1)
    for nm in FLOAT_ARRAYS + list(extra):
        assert np.array_equal(d.get(nm), g["sN_" + nm]), (f"step {steps}", nm)
    if case.contact is not None:
        assert (g["sN_m_mesh_in_contact"] >= 0).sum() > 0, "fixture should end with nodes in contact"
    if "sN_m_hg_q" in g:
        assert np.array_equal(d.get("m_hg_q")[: 2 * case.n_elems], g["sN_m_hg_q"])
    d.call("calcNodalPressureFromElemental")
    assert np.array_equal(d.get("p_node"), g["sN_p_node"])
    if "sN_min_edge" in g:
        d.call("calcMinEdgeLength")
        c = d.consts()
        assert np.array_equal([c["min_length"], c["min_height"]], g["sN_min_edge"])
        assert np.array_equal(d.get("m_elem_length"), g["sN_m_elem_length"])


LIVE = [dataclasses.replace(cases.c3_hexes(9), top_vel=-150.0), dataclasses.replace(cases.c2_tets(7), top_vel=-150.0),
        dataclasses.replace(cases.c4_axisymm_quads(20), top_vel=-40.0),
        cases.contact_tets(6, stab=dict(alpha_free=0.3, alpha_contact=0.6, hg_coeff_free=0.2, hg_coeff_contact=0.1,
                                        av_coeff_div=0.15, av_coeff_bulk=0.15, log_factor=0.8, pspg_scale=0.2,
                                        p_pspg_bulkfac=0.05, J_min=0.1)),
        cases.contact_quads(12), cases.contact_quads(10, domtype=cases.AXISYMM),
        # Johnson-Cook / GMT: hexes and quads only — the reference reads the NODAL temperature array with the ELEMENT id
        # (Mechanical.C:1731), which runs past its end on tet meshes (more elements than nodes)
        cases.with_johnson_cook(dataclasses.replace(cases.c3_hexes(6), top_vel=-150.0)),
        cases.with_gmt(dataclasses.replace(cases.c3_hexes(5), top_vel=-150.0)),
        cases.with_johnson_cook(dataclasses.replace(cases.c4_axisymm_quads(12), top_vel=-40.0)),
        cases.with_thermal(dataclasses.replace(cases.c3_hexes(5), top_vel=-150.0)),
        cases.with_thermal(cases.with_johnson_cook(dataclasses.replace(cases.c3_hexes(4), top_vel=-150.0)), T0=400.0),
        cases.with_thermal(cases.contact_tets(5), heat_cond=25000.0, T_die=200.0)]


@pytest.mark.parametrize("case", LIVE, ids=lambda c: c.name)
def test_oracle_matches_compiled_reference_live(case, oracle_port, oracle_ref):
    oracle_ref.set_threads(1)
    a, b = oracle_ref(), oracle_port()
    case.apply(a)
    case.apply(b)
    a.step(40)
    b.step(40)
    extra = (CONTACT_ARRAYS if case.contact is not None else []) + (THERMAL_ARRAYS if case.thermal is not None else [])
    for nm in FLOAT_ARRAYS + ["m_f_elem", "m_str_rate", "m_rot_rate", "m_detJ", "u_dt"] + list(extra):
        assert np.array_equal(a.get(nm), b.get(nm)), nm


def test_validation_file_pins(oracle_port):
    """validation/1elem_3d_red_int_f_0.06.txt (C++ block :5-42, "NO HG" block :73-83) were printed by older
    code (incremental pressure law); the current algorithm agrees to ~1 %."""
    d = oracle_port()
    cases.c1_one_hex().apply(d)
    d.step(126)
    u = d.get("u").reshape(-1, 3)
    assert abs(u[4, 2] - (-1.008000e-03)) < 1e-12                    # prescribed top displacement, exact
    assert abs(u[1, 0] / 2.991992e-04 - 1) < 0.015                   # :7
    assert abs(u[4, 0] / -6.768332e-06 - 1) < 0.015                  # :10
    v = d.get("v").reshape(-1, 3)
    assert abs(v[1, 0] / 3.039238e-01 - 1) < 0.015                   # :16
    f = d.get("m_fi").reshape(-1, 3)
    assert abs(f[0, 2] / 5.249056e+06 - 1) < 0.015                   # :34
    c = d.consts()                                                   # :112-114
    assert abs(c["alpha"] - 0.35001649984) < 1e-10 and abs(c["beta"] - 0.65152149311) < 1e-10
    assert abs(c["gamma"] - 1.1499835002) < 1e-9
    d2 = oracle_port()
    cases.c1_one_hex(hexa_hg=0.0).apply(d2)
    d2.step(126)
    assert abs(d2.get("u").reshape(-1, 3)[1, 0] / 1.530002e-04 - 1) < 0.015   # :77


def test_hourglass_orthogonal_to_rigid_and_linear_fields(oracle_port):
    """Property of the restated hexa hourglass force (f90_ver/src/Mechanical.f90:241-344): zero for any
    velocity field that is linear in the (undistorted) coordinates."""
    case = cases.c3_hexes(3)
    d = oracle_port()
    case.apply(d)
    x = d.get("x").reshape(-1, 3)
    A = np.array([[0.3, -1.0, 2.0], [0.5, 0.7, -0.2], [1.1, 0.0, -0.9]])
    d.set("v", (x @ A.T + np.array([1.0, -2.0, 3.0])).reshape(-1))
    d.call("calcElemJAndDerivatives"); d.call("CalcElemVol"); d.call("calcElemHourglassForces")
    fh = d.get("m_f_elem_hg")
    d.call("calcElemStrainRates")
    scale = np.abs(d.get("rho")).max() * 4600.0 * 1e-6
    assert np.abs(fh).max() < 1e-9 * scale
