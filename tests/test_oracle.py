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

from parity_util import relerr  # noqa: E402
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


def _validation_pins():
    import json
    return json.load(open(os.path.join(HERE, "golden", "validation_pins.json")))


def _historical_run(oracle_port, nsteps, hexa_hg=0.06, case=None):
    """The step sequence the validation/ files were printed with, driven member by member: the historical
    incremental pressure law (test-only press_algorithm 2 of the port = the commented-out calcElemPressure_Hex,
    Mechanical.C:576-603 = f90_ver/src/Mechanical.f90:525-547) and calcElemDensity EVERY step (the sequence of the
    reference's CUDA branch, Solver_explicit.C:601-608; its CPU branch freezes rho after init), everything else the
    current functions in the order of SolveChungHulbert."""
    case = case or dataclasses.replace(cases.c1_one_hex(hexa_hg=hexa_hg), press=2)
    d = oracle_port()
    case.apply(d)
    for _ in range(nsteps):
        d.call("UpdatePrediction"); d.call("ImposeBCVAllDim")
        d.call("calcElemJAndDerivatives"); d.call("CalcElemVol"); d.call("calcElemDensity")
        d.call("CalcNodalVol"); d.call("CalcNodalMassFromVol")
        d.call("calcElemStrainRates"); d.call("calcElemPressure"); d.call("CalcStressStrain", case.timestep)
        d.call("calcElemForces"); d.call("calcElemHourglassForces"); d.call("assemblyForces")
        d.call("calcAccel"); d.call("ImposeBCAAllDim"); d.call("UpdateCorrectionAccVel"); d.call("ImposeBCVAllDim")
        d.call("UpdateCorrectionPos")
    return d


def _printed_equal(got, printed, digits=7):
    """every number equals the printed one after rounding to the printed number of significant digits (%.6e)"""
    got, printed = np.asarray(got, dtype=np.float64), np.asarray(printed, dtype=np.float64)
    assert got.shape == printed.shape
    for g, w in zip(got.ravel(), printed.ravel()):
        if abs(w) < 1e-300:
            assert abs(g) < 1e-300, (g, w)
        else:
            assert float(f"{g:.{digits - 1}e}") == w or abs(g - w) <= 1.0000001 * 10.0 ** (np.floor(np.log10(abs(w))) - digits + 1), (g, w)


@pytest.mark.parametrize("block,hexa_hg", [("cxx_hg_0.06", 0.06), ("cxx_no_hg", 0.0)])
def test_validation_1elem_every_printed_digit(block, hexa_hg, oracle_port):
    """validation/1elem_3d_red_int_f_0.06.txt, C++ blocks (:5-42 with hourglass 0.06, :73-108 without), 126 steps
    of dt = 0.8e-5: displacements, velocities, accelerations AND forces agree with every printed digit (7 significant:
    <= 5e-7 relative per entry; the accelerations are differences of forces 1e6 times larger).  This pins the restated
    hexa viscous hourglass (f90_ver/src/Mechanical.f90:241-344: sign table, vol^0.6666666, rho, 0.25, cs0, the
    subtraction in assemblyForces) and the step sequence; a15b of SURVEY.md 8(a)."""
    pins = _validation_pins()[block]
    d = _historical_run(oracle_port, 126, hexa_hg)
    _printed_equal(d.get("u").reshape(-1, 3), pins["DISPLACEMENTS"])
    _printed_equal(d.get("v").reshape(-1, 3), pins["VELOCITIES"])
    acc = np.array(pins["ACCEL"])
    acc[np.abs(acc) < 1e-300] = 0.0          # node 0 prints denormal garbage (9.88e-324) on its constrained components
    _printed_equal(d.get("a").reshape(-1, 3), acc)
    _printed_equal(d.get("m_fi").reshape(-1, 3), pins["FORCES"])
    # a wrong exponent (2/3 instead of 0.6666666) or coefficient would not survive: the lateral displacement of the
    # top nodes exists only through the hourglass force
    if hexa_hg:
        assert abs(d.get("u").reshape(-1, 3)[4, 0] / -6.768332e-06 - 1) < 1e-6


def test_validation_f90_blocks(oracle_port):
    """The F90 program's own output (1elem_3d_red_int_f_0.06.txt:44-68, 'several dts.txt'): it carries single-precision
    contamination of its constants (M diag 0.98125004386 for 0.98125), so it agrees to ~1e-6 of the array maximum, not
    to the last digit.  Tolerance 1e-5 relative to max |array| (the metric of BASELINE.json)."""
    pins = _validation_pins()
    for key, nsteps in (("f90_10_steps", 10), ("f90_100_steps", 100), ("f90_126_steps", 126)):
        d = _historical_run(oracle_port, nsteps)
        want = np.array(pins[key]["Disp"])
        assert relerr(d.get("u").reshape(-1, 3), want) < 1e-5, (key, relerr(d.get("u").reshape(-1, 3), want))
        if "Vel" in pins[key]:
            assert relerr(d.get("v").reshape(-1, 3), np.array(pins[key]["Vel"])) < 1e-5


def test_validation_f90_8_elements_one_step(oracle_port):
    """validation/4elem_red_0.06_f90_1step.txt: 2x2x2 hexes of 0.05, bottom clamped, top v_z = -1, one step of
    dt = 0.8e-5/4: stress of the loaded elements (the F90 nodal mass is total mass / node count, so its accelerations
    are not this algorithm's; stress does not depend on the mass in the first step)."""
    pins = _validation_pins()["f90_8elem_1step"]

    class C8(cases.Case):
        def bc_nodes(self):
            out = []
            for n in range(27):
                if n // 9 == 0:
                    out += [(n, 0, 0.0), (n, 1, 0.0), (n, 2, 0.0)]
                if n // 9 == 2:
                    out.append((n, 2, -1.0))
            return out
    c = C8("c8", 3, (2, 2, 2), 0.05, E=206e9, nu=0.3, rho0=7850.0, model=cases.BILINEAR, sy0=1e10, K=0.0, m=1.0,
           dt=2e-6, hexa_hg=0.06, press=2)
    d = _historical_run(oracle_port, 1, case=c)
    sig = d.get("m_sigma").reshape(-1, 6)[4:]     # the four elements under the moving face
    tau = d.get("m_tau").reshape(-1, 6)[4:]
    assert np.allclose(sig[:, 0], pins["sigma_xx"], rtol=1e-7) and np.allclose(sig[:, 2], pins["sigma_zz"], rtol=1e-7)
    assert np.allclose(tau[:, 0], pins["tau_xx"], rtol=1e-7) and np.allclose(tau[:, 2], pins["tau_zz"], rtol=1e-7)


def test_validation_f90_one_element_second_step(oracle_port):
    """validation/2step_1elem_red_no_hg_DIV.txt: the F90 program's state after the SECOND step of the one-element
    compression (17 printed digits).  Intermediates of the run without hourglass forces: dH x detJ of the deformed element
    to 1e-11 and the strain-rate tensor to 5e-9 (xx = yy 4.4720252..., zz -10.0008000640..., xz = yz -0.745397...); state of
    the run with hourglass 0.06: displacements 2e-6, velocities 5e-6, accelerations 5e-5 of the array maximum (the F90
    constants carry single-precision contamination, e.g. dt = float32(0.8e-5))."""
    pins = _validation_pins()["f90_1elem_2_steps"]
    d = _historical_run(oracle_port, 2, 0.0)
    for c, key in (("x", "dHx_detJ"), ("y", "dHy_detJ")):
        got = d.get(f"m_dH_detJ_d{c}").reshape(-1, 8)[0]
        assert np.abs(got / np.array(pins[key]) - 1).max() < 1e-11, key
    sr = d.get("m_str_rate").reshape(-1, 6)[0]          # xx yy zz xy yz xz
    want = np.array(pins["strain_rate"])
    got = np.array([[sr[0], sr[3], sr[5]], [sr[3], sr[1], sr[4]], [sr[5], sr[4], sr[2]]])
    assert np.abs(got - want).max() < 5e-9 * np.abs(want).max()
    assert relerr(d.get("u").reshape(-1, 3), np.array(pins["Disp_no_hg"])) < 2e-6
    d = _historical_run(oracle_port, 2, 0.06)
    assert relerr(d.get("u").reshape(-1, 3), np.array(pins["Disp"])) < 2e-6
    assert relerr(d.get("v").reshape(-1, 3), np.array(pins["Vel"])) < 5e-6
    assert relerr(d.get("a").reshape(-1, 3), np.array(pins["Acc"])) < 5e-5
    # the two runs differ where the hourglass force acts: lateral displacement of the bottom node on the x axis
    assert abs(np.array(pins["Disp"])[1, 0] / np.array(pins["Disp_no_hg"])[1, 0] - 1) > 0.05


def test_validation_cxx_second_step_and_hourglass_force_values(oracle_port):
    """validation/2step_1elem_red_no_hg_DIV.txt:93-141, "C++ with hourglass": displacements, velocities, accelerations and
    forces after two steps to every printed digit, and the HOURGLASS FORCE of each element node itself — printed with six
    decimals, i.e. ten significant digits (3283.425868): the direct pin of calcElemHourglassForces' value, sign pattern and
    coefficient."""
    pins = _validation_pins()["cxx_hg_0.06_2_steps"]
    d = _historical_run(oracle_port, 2, 0.06)
    _printed_equal(d.get("u").reshape(-1, 3), pins["DISPLACEMENTS"])
    _printed_equal(d.get("v").reshape(-1, 3), pins["VELOCITIES"])
    _printed_equal(d.get("a").reshape(-1, 3), pins["ACCEL"])
    _printed_equal(d.get("m_fi").reshape(-1, 3), pins["FORCES"])
    hg = d.get("m_f_elem_hg").reshape(-1, 3)
    want = np.array(pins["HG_FORCES"])
    assert np.abs(hg - want).max() <= 5.0e-7 and np.abs(want).max() > 3283.0       # half a unit of the last printed decimal


def test_validation_integration_constants(oracle_port):
    """validation/cxx/2time_step.txt:61,79-81: the Chung-Hulbert alpha / beta / gamma (rho_b = 0.8182) and the sound speed
    the C++ solver printed for the one-element case (six decimals / six digits)."""
    pins = _validation_pins()["cxx_log_constants"]
    c = cases.c1_one_hex()
    d = oracle_port()
    c.apply(d)
    k = d.consts()
    for nm in ("alpha", "beta", "gamma"):
        assert abs(k[nm] - pins[nm]) <= 5.0e-7, nm
    cs0 = np.sqrt(c.E / (3.0 * (1.0 - 2.0 * c.nu)) / c.rho0)
    assert abs(cs0 - pins["CS_0"]) <= 5.0e-3


def test_validation_first_step_intermediates(oracle_port):
    """The FIRST step of the one-element compression with hourglass 0.06 (the hourglass force is still zero: no hourglass
    velocity yet).  validation/1step_red_int_cube3D_hf_c_0.06.txt (F90, 17 digits): element stress, deviatoric stress,
    pressure, global forces, accelerations, velocities, displacements to 2e-7 relative (its constants carry
    single-precision contamination: M 0.98125004386, p 13733333.91 for 13733333.33).  validation/cxx/2time_step.txt:377-433
    (the C++ solver's own log): corrected acceleration 37267.745852 to all eleven printed digits, velocity 0.342858."""
    pins = _validation_pins()
    f90, cxx = pins["f90_1elem_first_step"], pins["cxx_log_first_step"]
    d = _historical_run(oracle_port, 1, 0.06)
    sig, tau = d.get("m_sigma").reshape(-1, 6)[0], d.get("m_tau").reshape(-1, 6)[0]
    assert np.abs(sig[:3] / np.array(f90["sigma_diag"]) - 1).max() < 2e-7 and not sig[3:].any()
    assert np.abs(tau[:3] / np.array(f90["tau_diag"]) - 1).max() < 2e-7
    assert abs(d.get("p")[0] / f90["pressure"] - 1) < 2e-7
    # the F90 prints the internal force (+), rows in node order
    assert relerr(d.get("m_fi").reshape(-1, 3), np.array(f90["forces"])) < 2e-7
    assert relerr(d.get("a").reshape(-1, 3), np.array(f90["Acc"])) < 2e-7
    assert relerr(d.get("v").reshape(-1, 3), np.array(f90["Vel"])) < 2e-7
    assert relerr(d.get("u").reshape(-1, 3), np.array(f90["Disp"])) < 2e-7
    assert np.abs(d.get("a").reshape(-1, 3) - np.array(cxx["Acc"])).max() <= 5.0e-7
    assert np.abs(d.get("v").reshape(-1, 3) - np.array(cxx["Vel"])).max() <= 5.0e-7
    assert not d.get("m_f_elem_hg").any()


class _F90Cube8(cases.Case):
    """2x2x2 hexes of 0.05: symmetry conditions on the bottom layer only (z = 0: u_z = 0; its x = 0 nodes u_x = 0, its
    y = 0 nodes u_y = 0), top layer v_z = -1 — the conditions visible in validation/4_el_hg_1e-3.txt"""

    def bc_nodes(self):
        out = []
        for n in range(27):
            i, j, k = n % 3, (n // 3) % 3, n // 9
            if k == 0:
                out.append((n, 2, 0.0))
                if i == 0:
                    out.append((n, 0, 0.0))
                if j == 0:
                    out.append((n, 1, 0.0))
            if k == 2:
                out.append((n, 2, -1.0))
        return out


def _f90_cube8_run(oracle_port, nsteps, hexa_hg=0.06, press=2, equal_masses=True):
    c = _F90Cube8("c8", 3, (2, 2, 2), 0.05, E=206e9, nu=0.3, rho0=7850.0, model=cases.BILINEAR, sy0=1e10, K=0.0, m=1.0,
                  dt=2e-6, hexa_hg=hexa_hg, press=press)
    d = oracle_port()
    c.apply(d)
    for _ in range(nsteps):
        d.call("UpdatePrediction"); d.call("ImposeBCVAllDim")
        d.call("calcElemJAndDerivatives"); d.call("CalcElemVol"); d.call("calcElemDensity")
        d.call("CalcNodalVol"); d.call("CalcNodalMassFromVol")
        if equal_masses:      # the F90 program lumps total mass / node count ("Affecting masses to equal")
            m = d.get("m_mdiag")
            d.set("m_mdiag", np.full(m.size, m.sum() / m.size))
        d.call("calcElemStrainRates"); d.call("calcElemPressure"); d.call("CalcStressStrain", c.timestep)
        d.call("calcElemForces"); d.call("calcElemHourglassForces"); d.call("assemblyForces")
        d.call("calcAccel"); d.call("ImposeBCAAllDim"); d.call("UpdateCorrectionAccVel"); d.call("ImposeBCVAllDim")
        d.call("UpdateCorrectionPos")
    return d


def test_validation_f90_8_elements_501_steps(oracle_port):
    """validation/4_el_hg_1e-3.txt, F90 block: eight elements, hourglass 0.06, run to t = 1.002e-3 (501 steps of 2e-6; the
    printed top displacement 1.00199999747e-3 = 501 x float32(2e-6)).  With the F90 program's mass lumping (total mass /
    node count) the restated step reproduces its 27 nodal displacements to 3e-5 and velocities to 2e-4 of the array
    maximum — hourglass modes of NEIGHBOURING elements interacting over 500 steps, which the one-element files cannot
    show.  Negative controls: without hourglass forces 20 % off, with the current pressure law 1.2e-3, with the C++ nodal
    masses 7e-4 (lateral displacements 2.5e-3).  (The file's C++ block was printed by an intermediate, x/y-asymmetric state
    of the C++ code and is not reproducible by the algorithm at this commit.)"""
    pins = _validation_pins()["f90_8elem_501_steps"]
    want_u, want_v = np.array(pins["Disp"]), np.array(pins["Vel"])
    d = _f90_cube8_run(oracle_port, 501)
    u, v = d.get("u").reshape(-1, 3), d.get("v").reshape(-1, 3)
    assert relerr(u, want_u) < 3e-5, relerr(u, want_u)
    assert relerr(v, want_v) < 2e-4, relerr(v, want_v)
    assert relerr(u[:, :2], want_u[:, :2]) < 1e-4      # the lateral field alone (it exists through Poisson + hourglass)
    u_nohg = _f90_cube8_run(oracle_port, 501, hexa_hg=0.0).get("u").reshape(-1, 3)
    assert relerr(u_nohg, want_u) > 0.1
    # validation/4_el_NO_hg_1e-3.txt: the same run without hourglass forces (its free hourglass modes amplify the
    # single-precision contamination of the F90 constants: 2e-4)
    assert relerr(u_nohg, np.array(_validation_pins()["f90_8elem_501_steps_no_hg"]["Disp"])) < 1e-3
    assert relerr(_f90_cube8_run(oracle_port, 501, press=0).get("u").reshape(-1, 3), want_u) > 5e-4
    assert relerr(_f90_cube8_run(oracle_port, 501, equal_masses=False).get("u").reshape(-1, 3), want_u) > 3e-4


def test_validation_shape_derivative_matrix(oracle_port):
    """validation/1step_red_int_cube3D_hf_c_0.06.txt:6-9: dHdx * detJ of the 0.1 cube (+-3.125e-4, signs per node)."""
    want = np.array(_validation_pins()["dHdx_detJ"])
    d = oracle_port()
    cases.c1_one_hex().apply(d)
    d.call("calcElemJAndDerivatives")
    got = np.stack([d.get(f"m_dH_detJ_d{c}").reshape(-1, 8)[0] for c in "xyz"])
    assert np.array_equal(np.sign(got), np.sign(want))
    assert np.abs(got / want - 1).max() < 1e-6      # the box is padded by 1e-6 (cases.Case.apply); F90 prints ...0006E-004


def test_validation_fixture_is_the_reference_file():
    """the committed pins are a transcription of the reference's files (checked whenever the reference tree is here)"""
    if not os.path.isdir("/root/reference/validation"):
        pytest.skip("reference tree not present")
    import json, subprocess, sys, tempfile
    before = _validation_pins()
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_validation_golden as mk
    out = mk.OUT
    try:
        with tempfile.TemporaryDirectory() as td:
            mk.OUT = os.path.join(td, "pins.json")
            mk.main()
            assert json.load(open(mk.OUT)) == before
    finally:
        mk.OUT = out


def test_current_pressure_law_differs_from_the_validation_files(oracle_port):
    """With the pressure law the reference ships today the same run is ~1 % off the printed numbers:
    documents why the tight pin needs the historical law."""
    d = oracle_port()
    cases.c1_one_hex().apply(d)
    d.step(126)
    u = d.get("u").reshape(-1, 3)
    assert abs(u[4, 2] - (-1.008000e-03)) < 1e-12
    assert 1e-3 < abs(u[1, 0] / 2.991992e-04 - 1) < 0.015
    c = d.consts()                                                   # :112-114
    assert abs(c["alpha"] - 0.35001649984) < 1e-10 and abs(c["beta"] - 0.65152149311) < 1e-10
    assert abs(c["gamma"] - 1.1499835002) < 1e-9


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
