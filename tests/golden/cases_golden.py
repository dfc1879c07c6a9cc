"""The small cases whose reference outputs are committed under tests/golden/ (shared by the generator
and by the tests)."""
import dataclasses

from weldformfem_b200 import cases

R = dataclasses.replace
STAB = dict(alpha_free=0.3, hg_coeff_free=0.2, av_coeff_div=0.15, av_coeff_bulk=0.15, log_factor=0.8,
            pspg_scale=0.2, p_pspg_bulkfac=0.05, J_min=0.1)

GOLDEN = {
    # name: (case, steps)
    "c1_1hex_hg006": (cases.c1_one_hex(), 126),
    "c1_1hex_nohg": (cases.c1_one_hex(hexa_hg=0.0), 126),
    "hex_n4_plastic": (R(cases.c3_hexes(4), top_vel=-200.0), 80),
    "tet_n3_plastic": (R(cases.c2_tets(3), top_vel=-200.0), 80),
    "tet_n3_anp_nodal": (R(cases.c2_tets(3, press=3), top_vel=-200.0), 80),
    "tet_n2_anp_shipped": (R(cases.c2_tets(2, press=1), top_vel=-200.0), 5),
    "axiquad_n6": (R(cases.c4_axisymm_quads(6), top_vel=-50.0), 80),
    "psquad_n6": (R(cases.plane_strain_quads(6), top_vel=-50.0), 80),
    "pstri_n6": (R(cases.plane_strain_tris(6), top_vel=-50.0), 80),
    "hex_n3_stab_av": (R(cases.c3_hexes(3), top_vel=-200.0, stab=STAB, av=(1.0, 0.2)), 60),
    # penalty contact with rigid surfaces: two planes + friction + contact-dependent stabilisation
    "contact_tet_n3": (cases.contact_tets(3, stab=dict(STAB, alpha_contact=0.6, hg_coeff_contact=0.1)), 80),
    "contact_quad_n6": (cases.contact_quads(6), 80),
    # rate-dependent flow stresses: Johnson-Cook and GMT at a uniform temperature
    "hex_n3_jc": (cases.with_johnson_cook(R(cases.c3_hexes(3), top_vel=-200.0)), 80),
    "psquad_n6_gmt": (cases.with_gmt(R(cases.plane_strain_quads(6), top_vel=-50.0)), 80),
    # thermal coupling (Thermal.C): conduction, plastic heating, thermal expansion; with contact heat exchange
    "tet_n3_thermal": (cases.with_thermal(R(cases.c2_tets(3), top_vel=-200.0)), 80),
    "contact_quad_n6_thermal": (cases.with_thermal(cases.contact_quads(6), heat_cond=25000.0, T_die=200.0), 80),
}
THERMAL_ARRAYS = "T m_dTedt m_q_plheat q_cont_conv".split()
CONTACT_ARRAYS = ("contforce ut_prev node_area m_elem_area m_mesh_in_contact ext_nodes trimesh.node trimesh.node_v "
                  "trimesh.normal trimesh.pplane").split()

FLOAT_ARRAYS = "x v a u prev_a m_fi m_mdiag vol p pl_strain sigma_y m_sigma m_tau m_eps".split()
INT_ARRAYS = "m_elnod m_nodel m_nodel_loc m_nodel_offset m_nodel_count".split()
