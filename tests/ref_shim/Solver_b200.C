// Solver_b200.C — the reference-side binding of libwf_b200.so: the file a WeldFormFEM maintainer adds next to
// src/explicit/Solver_explicit.C.  It replaces the body of Domain_d::SolveChungHulbert()
// (src/explicit/Solver_explicit.C:101: initialisation :115-292, one step :524-978) by calls through the C ABI of
// include/wf_engine.h and leaves everything else of the reference (JSON deck, LS-Dyna reader, materials, VTK
// output) untouched.  Compiled and run against the unmodified reference tree by tests/test_ref_shim.py
// (oracle/Makefile target `shim`); INTEGRATION.md quotes this file.
//
// Domain_d's state is `protected`, so a subclass reaches every array and option (the pattern of the reference's own
// src/common/test_1el_3D.cpp:67).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "Domain_d.h"
#include "VTKWriter.h"
#include "wf_engine.h"  // this repository's include/

namespace MetFEM {

class Domain_b200 : public Domain_d {
 protected:
  wf_engine *eng = nullptr;
  static void ck(wf_engine *e, int rc) {
    if (rc) { fprintf(stderr, "wf_b200: %s\n", wf_last_error(e)); exit(1); }
  }

 public:
  // hexa viscous hourglass coefficient (f90_ver/src/Mechanical.f90:307 uses 0.06); the C++ reference at this commit has
  // no hexa hourglass, so 0 reproduces it
  double hexa_hg_coeff = 0.0;
  int strict_reference = 0;  // 1 = operation-for-operation flavour for regression runs against the CPU build

  // call once after the deck has been read: mesh, material, BCs and options are already in *this
  void AttachB200(int device) {
    // dom_type (Domain_d.h:101).  main.C leaves m_domtype at its 3D default for a "plStrain" deck (main.C:325-327 only
    // prints), and the reference's 2D kernels never read it: a 2D domain that is not axisymmetric is plane strain.
    const int domtype = (m_dim == 2 && m_domtype != _Axi_Symm_) ? WF_PLANE_STRAIN : (int)m_domtype;
    ck(nullptr, wf_create(&eng, m_dim, m_nodxelem, domtype, device));
    ck(eng, wf_set_mesh(eng, m_node_count, m_elem_count, x, m_elnod));             // rebuilds nodel* like setNodElem
    if (m_domtype == _Axi_Symm_) ck(eng, wf_set_axisymm_vol_weight(eng, m_axisymm_vol_weight ? 1 : 0));
    const Material_ *mt = &materials[0];   // AssignMaterial (Domain_d.C:903); mat[e] is only wired by SolveChungHulbert
    wf_material wm = {};
    wm.model = mt->Material_model == HOLLOMON ? WF_HOLLOMON : WF_BILINEAR;         // Material.cuh:9-13
    wm.E = mt->Elastic().E(); wm.nu = mt->Elastic().Poisson();
    wm.rho0 = rho_0[0];                                                            // setDensity, Domain_d.C:951
    wm.sy0 = mt->sy0; wm.K = mt->K; wm.m = mt->m;                                  // Material.cuh:50-52
    wm.max_edot = m_max_edot;
    ck(eng, wf_set_material(eng, &wm));
    wf_stab st = {m_stab.alpha_free, m_stab.alpha_contact, m_stab.hg_coeff_free, m_stab.hg_coeff_contact,
                  m_stab.av_coeff_div, m_stab.av_coeff_bulk, m_stab.log_factor, m_stab.pspg_scale,
                  m_stab.p_pspg_bulkfac, m_stab.J_min, m_stab.hg_visc, m_stab.hg_stiff, hexa_hg_coeff};
    ck(eng, wf_set_stab(eng, &st));                                                // m_stab, Domain_d.h:140-153
    ck(eng, wf_set_options(eng, m_press_algorithm, m_artifvisc[0], m_artifvisc[1], strict_reference));
    for (int d = 0; d < m_dim; d++) {                                              // AllocateBCs lists, Domain_d.C:1063
      const int *nod = d == 0 ? bcx_nod : d == 1 ? bcy_nod : bcz_nod;
      const double *val = d == 0 ? bcx_val : d == 1 ? bcy_val : bcz_val;
      for (int i = 0; i < bc_count[d]; i++) ck(eng, wf_add_bc_vel(eng, nod[i], d, val[i]));
    }
    ck(eng, wf_allocate_bcs(eng));
  }

  // replaces Domain_d::SolveChungHulbert() (Solver_explicit.C:101) for a fixed time step
  void SolveChungHulbert_b200() {
    ck(eng, wf_init(eng, dt));                                                     // :115-292
    Time = 0.0;
    double tout = 0.0;
    long step_count = 0;
    while (Time < end_t) {                                                         // :346
      if (Time >= tout) {                                                          // the reference's output cadence
        PullState();
        char name[64];
        snprintf(name, sizeof(name), "out_%.4e.vtk", Time);
        VTKWriter writer(this, name);
        writer.writeFile();
        tout += m_dtout;
      }
      // steps until the next output / the end, counted with the reference's own clock arithmetic (Time += dt, :1166)
      int n = 0;
      double t = Time;
      const double stop = tout < end_t ? tout : end_t;
      while (t < stop) { t += dt; n++; }
      ck(eng, wf_step(eng, n));                                                    // rows 1-22 of :524-978, n times
      ck(eng, wf_get_time(eng, &Time, &step_count));
      int bad = 0;
      ck(eng, wf_nonfinite_flag(eng, &bad));                                       // :779-784
      if (bad) printf("Nonfinite internal force\n");
    }
    PullState();
    printf("wf_b200: %ld steps, Time %.6e\n", step_count, Time);
  }

  // member names == array names: no layout code on this side
  void PullState() {
    const size_t nd = (size_t)m_node_count * m_dim, ne = (size_t)m_elem_count;
    struct { const char *nm; double *p; size_t n; } arr[] = {
        {"x", x, nd}, {"v", v, nd}, {"u", u, nd}, {"a", this->a, nd}, {"prev_a", prev_a, nd},
        {"m_sigma", m_sigma, ne * 6}, {"m_tau", m_tau, ne * 6}, {"pl_strain", pl_strain, ne},
        {"p", p, ne}, {"sigma_y", sigma_y, ne}, {"vol", vol, ne}};
    for (auto &q : arr) ck(eng, wf_get_array(eng, q.nm, q.p, q.n * sizeof(double)));
  }

  ~Domain_b200() { if (eng) wf_destroy(eng); }
};

}  // namespace MetFEM
