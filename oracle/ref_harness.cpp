// TEST INFRASTRUCTURE (oracle) — not product code, never linked into the engine.
//
// Drives the UNMODIFIED reference CPU sources (compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libwf_ref.so) through a
// subclass of MetFEM::Domain_d, the pattern the reference itself uses in
// src/common/test_1el_3D.cpp:67.  The step is the member-by-member sequence of
// Domain_d::SolveChungHulbert() (src/explicit/Solver_explicit.C:115-292 for the
// initialisation, :524-978 for one time step, CPU branch, remesh / contact /
// thermal off).  Only two things are added on top of the reference:
//   * zero-filling of the malloc'ed state the reference never initialises
//     (Domain_d.C:457-621 allocates with malloc);
//   * the 3D hexa viscous hourglass force, which does not exist in the C++ at
//     this commit (Mechanical.C:1842-1943 only acts for 2D quads) and is
//     restated here from f90_ver/src/Mechanical.f90:241-344, written into
//     m_f_elem_hg so that the reference's own assemblyForces() (Matrices.C:42)
//     subtracts it exactly as f90_ver/src/Matrices.f90:639-643 does.
// Everything is exported with a plain C ABI so tests can reach it via ctypes.

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <omp.h>
#include <unistd.h>

#include "Domain_d.h"
#include "Mesh.h"

using namespace MetFEM;

namespace {

struct CoutSilencer {
  std::streambuf *old_;
  CoutSilencer() : old_(std::cout.rdbuf(nullptr)) {}
  ~CoutSilencer() { std::cout.rdbuf(old_); }
};

// silence C stdio printf from the reference during setup/steps
struct StdoutSilencer {
  int saved_;
  StdoutSilencer() {
    fflush(stdout);
    saved_ = dup(1);
    FILE *nul = fopen("/dev/null", "w");
    dup2(fileno(nul), 1);
    fclose(nul);
  }
  ~StdoutSilencer() {
    fflush(stdout);
    dup2(saved_, 1);
    close(saved_);
  }
};

class Harness : public Domain_d {
 public:
  int press_variant = 0;     // 0: solver default (alg 0), 1: ANP as shipped, 3: ANP_Nodal
  double hexa_hg_coeff = 0;  // 0 = off (reference C++ behaviour); 0.06 = F90 value
  Material_ *mat_h = nullptr;
  double time_ = 0.0;
  long step_count_ = 0;
  double3 *v_orig_ = nullptr;  // m_v_orig, Solver_explicit.C:168-173

  Harness() {
    m_faceCount = 0;
    contact = false;
    m_thermal = false;
    m_artifvisc[0] = m_artifvisc[1] = 0.0;
    m_nodxelem = 0;
    m_gp_count = 1;
    trimesh = nullptr;
    m_timeint_type = TimeInt::EXPLICIT;
  }

  void zero_state() {
    const size_t nd = (size_t)m_node_count * m_dim;
    const size_t ne = (size_t)m_elem_count;
    const size_t nk = ne * m_nodxelem * m_dim;
    auto z = [](double *q, size_t n) { if (q) memset(q, 0, n * sizeof(double)); };
    z(prev_a, nd); z(m_fe, nd); z(m_fi, nd); z(a, nd); z(v, nd); z(u, nd); z(u_dt, nd);
    z(contforce, nd); z(ut_prev, nd);
    z(m_tau, 6 * ne); z(m_sigma, 6 * ne); z(m_eps, 6 * ne); z(m_str_rate, 6 * ne);
    z(m_rot_rate, 6 * ne); z(m_strain_pl_incr, 6 * ne);
    z(p, ne); z(pl_strain, ne); z(sigma_y, ne); z(m_radius, ne); z(rho, ne); z(rho_0, ne);
    z(vol, ne); z(vol_0, ne); z(m_detJ, ne);
    z(m_f_elem, nk); z(m_f_elem_hg, nk);
    z(m_mdiag, m_node_count); z(m_voln, m_node_count); z(p_node, m_node_count);
    z(m_dTedt, ne * m_nodxelem); z(m_q_plheat, ne);
    z(T, m_node_count); z(node_area, m_node_count); z(q_cont_conv, m_node_count); z(m_elem_area, ne);
    z(m_elem_length, ne);
    if (ext_nodes) memset(ext_nodes, 0, sizeof(bool) * m_node_count);
    if (m_mesh_in_contact) memset(m_mesh_in_contact, 0xff, sizeof(int) * m_node_count);
    if (m_dim == 2 && m_hg_q) z(m_hg_q, nk);
    m_faceCount = 0;
    contact = false;
    m_thermal = false;
  }

  void box(const double *V, const double *L, double r, int tritet) {
    CoutSilencer s; StdoutSilencer s2;
    AddBoxLength(make_double3(V[0], V[1], V[2]), make_double3(L[0], L[1], L[2]), r, true, tritet != 0);
    zero_state();
  }

  // same steps as Domain_d::CreateFromLSDyna (Domain_d.C:1647-1699) for any dim
  void mesh(int dim, int k, int nn, int ne, const double *xin, const int *elnod) {
    CoutSilencer s; StdoutSilencer s2;
    m_dim = dim;
    m_gp_count = 1;
    m_nodxelem = k;
    SetDimension(nn, ne);
    memcpy(x, xin, sizeof(double) * (size_t)nn * dim);
    std::vector<int> el(elnod, elnod + (size_t)ne * k);
    setNodElem(el.data());
    zero_state();
  }

  // ---- hooks for the reference's own front-end --------------------------------------------------------
  // src/explicit/main.C is compiled UNMODIFIED at the end of this file with `Domain_d` spelled `Harness`, so that its
  // `new Domain_d` creates this subclass and the three calls below resolve (by name hiding) to these wrappers: the two
  // mesh builders add the zero-fill of malloc'ed state (as box() / mesh() above do), and the final
  // `dom_d->SolveChungHulbert()` hands the fully set-up domain to the test instead of running the VTK-writing loop.
  static Harness *&deck_captured() { static Harness *p = nullptr; return p; }
  void AddBoxLength(double3 const &V, double3 const &L, const double &r, const bool &red_int = true, const bool &tritetra = false) {
    MetFEM::Domain_d::AddBoxLength(V, L, r, red_int, tritetra);
    zero_state();
  }
  void CreateFromLSDyna(LS_Dyna::lsdynaReader &reader) {
    MetFEM::Domain_d::CreateFromLSDyna(reader);
    zero_state();
  }
  void SolveChungHulbert() {
    press_variant = m_press_algorithm == 1 ? 1 : 0;
    deck_captured() = this;
    // leave main.C by unwinding: renamed, its `int main` falls off the end without a return (fine for main, undefined
    // for any other function)
    throw DeckDone{};
  }
  struct DeckDone {};
  double deck_dt() const { return dt; }
  double deck_end_t() const { return end_t; }

  // src/explicit/main.C:460-581
  void material(double E, double nu, double rho0, int model, double sy0v, double K, double mexp) {
    CoutSilencer s; StdoutSilencer s2;
    setDensity(rho0);
    Elastic_ el(E, nu);
    if (model == HOLLOMON) {
      mat_h = new Hollomon(el, sy0v, K, mexp);
      mat_h->InitHollomon(el, sy0v, K, mexp);
      mat_h->Material_model = HOLLOMON;
    } else {
      mat_h = new Material_(el);
      mat_h->Ep = 0.0;
      mat_h->Material_model = BILINEAR;
    }
    mat_h->cs0 = sqrt(mat_h->Elastic().BulkMod() / rho0);
    mat_h->sy0 = sy0v;
    AssignMaterial(mat_h);
  }

  // Johnson-Cook / GMT.  main.C:535-558 builds `JohnsonCook` / `GMT` objects whose constructor arguments
  // land in PRIVATE members that shadow the public Material_ fields of the same names (Material.cuh:176-197, 237-274),
  // while the free functions CalcStressStrain calls (Material.cuh:377-483) read the PUBLIC Material_ fields — which
  // that path leaves uninitialised.  The defined behaviour of those functions is therefore pinned by filling the public
  // fields: Material_::Init_JohnsonCook (Material.cuh:105-116) for JC, plain assignment for GMT.  The temperature is
  // the nodal array T indexed by ELEMENT id (Mechanical.C:1731); with thermal coupling off it is uniform.
  void material_ext(double E, double nu, double rho0, int model, double sy0v, const double *q, double temp) {
    CoutSilencer s; StdoutSilencer s2;
    setDensity(rho0);
    Elastic_ el(E, nu);
    mat_h = new Material_(el);
    if (model == JOHNSON_COOK) {
      mat_h->Init_JohnsonCook(el, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);  // a, b, n, c, eps_0, m, T_m, T_t
    } else {
      mat_h->Material_model = _GMT_;
      mat_h->n1 = q[0]; mat_h->n2 = q[1]; mat_h->C1 = q[2]; mat_h->C2 = q[3]; mat_h->m1 = q[4]; mat_h->m2 = q[5];
      mat_h->I1 = q[6]; mat_h->I2 = q[7];
      mat_h->e_min = q[8]; mat_h->e_max = q[9]; mat_h->er_min = q[10]; mat_h->er_max = q[11];
      mat_h->T_min = q[12]; mat_h->T_max = q[13];
    }
    mat_h->cs0 = sqrt(mat_h->Elastic().BulkMod() / rho0);
    mat_h->sy0 = sy0v;
    AssignMaterial(mat_h);
    for (int n = 0; n < m_node_count; n++) T[n] = temp;
  }

  // thermal coupling: main.C:218, 436-441, 567-570 (plHeatFrac, setThermalOn + setTemp, k_T / cp_T / exp_T)
  void thermal_on(double k_T, double cp_T, double exp_T, double plheatfrac, double T0) {
    setThermalOn();
    setTemp(T0);
    // main.C:567-570 fills these BEFORE AssignMaterial copies the object (Domain_d.C:903-907); here the copy exists already
    mat_h->k_T = k_T; mat_h->cp_T = cp_T; mat_h->exp_T = exp_T;
    materials[0].k_T = k_T; materials[0].cp_T = cp_T; materials[0].exp_T = exp_T;
    m_plheatfraction = plheatfrac;
  }

  // src/explicit/Solver_explicit.C:115-292, CPU branch
  void init(double dt_) {
    CoutSilencer s; StdoutSilencer s2;
    SetDT(dt_);
    AssignMatAddress();
    InitValues();
    for (int d = 0; d < m_dim; d++) {
      for (int n = 0; n < m_node_count * m_dim; n++) v[n] = a[n] = u[n] = 0.0;
      ImposeBCV(d);
    }
    double rho_b = 0.818200;
    m_alpha = (2.0 * rho_b - 1.0) / (1.0 + rho_b);
    m_beta = (5.0 - 3.0 * rho_b) / ((1.0 + rho_b) * (1.0 + rho_b) * (2.0 - rho_b));
    m_gamma = 1.5 - m_alpha;
    calcElemJAndDerivatives();
    if (m_dim == 2 && m_domtype == _Axi_Symm_) Calc_Element_Radius();
    CalcElemInitialVol();
    CalcElemVol();
    calcElemDensity();
    CalcNodalVol();
    CalcNodalMassFromVol();
    for (int n = 0; n < m_node_count * m_dim; n++) ut_prev[n] = 0.0;
    time_ = 0.0;
    step_count_ = 0;
    if (contact) {  // Solver_explicit.C:168-173
      v_orig_ = new double3[trimesh->nodecount];
      for (int n = 0; n < trimesh->nodecount; n++) v_orig_[n] = trimesh->node_v[n];
    }
  }

  // ---- contact with rigid surfaces ----------------------------------------------------
  // main.C:672-708 (first body) / :775-828 (further bodies): AxisPlaneMesh, node velocities, AddMesh
  void add_plane(int dimension, int id, int axis, int positaxisorent, const double *p1, const double *p2, int dens,
                 const double *vel) {
    CoutSilencer s; StdoutSilencer s2;
    double3 a = make_double3(p1[0], p1[1], p1[2]), b = make_double3(p2[0], p2[1], p2[2]);
    double3 vv = make_double3(vel[0], vel[1], vel[2]);
    if (!trimesh) {
      TriMesh_d *m = new TriMesh_d();
      m->dimension = dimension;
      m->AxisPlaneMesh(id, axis, positaxisorent != 0, a, b, dens);
      setTriMesh(m);
      for (int nc = 0; nc < m->nodecount; nc++) m->node_v[nc] = vv;
      m->mu_sta[0] = m->mu_dyn[0] = 0.0;
    } else {
      TriMesh_d *m = new TriMesh_d();
      m->dimension = dimension;
      m->AxisPlaneMesh(id, axis, positaxisorent != 0, a, b, dens);
      m->SetMeshVel(vv);
      addMeshData(*m);
      delete m;
    }
  }
  // the same TriMesh_d state from raw arrays (what a caller of the engine's wf_set_trimesh holds)
  void set_trimesh(int dimension, int nn, int ne, const double *node, const double *node_v, const int *elnode,
                   const double *normal, const int *mesh_id) {
    TriMesh_d *m = new TriMesh_d();
    const int nen = dimension == 3 ? 3 : 2;
    m->dimension = dimension; m->nodecount = nn; m->elemcount = ne; m->mesh_count = 1;
    m->node = (double3 *)malloc(sizeof(double3) * nn);
    m->node_v = (double3 *)malloc(sizeof(double3) * nn);
    m->elnode = (int *)malloc(sizeof(int) * nen * ne);
    m->centroid = (double3 *)malloc(sizeof(double3) * ne);
    m->normal = (double3 *)malloc(sizeof(double3) * ne);
    m->pplane = (double *)malloc(sizeof(double) * ne);
    m->nfar = (int *)malloc(sizeof(int) * ne);
    m->ele_mesh_id = (int *)malloc(sizeof(int) * ne);
    m->mu_sta = (double *)calloc(1, sizeof(double));
    m->mu_dyn = (double *)calloc(1, sizeof(double));
    m->react_force = (double3 *)calloc(1, sizeof(double3));
    m->react_p_force = (double *)calloc(1, sizeof(double));
    for (int i = 0; i < nn; i++) {
      m->node[i] = make_double3(node[3 * i], node[3 * i + 1], node[3 * i + 2]);
      m->node_v[i] = make_double3(node_v[3 * i], node_v[3 * i + 1], node_v[3 * i + 2]);
    }
    memcpy(m->elnode, elnode, sizeof(int) * nen * ne);
    for (int e = 0; e < ne; e++) {
      m->normal[e] = make_double3(normal[3 * e], normal[3 * e + 1], normal[3 * e + 2]);
      m->ele_mesh_id[e] = mesh_id[e];
    }
    m->CalcCentroids();
    setTriMesh(m);
  }
  // main.C:716-725 (friction, penalty factor), :842-847 (CalcSpheres, setContactOn), SetEndTime
  void contact_on(double mu_sta, double mu_dyn, double pf, double end_time) {
    CoutSilencer s; StdoutSilencer s2;
    trimesh->mu_sta[0] = mu_sta; trimesh->mu_dyn[0] = mu_dyn;
    if (pf > -1.0) setContactPF(pf);
    trimesh->CalcSpheres();
    setContactOn();
    SetEndTime(end_time);
  }
  void search_ext_nodes() { CoutSilencer s; StdoutSilencer s2; SearchExtNodes(); }  // main.C:650
  // Solver_explicit.C:981-1005: velocity ramp of the rigid surfaces, Move, centroids, normals, plane coefficients
  void move_trimesh() {
    const double RAMP_FRACTION = 1.0e-2;  // Solver_explicit.C:309
    double f = 1.0;
    if (time_ < RAMP_FRACTION * end_t) f = pow(time_ / (RAMP_FRACTION * end_t), 0.5);
    for (int n = 0; n < trimesh->nodecount; n++) trimesh->node_v[n] = f * v_orig_[n];
    trimesh->Move(dt);
    trimesh->CalcCentroids();
    trimesh->CalcNormals();
    trimesh->UpdatePlaneCoeff();
  }

  // 3D hexa viscous hourglass, f90_ver/src/Mechanical.f90:241-344 (Goudreau 1982):
  //   Sig = 4x8 table of +-1 (0.125 * 8), hmod(d,j) = sum_n v(n,d)*Sig(j,n),
  //   f(n,d) = (0 - sum_j hmod(d,j)*Sig(j,n)) * c_h,
  //   c_h = coeff * vol**0.6666666 * rho * 0.25 * cs0
  void hexa_hourglass() {
    static const double Sig[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1},
                                     {1, -1, -1, 1, -1, 1, 1, -1},
                                     {1, -1, 1, -1, 1, -1, 1, -1},
                                     {-1, 1, -1, 1, 1, -1, 1, -1}};
    const double cs0 = mat[0]->cs0;
#pragma omp parallel for
    for (int e = 0; e < m_elem_count; e++) {
      double vel[8][3], hmod[3][4], f[8][3];
      for (int n = 0; n < 8; n++)
        for (int d = 0; d < 3; d++) vel[n][d] = v[3 * m_elnod[8 * e + n] + d];
      for (int d = 0; d < 3; d++)
        for (int j = 0; j < 4; j++) hmod[d][j] = 0.0;
      for (int j = 0; j < 4; j++)
        for (int n = 0; n < 8; n++)
          for (int d = 0; d < 3; d++) hmod[d][j] = hmod[d][j] + vel[n][d] * Sig[j][n];
      for (int n = 0; n < 8; n++) {
        for (int d = 0; d < 3; d++) f[n][d] = 0.0;
        for (int j = 0; j < 4; j++)
          for (int d = 0; d < 3; d++) f[n][d] = f[n][d] - hmod[d][j] * Sig[j][n];
      }
      double c_h = hexa_hg_coeff * pow(vol[e], 0.6666666) * rho[e] * 0.25 * cs0;
      for (int n = 0; n < 8; n++)
        for (int d = 0; d < 3; d++) m_f_elem_hg[e * 24 + n * 3 + d] = f[n][d] * c_h;
    }
  }

  void pressure() {
    if (press_variant == 0) {
      if (m_dim == 3) calcElemPressure(); else calcElemPressureLocal();
    } else if (press_variant == 1) {
      calcElemPressureANP();
    } else if (press_variant == 3) {
      calcElemPressureANP_Nodal();
    }
  }

  void hourglass() {
    calcElemHourglassForces();
    if (m_dim == 3 && m_nodxelem == 8 && hexa_hg_coeff != 0.0) hexa_hourglass();
  }

  void axis_constraint() {  // Solver_explicit.C:953-969
    if (m_domtype == _Axi_Symm_) {
      double xmin = 1000.0;
      for (int i = 0; i < getNodeCount(); i++)
        if (getPosVec2(i).x < xmin) xmin = getPosVec2(i).x;
      for (int i = 0; i < getNodeCount(); i++)
        if (getPosVec2(i).x <= xmin + 1.e-6) { a[m_dim * i] = 0.0; v[m_dim * i] = 0.0; }
    }
  }

  // one time step: Solver_explicit.C:524-978 (rows 1-22 of SURVEY.md §3.3)
  void step_once() {
    if (m_dim > 2 && m_faceCount > 0 && step_count_ % 10 == 0) CalcExtFaceAreas();  // Solver_explicit.C:445-450
    UpdatePrediction();
    for (int d = 0; d < m_dim; d++) ImposeBCV(d);
    calcElemJAndDerivatives();
    if (m_dim == 2 && m_domtype == _Axi_Symm_) Calc_Element_Radius();
    CalcElemVol();
    CalcNodalVol();
    CalcNodalMassFromVol();
    calcElemStrainRates();
    if (m_thermal) calcThermalExpansion();  // Solver_explicit.C:719-720
    pressure();
    calcNodalPressureFromElemental();
    CalcStressStrain(dt);
    calcArtificialViscosity();
    calcElemForces();
    hourglass();
    if (contact) CalcContactForces();  // Solver_explicit.C:769-770
    assemblyForces();
    for (int i = 0; i < m_node_count * m_dim; ++i)
      if (!std::isfinite(m_fi[i])) m_fi[i] = 0.0;
    calcAccel();
    ImposeBCAAllDim();
    UpdateCorrectionAccVel();
    ImposeBCVAllDim();
    axis_constraint();
    UpdateCorrectionPos();
    if (contact) move_trimesh();
    if (m_thermal) ThermalCalcs();  // Solver_explicit.C:1008-1012
    time_ += dt;
    step_count_++;
  }

  void steps(int n) {
    StdoutSilencer s2; CoutSilencer s;
    for (int i = 0; i < n; i++) step_once();
  }

  int call(const std::string &f, double arg) {
    StdoutSilencer s2; CoutSilencer s;
    if (f == "UpdatePrediction") UpdatePrediction();
    else if (f == "ImposeBCV") ImposeBCV((int)arg);
    else if (f == "ImposeBCVAllDim") ImposeBCVAllDim();
    else if (f == "ImposeBCA") ImposeBCA((int)arg);
    else if (f == "ImposeBCAAllDim") ImposeBCAAllDim();
    else if (f == "calcElemJAndDerivatives") calcElemJAndDerivatives();
    else if (f == "Calc_Element_Radius") Calc_Element_Radius();
    else if (f == "CalcElemVol") CalcElemVol();
    else if (f == "CalcElemInitialVol") CalcElemInitialVol();
    else if (f == "calcElemDensity") calcElemDensity();
    else if (f == "CalcNodalVol") CalcNodalVol();
    else if (f == "CalcNodalMassFromVol") CalcNodalMassFromVol();
    else if (f == "calcElemStrainRates") calcElemStrainRates();
    else if (f == "calcElemPressure") pressure();
    else if (f == "calcNodalPressureFromElemental") calcNodalPressureFromElemental();
    else if (f == "SetDT") SetDT(arg);
    else if (f == "calcMinEdgeLength") { if (m_dim == 2 && m_nodxelem != 4) return -1; calcMinEdgeLength(); }
    else if (f == "CalcStressStrain") CalcStressStrain(arg);
    else if (f == "calcArtificialViscosity") calcArtificialViscosity();
    else if (f == "calcElemForces") calcElemForces();
    else if (f == "calcElemHourglassForces") hourglass();
    else if (f == "assemblyForces") assemblyForces();
    else if (f == "calcAccel") calcAccel();
    else if (f == "UpdateCorrectionAccVel") UpdateCorrectionAccVel();
    else if (f == "AxisConstraint") axis_constraint();
    else if (f == "UpdateCorrectionPos") UpdateCorrectionPos();
    else if (f == "calcThermalExpansion") calcThermalExpansion();
    else if (f == "ThermalCalcs") ThermalCalcs();
    else if (f == "SearchExtNodes") SearchExtNodes();
    else if (f == "CalcExtFaceAreas") CalcExtFaceAreas();
    else if (f == "CalcContactForces") CalcContactForces();
    else if (f == "MoveTriMesh") move_trimesh();
    else return -1;
    return 0;
  }

  struct View { const void *ptr; size_t bytes; };
  View view(const std::string &nm) {
    const size_t nd = sizeof(double) * (size_t)m_node_count * m_dim;
    const size_t nn = sizeof(double) * (size_t)m_node_count;
    const size_t ne = sizeof(double) * (size_t)m_elem_count;
    const size_t nk = ne * m_nodxelem;
    size_t ntot = 0;
    if (m_nodel_offset && m_node_count > 0)
      ntot = (size_t)m_nodel_offset[m_node_count - 1] + m_nodel_count[m_node_count - 1];
    if (nm == "x") return {x, nd};
    if (nm == "v") return {v, nd};
    if (nm == "a") return {a, nd};
    if (nm == "u") return {u, nd};
    if (nm == "u_dt") return {u_dt, nd};
    if (nm == "prev_a") return {prev_a, nd};
    if (nm == "m_fi") return {m_fi, nd};
    if (nm == "m_fe") return {m_fe, nd};
    if (nm == "m_mdiag") return {m_mdiag, nn};
    if (nm == "m_voln") return {m_voln, nn};
    if (nm == "p_node") return {p_node, nn};
    if (nm == "m_dH_detJ_dx") return {m_dH_detJ_dx, nk};
    if (nm == "m_dH_detJ_dy") return {m_dH_detJ_dy, nk};
    if (nm == "m_dH_detJ_dz") return {m_dH_detJ_dz, nk};
    if (nm == "m_detJ") return {m_detJ, ne};
    if (nm == "vol") return {vol, ne};
    if (nm == "vol_0") return {vol_0, ne};
    if (nm == "rho") return {rho, ne};
    if (nm == "rho_0") return {rho_0, ne};
    if (nm == "p") return {p, ne};
    if (nm == "pl_strain") return {pl_strain, ne};
    if (nm == "sigma_y") return {sigma_y, ne};
    if (nm == "m_radius") return {m_radius, ne};
    if (nm == "m_str_rate") return {m_str_rate, 6 * ne};
    if (nm == "m_rot_rate") return {m_rot_rate, 6 * ne};
    if (nm == "m_sigma") return {m_sigma, 6 * ne};
    if (nm == "m_tau") return {m_tau, 6 * ne};
    if (nm == "m_eps") return {m_eps, 6 * ne};
    if (nm == "m_f_elem") return {m_f_elem, nk * m_dim};
    if (nm == "m_f_elem_hg") return {m_f_elem_hg, nk * m_dim};
    if (nm == "m_hg_q") return {m_dim == 2 ? m_hg_q : nullptr, m_dim == 2 ? nk * m_dim : 0};
    if (nm == "m_elem_length") return {m_elem_length, ne};
    if (nm == "T") return {T, nn};
    if (nm == "m_dTedt") return {m_dTedt, nk};
    if (nm == "m_q_plheat") return {m_q_plheat, ne};
    if (nm == "q_cont_conv") return {q_cont_conv, nn};
    if (nm == "contforce") return {contforce, nd};
    if (nm == "ut_prev") return {ut_prev, nd};
    if (nm == "node_area") return {node_area, nn};
    if (nm == "m_elem_area") return {m_elem_area, ne};
    if (nm == "ext_nodes") return {ext_nodes, sizeof(bool) * (size_t)m_node_count};
    if (nm == "m_mesh_in_contact") return {m_mesh_in_contact, sizeof(int) * (size_t)m_node_count};
    if (trimesh) {
      const size_t tn = sizeof(double3) * (size_t)trimesh->nodecount, te = (size_t)trimesh->elemcount;
      if (nm == "trimesh.node") return {trimesh->node, tn};
      if (nm == "trimesh.node_v") return {trimesh->node_v, tn};
      if (nm == "trimesh.normal") return {trimesh->normal, sizeof(double3) * te};
      if (nm == "trimesh.pplane") return {trimesh->pplane, sizeof(double) * te};
      if (nm == "trimesh.elnode") return {trimesh->elnode, sizeof(int) * te * (trimesh->dimension == 3 ? 3 : 2)};
      if (nm == "trimesh.ele_mesh_id") return {trimesh->ele_mesh_id, sizeof(int) * te};
    }
    if (nm == "bcx_val") return {bcx_val, sizeof(double) * (size_t)bc_count[0]};
    if (nm == "bcy_val") return {bcy_val, sizeof(double) * (size_t)bc_count[1]};
    if (nm == "bcz_val") return {bcz_val, sizeof(double) * (size_t)bc_count[2]};
    if (nm == "m_elnod") return {m_elnod, sizeof(unsigned) * (size_t)m_elem_count * m_nodxelem};
    if (nm == "m_nodel") return {m_nodel, sizeof(int) * ntot};
    if (nm == "m_nodel_loc") return {m_nodel_loc, sizeof(int) * ntot};
    if (nm == "m_nodel_offset") return {m_nodel_offset, sizeof(int) * (size_t)m_node_count};
    if (nm == "m_nodel_count") return {m_nodel_count, sizeof(int) * (size_t)m_node_count};
    return {nullptr, 0};
  }

  void info(int *out) {
    out[0] = m_dim; out[1] = m_nodxelem; out[2] = m_node_count; out[3] = m_elem_count;
    out[4] = bc_count[0]; out[5] = bc_count[1]; out[6] = bc_count[2];
    out[7] = (int)m_domtype;
  }
  void consts(double *out) {
    out[0] = m_alpha; out[1] = m_beta; out[2] = m_gamma; out[3] = dt; out[4] = time_;
    out[5] = m_min_length; out[6] = m_min_height;
  }
  void energies(double *ek, double *dei) {
    double ev = 0.0;
    computeEnergies(dt, *ek, *dei, ev);
  }
};

}  // namespace

extern "C" {

void *wfref_new() { CoutSilencer s; return new Harness(); }
void wfref_free(void *h) { delete (Harness *)h; }  // arrays are leaked on purpose (reference Free() is partial)
void wfref_set_threads(int n) { omp_set_num_threads(n); }
int wfref_max_threads() { return omp_get_max_threads(); }

// domtype: 0 plane strain, 2 axisymmetric, 3 3D (Domain_d.h:101); call BEFORE meshing
void wfref_set_domtype(void *h, int domtype, int vol_weight) {
  Harness *d = (Harness *)h;
  if (domtype == 2) d->setAxiSymm(vol_weight != 0);
  else d->m_domtype = (dom_type)domtype;
}
void wfref_box(void *h, const double *V, const double *L, double r, int tritet) { ((Harness *)h)->box(V, L, r, tritet); }
void wfref_set_mesh(void *h, int dim, int k, int nn, int ne, const double *x, const int *elnod) {
  ((Harness *)h)->mesh(dim, k, nn, ne, x, elnod);
}
void wfref_set_material(void *h, double E, double nu, double rho0, int model, double sy0, double K, double m) {
  ((Harness *)h)->material(E, nu, rho0, model, sy0, K, m);
}
void wfref_set_material_ext(void *h, double E, double nu, double rho0, int model, double sy0, const double *q, double temp) {
  ((Harness *)h)->material_ext(E, nu, rho0, model, sy0, q, temp);
}
void wfref_thermal_on(void *h, double k_T, double cp_T, double exp_T, double plheatfrac, double T0) {
  ((Harness *)h)->thermal_on(k_T, cp_T, exp_T, plheatfrac, T0);
}
void wfref_set_contact_heat(void *h, double heat_cond, double T_const) {
  TriMesh_d *m = ((Harness *)h)->getTriMesh();
  if (m) { m->heat_cond = heat_cond; m->T_const = T_const; }
}
void wfref_set_max_edot(void *h, double v) { ((Harness *)h)->m_max_edot = v; }
// order = StabilizationParams fields (Domain_d.h:140-153)
void wfref_set_stab(void *h, const double *s) {
  StabilizationParams &p = ((Harness *)h)->m_stab;
  p.alpha_free = s[0]; p.alpha_contact = s[1]; p.hg_coeff_free = s[2]; p.hg_coeff_contact = s[3];
  p.av_coeff_div = s[4]; p.av_coeff_bulk = s[5]; p.log_factor = s[6]; p.pspg_scale = s[7];
  p.p_pspg_bulkfac = s[8]; p.J_min = s[9]; p.hg_visc = s[10]; p.hg_stiff = s[11];
}
void wfref_set_options(void *h, int press_variant, double av_alpha, double av_beta, double hexa_hg_coeff) {
  Harness *d = (Harness *)h;
  d->press_variant = press_variant;
  d->m_press_algorithm = press_variant == 1 ? 1 : 0;
  d->m_artifvisc[0] = av_alpha;
  d->m_artifvisc[1] = av_beta;
  d->hexa_hg_coeff = hexa_hg_coeff;
}
void wfref_add_plane(void *h, int dimension, int id, int axis, int positaxisorent, const double *p1, const double *p2,
                     int dens, const double *vel) {
  ((Harness *)h)->add_plane(dimension, id, axis, positaxisorent, p1, p2, dens, vel);
}
void wfref_set_trimesh(void *h, int dimension, int nn, int ne, const double *node, const double *node_v,
                       const int *elnode, const double *normal, const int *mesh_id) {
  ((Harness *)h)->set_trimesh(dimension, nn, ne, node, node_v, elnode, normal, mesh_id);
}
void wfref_contact_on(void *h, double mu_sta, double mu_dyn, double pf, double end_time) {
  ((Harness *)h)->contact_on(mu_sta, mu_dyn, pf, end_time);
}
void wfref_trimesh_counts(void *h, int *out) {
  Harness *d = (Harness *)h;
  TriMesh_d *m = d->getTriMesh();
  out[0] = m ? m->dimension : 0;
  out[1] = m ? m->nodecount : 0;
  out[2] = m ? m->elemcount : 0;
}
void wfref_add_bc(void *h, int node, int dim, double val) { ((Harness *)h)->AddBCVelNode(node, dim, val); }
void wfref_allocate_bcs(void *h) { CoutSilencer s; ((Harness *)h)->AllocateBCs(); }
void wfref_init(void *h, double dt) { ((Harness *)h)->init(dt); }
void wfref_step(void *h, int n) { ((Harness *)h)->steps(n); }
double wfref_time_steps(void *h, int n) {
  Harness *d = (Harness *)h;
  StdoutSilencer s2; CoutSilencer s;
  double t0 = omp_get_wtime();
  for (int i = 0; i < n; i++) d->step_once();
  return omp_get_wtime() - t0;
}
// the reference's own whole-solver entry point (prints + VTK; for the bit-equality probe only)
void wfref_solve_chung_hulbert(void *h, double dt, double end_t) {
  Harness *d = (Harness *)h;
  StdoutSilencer s2; CoutSilencer s;
  d->SetDT(dt); d->SetEndTime(end_t); d->setdtOut(1.0e10); d->setFixedDt(true);
  d->MetFEM::Domain_d::SolveChungHulbert();
}
int wfref_call(void *h, const char *fn, double arg) { return ((Harness *)h)->call(fn, arg); }
long wfref_get(void *h, const char *name, void *dst, long cap) {
  Harness::View w = ((Harness *)h)->view(name);
  if (!w.ptr) return -1;
  if ((long)w.bytes > cap) return -(long)w.bytes;
  memcpy(dst, w.ptr, w.bytes);
  return (long)w.bytes;
}
long wfref_set(void *h, const char *name, const void *src, long bytes) {
  Harness::View w = ((Harness *)h)->view(name);
  if (!w.ptr || (long)w.bytes != bytes) return -1;
  memcpy(const_cast<void *>(w.ptr), src, w.bytes);
  return bytes;
}
void wfref_info(void *h, int *out) { ((Harness *)h)->info(out); }
void wfref_consts(void *h, double *out) { ((Harness *)h)->consts(out); }
void wfref_energies(void *h, double *ek, double *dei) { ((Harness *)h)->energies(ek, dei); }

}  // extern "C"

// ---- the reference's deck front-end: src/explicit/main.C, unmodified, with its Domain_d spelled Harness ------------
// (`#include "Domain_d.h"` inside main.C is not macro-expanded and is include-guarded.)  main() becomes wf_ref_main().
// NastranReader::read is declared `inline` in src/common/NastranReader.cpp:34, so its definition must be visible in the
// translation unit that calls it (main.C:690, rigid bodies of type "File").
#include "src/common/NastranReader.cpp"
#define Domain_d Harness
#define main wf_ref_main
#include "src/explicit/main.C"
#undef main
#undef Domain_d

extern "C" {
// Run main.C on `deck` (path relative to the current directory, as the reference resolves `fileName` relative to it).
// Returns the set-up domain (not initialised, not stepped) or NULL.  out[0] = dt chosen by main.C, out[1] = end time.
void *wfref_load_deck(const char *deck, double *out) {
  Harness::deck_captured() = nullptr;
  const int threads = omp_get_max_threads();
  char a0[] = "WeldFormFEM";
  std::string d = deck;
  char *argv[] = {a0, d.data(), nullptr};
  std::streambuf *olderr = std::cerr.rdbuf();
  try {
    if (getenv("WF_REF_VERBOSE")) {
      wf_ref_main(2, argv);
    } else {
      StdoutSilencer s2; CoutSilencer s;
      std::cerr.rdbuf(nullptr);
      wf_ref_main(2, argv);
    }
    Harness::deck_captured() = nullptr;  // main.C returned without reaching the solver call
  } catch (const Harness::DeckDone &) {
  } catch (const std::exception &e) {  // nlohmann::json parse / type errors
    std::cerr.rdbuf(olderr);
    fprintf(stderr, "wfref_load_deck: %s\n", e.what());
    Harness::deck_captured() = nullptr;
  }
  std::cerr.rdbuf(olderr);
  omp_set_num_threads(threads);  // main.C:338 sets it from "Nproc"
  Harness *h = Harness::deck_captured();
  if (h && out) { out[0] = h->deck_dt(); out[1] = h->deck_end_t(); }
  return h;
}
}
