// wf_domain.hpp — C++ host side of the B200 explicit engine: a `Domain_d` with the reference's method names
// (include/common/Domain_d.h:231-700 of luchete80/WeldFormFEM) whose every call forwards to the C ABI of
// include/wf_engine.h.  Header-only; link with -lwf_b200.  The drivers in this directory
// (main_1_elem_3d.cpp, wf_explicit.cpp) read like the reference's own src/common/main_1_elem_3d.C.
//
// There is no CPU path behind this class: without libwf_b200.so / a CUDA device every call fails and
// the error text of wf_last_error() is thrown as std::runtime_error (the reference prints and goes on;
// a caller that wants that behaviour catches and prints).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "wf_engine.h"

namespace wf_b200 {

struct double3 {
  double x, y, z;
};
inline double3 make_double3(double x, double y, double z) { return double3{x, y, z}; }

// Elastic_ (include/common/Material.cuh:15-40)
class Elastic_ {
 public:
  Elastic_() = default;
  Elastic_(double e, double nu) : E_(e), nu_(nu) {}
  double E() const { return E_; }
  double Poisson() const { return nu_; }
  double BulkMod() const { return E_ / (3.0 * (1.0 - 2.0 * nu_)); }
  double G() const { return E_ / (2.0 * (1.0 + nu_)); }

 private:
  double E_ = 0.0, nu_ = 0.0;
};

enum Material_model_t { BILINEAR = WF_BILINEAR, HOLLOMON = WF_HOLLOMON, JOHNSON_COOK = 2, GMT = 3 };

// Material_ (include/common/Material.cuh:42-156): only the members the explicit step reads
class Material_ {
 public:
  Material_() = default;
  explicit Material_(const Elastic_ &el) : elastic_(el) {}
  // Material_::InitHollomon (Material.cuh:90-104): eps0 / eps1 are derived inside the engine the same way
  void InitHollomon(const Elastic_ &el, double sy0_, double K_, double m_) {
    elastic_ = el;
    sy0 = sy0_;
    K = K_;
    m = m_;
    Material_model = HOLLOMON;
  }
  const Elastic_ &Elastic() const { return elastic_; }
  // Material_::Init_JohnsonCook (Material.cuh:105-116)
  void Init_JohnsonCook(const Elastic_ &el, double a, double b, double n_, double c, double eps_0_, double m_, double T_m_, double T_t_) {
    elastic_ = el;
    Material_model = JOHNSON_COOK;
    A = a; B = b; C = c; m = m_; n = n_; eps_0 = eps_0_; T_m = T_m_; T_t = T_t_;
  }
  int Material_model = BILINEAR;
  double sy0 = 1.0e10, K = 0.0, m = 1.0, cs0 = 0.0, Ep = 0.0;
  double A = 0, B = 0, C = 0, n = 0, eps_0 = 1.0, T_m = 1.0e10, T_t = 0.0;                   // Johnson-Cook (Material.cuh:67-72)
  double C1 = 0, C2 = 0, m1 = 0, m2 = 0, n1 = 0, n2 = 0, I1 = 0, I2 = 0;                    // GMT (Material.cuh:76)
  double e_min = 0, e_max = 1.0e10, er_min = 0, er_max = 1.0e10, T_min = 0, T_max = 1.0e10;  // Material.cuh:63-65
  double k_T = 0.0, cp_T = 0.0, exp_T = 0.0;                                                 // thermal (Material.cuh:79-81)

 private:
  Elastic_ elastic_;
};

// StabilizationParams (include/common/Domain_d.h:140-153)
struct StabilizationParams {
  double alpha_free = 0, alpha_contact = 0, hg_coeff_free = 0, hg_coeff_contact = 0, av_coeff_div = 0,
         av_coeff_bulk = 0, log_factor = 0, pspg_scale = 0, p_pspg_bulkfac = 0, J_min = 0, hg_visc = 0, hg_stiff = 0;
};

// TriMesh_d (include/common/Mesh.h:72-140, src/common/Mesh.C): the rigid tool surfaces, flattened the way the
// reference keeps them (all bodies in one node / facet list).  Built on the host, handed to the engine by
// Domain_d::setTriMesh; from then on the engine moves the surfaces itself (Solver_explicit.C:981-1005).
class TriMesh_d {
 public:
  int dimension = 3;
  int nodecount = 0, elemcount = 0, mesh_count = 0;
  std::vector<double> node, node_v, normal;  // xyz triples
  std::vector<int> elnode, ele_mesh_id;
  double3 m_v{0, 0, 0};
  double mu_sta[1] = {0.0}, mu_dyn[1] = {0.0};
  double heat_cond = 0.0, T_const = 0.0;
  // TriMesh_d::AxisPlaneMesh (Mesh.C:48-283)
  void AxisPlaneMesh(int id, int axis, bool positaxisorent, double3 p1, double3 p2, int dens) {
    int nn = 0, ne = 0;
    if (wf_host_axis_plane_counts(dimension, dens, &nn, &ne)) throw std::runtime_error("AxisPlaneMesh: bad arguments");
    const int nen = dimension == 3 ? 3 : 2;
    node.assign(3 * (size_t)nn, 0.0);
    node_v.assign(3 * (size_t)nn, 0.0);
    normal.assign(3 * (size_t)ne, 0.0);
    elnode.assign((size_t)nen * ne, 0);
    ele_mesh_id.assign(ne, id);
    double a[3] = {p1.x, p1.y, p1.z}, b[3] = {p2.x, p2.y, p2.z};
    if (wf_host_axis_plane_mesh(dimension, id, axis, positaxisorent ? 1 : 0, a, b, dens, node.data(), elnode.data(), normal.data(),
                                ele_mesh_id.data()))
      throw std::runtime_error("AxisPlaneMesh failed");
    nodecount = nn;
    elemcount = ne;
    mesh_count = 1;
  }
  void SetMeshVel(double3 v) { m_v = v; }  // Mesh.h:117
  void SetNodesVel(double3 v) {            // main.C:707-708
    m_v = v;
    for (int n = 0; n < nodecount; n++) { node_v[3 * n] = v.x; node_v[3 * n + 1] = v.y; node_v[3 * n + 2] = v.z; }
  }
  // TriMesh_d::AddMesh (Mesh.C:438-539): node ids offset by the nodes already present, new nodes move with m.m_v
  void AddMesh(const TriMesh_d &m) {
    if (m.dimension != dimension) throw std::runtime_error("AddMesh: dimension mismatch");
    for (int q : m.elnode) elnode.push_back(q + nodecount);
    node.insert(node.end(), m.node.begin(), m.node.end());
    for (int n = 0; n < m.nodecount; n++) { node_v.push_back(m.m_v.x); node_v.push_back(m.m_v.y); node_v.push_back(m.m_v.z); }
    normal.insert(normal.end(), m.normal.begin(), m.normal.end());
    ele_mesh_id.insert(ele_mesh_id.end(), m.ele_mesh_id.begin(), m.ele_mesh_id.end());
    nodecount += m.nodecount;
    elemcount += m.elemcount;
    mesh_count++;
  }
};

enum dom_type { _Plane_Strain_ = WF_PLANE_STRAIN, _Plane_Stress_ = WF_PLANE_STRESS, _Axi_Symm_ = WF_AXISYMM, _3D_ = WF_3D };

class Domain_d {
 public:
  explicit Domain_d(int device = 0) : device_(device) {}
  Domain_d(const Domain_d &) = delete;
  Domain_d &operator=(const Domain_d &) = delete;
  ~Domain_d() {
    if (eng_) wf_destroy(eng_);
  }

  // ---- setup, names of the reference ---------------------------------------------------------------
  void setAxiSymm(bool vol_weight = false) {  // Domain_d.h:666
    domtype_ = _Axi_Symm_;
    vol_weight_ = vol_weight;
  }
  void setDomType(dom_type t) { domtype_ = t; }
  void setStrict(bool on) { strict_ = on; }  // numerics flavour, wf_engine.h WF_STRICT / WF_FAST
  void setHexaHourglass(double c) { hexa_hg_ = c; }  // f90_ver/src/Mechanical.f90:307

  // Domain_d::AddBoxLength (src/common/Domain_d.C:1136)
  void AddBoxLength(double3 V, double3 L, double r, bool red_int = true, bool tritetra = false) {
    if (!red_int) throw std::runtime_error("full integration is not implemented by the reference step");
    int dim = L.z > 0.0 ? 3 : 2;
    int k = dim == 3 ? (tritetra ? 4 : 8) : (tritetra ? 3 : 4);
    create(dim, k);
    double v[3] = {V.x, V.y, V.z}, l[3] = {L.x, L.y, L.z};
    ck(wf_gen_box(eng_, v, l, r, tritetra ? 1 : 0));
    fetchCounts();
  }
  // Domain_d::CreateFromLSDyna path (Domain_d.C:1647): nodes + connectivity given by the caller
  void SetMesh(int dim, int nodxelem, int n_nodes, int n_elems, const double *x, const unsigned *elnod) {
    create(dim, nodxelem);
    ck(wf_set_mesh(eng_, n_nodes, n_elems, x, elnod));
    fetchCounts();
  }
  // Remesh hand-off (the engine-side half of ReMesher::WriteDomain, ReMesher.C:339): the current engine is dropped and
  // a new one is created on the new mesh; material, stabilisation, options, contact surfaces and the clock carry
  // over.  The caller then adds the boundary conditions of the new mesh (AddBCVelNode / AllocateBCs, SearchExtNodes
  // when contact is on), calls InitSolve() and uploads the mapped fields with set(): "u" "v" "prev_a" ("T"),
  // "m_tau" "pl_strain" "p" "sigma_y" "rho" "vol_0" — the arrays WriteDomain maps.  Uploading vol_0 / rho refreshes
  // the nodal reference sums on the device.
  void ReplaceMesh(int n_nodes, int n_elems, const double *x, const unsigned *elnod) {
    if (!eng_) throw std::runtime_error("ReplaceMesh: no mesh to replace");
    const int dim = dim_, k = nodxelem_;
    carry_clock_ = true;  // InitSolve() hands Time / step_count to the new engine
    wf_destroy(eng_);
    eng_ = nullptr;
    inited_ = false;
    SetMesh(dim, k, n_nodes, n_elems, x, elnod);
  }
  void setDensity(double rho) { rho0_ = rho; }  // Domain_d.C:951
  void setTemp(double T) { temp_ = T; }         // Domain_d::setTemp (uniform initial temperature, main.C:441)
  void setThermalOn() { m_thermal = true; }     // Domain_d.h:676
  double m_plheatfraction = 0.9;                // Domain_d.h:271 (config "plHeatFrac")
  double m_max_edot = 1.0e6;                    // Domain_d.h:824
  void AssignMaterial(const Material_ *mat) {    // Domain_d.C:903
    mat_ = *mat;
    have_mat_ = true;
  }
  StabilizationParams m_stab;                       // main.C:84-120
  int m_press_algorithm = 0;                        // main.C:211
  double m_artifvisc[2] = {0.0, 0.0};               // main.C:340-347
  void AddBCVelNode(int node, int dim, double val) { ck(wf_add_bc_vel(need(), node, dim, val)); }  // Domain_d.C:1057
  void AllocateBCs() { ck(wf_allocate_bcs(need())); }                                              // Domain_d.C:1063
  // ---- contact with rigid tool surfaces (main.C:636-848) -------------------------------------------
  void SearchExtNodes() { ck(wf_SearchExtNodes(need())); }  // Domain_d.C:110
  void setTriMesh(TriMesh_d *m) { trimesh = m; }            // Domain_d.h:464
  void addMeshData(const TriMesh_d &m) {                    // Domain_d.C:2780
    if (!trimesh) throw std::runtime_error("addMeshData: setTriMesh first");
    trimesh->AddMesh(m);
  }
  TriMesh_d *getTriMesh() { return trimesh; }
  void setContactPF(double pf) { m_contPF = pf; }           // Domain_d.h:771
  void setContactOn() { contact = true; }                   // Domain_d.h:642
  bool isContactOn() const { return contact; }
  void SetDT(double dt) { dt_ = dt; }            // Domain_d.h:629
  void SetEndTime(double t) { end_t_ = t; }      // Domain_d.h:630
  int getElemCount() const { return n_elems_; }
  int getNodeCount() const { return n_nodes_; }
  int getDim() const { return dim_; }
  int getNodxElem() const { return nodxelem_; }
  double getTime() const { return Time; }
  long getStepCount() const { return step_count; }

  // ---- solve -----------------------------------------------------------------------------------------
  // initialisation part of SolveChungHulbert (Solver_explicit.C:115-292)
  void InitSolve() {
    pushSettings();
    if (!(dt_ > 0.0)) throw std::runtime_error("SetDT first");
    if (contact) {  // main.C:700-725, :842-847, :862 (m_elem_length for the contact stiffness)
      if (!trimesh) throw std::runtime_error("contact is on but no TriMesh_d was set");
      ck(wf_set_trimesh(eng_, trimesh->dimension, trimesh->nodecount, trimesh->elemcount, trimesh->node.data(),
                        trimesh->node_v.data(), trimesh->elnode.data(), trimesh->normal.data(), trimesh->ele_mesh_id.data()));
      ck(wf_set_contact(eng_, trimesh->mu_sta[0], trimesh->mu_dyn[0], m_contPF, end_t_));
      if (m_thermal) ck(wf_set_contact_heat(eng_, trimesh->heat_cond, trimesh->T_const));
      double ml = 0, mh = 0;
      ck(wf_calcMinEdgeLength(eng_, &ml, &mh));
    }
    ck(wf_init(eng_, dt_));
    if (carry_clock_) { ck(wf_set_time(eng_, Time, step_count)); carry_clock_ = false; }
    inited_ = true;
  }
  // n fused steps (loop body, Solver_explicit.C:524-978)
  void Step(int n = 1) {
    ck(wf_step(eng_, n));
    ck(wf_get_time(eng_, &Time, &step_count));
  }
  // Domain_d::SolveChungHulbert (Solver_explicit.C:101): `while (Time < end_t)` with the fixed step of SetDT;
  // the non-finite internal-force scrub of :779-784 is reported like the reference does (printf and go on)
  void SolveChungHulbert() {
    if (!inited_) InitSolve();
    int n = 0;
    for (double t = Time; t < end_t_; t += dt_) n++;
    const int chunk = 1000;
    for (int done = 0; done < n; done += chunk) {
      Step(n - done < chunk ? n - done : chunk);
      int bad = 0;
      ck(wf_nonfinite_flag(eng_, &bad));
      if (bad) printf("Nonfinite internal force, step %ld\n", step_count);
    }
  }
  void computeEnergies(double *Ekin, double *dEint) { ck(wf_energies(need(), Ekin, dEint)); }  // Mechanical.C:2145
  void calcMinEdgeLength(double *min_len, double *min_h) { ck(wf_calcMinEdgeLength(need(), min_len, min_h)); }
  double cflDt(double factor) {  // Solver_explicit.C:579-598
    double d = 0;
    ck(wf_cfl_dt(need(), factor, &d));
    return d;
  }

  // ---- state: names = Domain_d member names, reference layouts --------------------------------------
  std::vector<double> get(const char *name) {
    size_t nb = wf_array_bytes(need(), name);
    std::vector<double> out(nb / sizeof(double));
    ck(wf_get_array(eng_, name, out.data(), nb));
    return out;
  }
  std::vector<int> getInt(const char *name) {
    size_t nb = wf_array_bytes(need(), name);
    std::vector<int> out(nb / sizeof(int));
    ck(wf_get_array(eng_, name, out.data(), nb));
    return out;
  }
  void set(const char *name, const std::vector<double> &v) { ck(wf_set_array(need(), name, v.data(), v.size() * sizeof(double))); }
  wf_engine *handle() { return eng_; }

  double Time = 0.0;
  long step_count = 0;

 private:
  wf_engine *need() {
    if (!eng_) throw std::runtime_error("no mesh yet (AddBoxLength / SetMesh first)");
    return eng_;
  }
  void ck(int rc) {
    if (rc) {
      const char *m = wf_last_error(eng_);
      throw std::runtime_error(std::string("wf_b200: ") + (m ? m : "error"));
    }
  }
  void create(int dim, int k) {
    if (eng_) throw std::runtime_error("mesh already created");
    int dt = dim == 3 ? (int)_3D_ : (domtype_ == _3D_ ? (int)_Plane_Strain_ : (int)domtype_);
    int rc = wf_create(&eng_, dim, k, dt, device_);
    if (rc) {
      const char *m = wf_last_error(nullptr);
      eng_ = nullptr;
      throw std::runtime_error(std::string("wf_b200: ") + (m ? m : "wf_create failed"));
    }
    dim_ = dim;
    nodxelem_ = k;
    if (dt == _Axi_Symm_ && vol_weight_) ck(wf_set_axisymm_vol_weight(eng_, 1));
  }
  void fetchCounts() {
    int tot = 0;
    ck(wf_get_counts(eng_, &n_nodes_, &n_elems_, &tot));
  }
  void pushSettings() {
    need();
    if (!have_mat_) throw std::runtime_error("AssignMaterial first");
    wf_material m{};
    m.model = mat_.Material_model;
    m.E = mat_.Elastic().E();
    m.nu = mat_.Elastic().Poisson();
    m.rho0 = rho0_;
    m.sy0 = mat_.sy0;
    m.K = mat_.K;
    m.m = mat_.m;
    if (mat_.Material_model == JOHNSON_COOK) {
      const double q[8] = {mat_.A, mat_.B, mat_.n, mat_.C, mat_.eps_0, mat_.m, mat_.T_m, mat_.T_t};
      for (int i = 0; i < 8; i++) m.q[i] = q[i];
    } else if (mat_.Material_model == GMT) {
      const double q[14] = {mat_.n1, mat_.n2, mat_.C1, mat_.C2, mat_.m1, mat_.m2, mat_.I1, mat_.I2,
                            mat_.e_min, mat_.e_max, mat_.er_min, mat_.er_max, mat_.T_min, mat_.T_max};
      for (int i = 0; i < 14; i++) m.q[i] = q[i];
    }
    m.temp = temp_;
    m.max_edot = m_max_edot;
    ck(wf_set_material(eng_, &m));
    wf_stab s{m_stab.alpha_free, m_stab.alpha_contact, m_stab.hg_coeff_free, m_stab.hg_coeff_contact,
              m_stab.av_coeff_div, m_stab.av_coeff_bulk, m_stab.log_factor, m_stab.pspg_scale,
              m_stab.p_pspg_bulkfac, m_stab.J_min, m_stab.hg_visc, m_stab.hg_stiff, hexa_hg_};
    ck(wf_set_stab(eng_, &s));
    ck(wf_set_options(eng_, m_press_algorithm, m_artifvisc[0], m_artifvisc[1], strict_ ? WF_STRICT : WF_FAST));
    if (m_thermal) ck(wf_set_thermal(eng_, mat_.k_T, mat_.cp_T, mat_.exp_T, m_plheatfraction, temp_));
  }

  TriMesh_d *trimesh = nullptr;
  bool contact = false, m_thermal = false;
  double m_contPF = 0.1;  // Domain_d.h:256
  wf_engine *eng_ = nullptr;
  int device_ = 0, dim_ = 0, nodxelem_ = 0, n_nodes_ = 0, n_elems_ = 0;
  dom_type domtype_ = _3D_;
  bool vol_weight_ = false, strict_ = false, have_mat_ = false, inited_ = false;
  double rho0_ = 0.0, dt_ = 0.0, end_t_ = 0.0, hexa_hg_ = 0.0, temp_ = 20.0;
  bool carry_clock_ = false;
  Material_ mat_;
};

}  // namespace wf_b200
