// wf_deck.hpp — the reference's JSON / LS-Dyna `.k` front-end over wf_domain.hpp: reads a WeldFormFEM
// input deck (examples/input/*.json) and drives the B200 engine through the same sequence of Domain_d calls as
// src/explicit/main.C:125-1000 does for the CPU solver.
//
// Parity status: the reference's own `.k` reader is the un-served submodule lib/LSDynaReader, so the
// `.k` path is PARITY UNPINNED; the reader here follows the keyword format visible in examples/input/*.k (node ids
// mapped to 0-based indices in order of appearance, tetrahedra written as degenerate 8-node solids truncated at the
// first repeated node).  Everything after the mesh is pinned indirectly: tests drive the oracle with the same settings
// and compare the resulting state (tests/test_deck.py).
//
// Deviations from main.C, on purpose:
//   * "JohnsonCook" / "GMT" materials: main.C:535-558 leaves the fields the step reads uninitialised (private members
//     shadow the public ones, Material.cuh:176-274); here the constants land in the public fields (the behaviour the
//     free functions of Material.cuh:377-483 define).
//   * a "File" rigid body (Nastran surface, main.C:686-697) and remeshing ("Meshing" block) are not supported.
//   * hexahedral decks: main.C:650 calls SearchExtNodes on every mesh, which overruns `elements[ELNOD]` for 8-node
//     elements (Domain_d.C:116-127); it is only called here when contact is requested.
#pragma once

#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "wf_domain.hpp"
#include "wf_json.hpp"

namespace wf_b200 {

// ---- LS-Dyna keyword mesh: *NODE and *ELEMENT_SOLID (examples/input/tetra_cyl.k, cyl_hex.k) -----------------------
struct KMesh {
  std::vector<double> x;         // xyz per node
  std::vector<unsigned> elnod;   // nodxelem ids per element, 0-based
  int nodxelem = 0;
  int n_nodes() const { return (int)(x.size() / 3); }
  int n_elems() const { return nodxelem ? (int)(elnod.size() / nodxelem) : 0; }
};

inline std::vector<std::string> k_fields(const std::string &line, const int *widths, int nw) {
  std::vector<std::string> out;
  if (line.find(',') != std::string::npos) {  // free format
    std::stringstream ss(line);
    std::string tok;
    while (std::getline(ss, tok, ',')) out.push_back(tok);
    return out;
  }
  size_t pos = 0;
  for (int i = 0; i < nw && pos < line.size(); i++) {  // fixed format
    out.push_back(line.substr(pos, (size_t)widths[i]));
    pos += (size_t)widths[i];
  }
  return out;
}

inline KMesh read_k(const std::string &path, double scale = 1.0) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  KMesh m;
  std::map<long, unsigned> index_of;
  std::vector<std::vector<long>> elems;
  enum { NONE, NODE, SOLID } sect = NONE;
  static const int wn[] = {8, 16, 16, 16, 8, 8}, we[] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 8};
  std::string line;
  std::vector<long> pending;  // "eid pid" line of the two-line solid format
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '$') continue;
    if (line[0] == '*') {
      if (line.compare(0, 5, "*NODE") == 0 && line.compare(0, 6, "*NODE_") != 0) sect = NODE;
      else if (line.compare(0, 14, "*ELEMENT_SOLID") == 0) sect = SOLID;
      else sect = NONE;
      continue;
    }
    if (sect == NODE) {
      auto t = k_fields(line, wn, 6);
      if (t.size() < 4) continue;
      long id = atol(t[0].c_str());
      index_of[id] = (unsigned)(m.x.size() / 3);
      for (int c = 1; c <= 3; c++) m.x.push_back(atof(t[c].c_str()) * scale);
    } else if (sect == SOLID) {
      auto t = k_fields(line, we, 10);
      std::vector<long> v;
      for (auto &q : t) {
        if (q.find_first_not_of(" \t") == std::string::npos) continue;
        v.push_back(atol(q.c_str()));
      }
      if (v.empty()) continue;
      if (pending.empty() && v.size() == 2) { pending = v; continue; }  // nodes follow on the next line
      std::vector<long> nodes;
      if (!pending.empty()) { nodes = v; pending.clear(); }
      else nodes.assign(v.begin() + 2, v.end());
      std::vector<long> uniq;  // degenerate solids: cut at the first repeated node
      for (long n : nodes) {
        bool seen = false;
        for (long u : uniq) seen = seen || (u == n);
        if (seen) break;
        uniq.push_back(n);
      }
      elems.push_back(uniq);
    }
  }
  if (elems.empty()) throw std::runtime_error(path + ": no *ELEMENT_SOLID records");
  m.nodxelem = (int)elems[0].size();  // Domain_d::CreateFromLSDyna, Domain_d.C:1653
  if (m.nodxelem != 4 && m.nodxelem != 8) throw std::runtime_error(path + ": solids must be tetrahedra or hexahedra");
  for (auto &e : elems) {
    if ((int)e.size() != m.nodxelem) throw std::runtime_error(path + ": mixed element types");
    for (long n : e) {
      auto it = index_of.find(n);
      if (it == index_of.end()) throw std::runtime_error(path + ": element references an unknown node id");
      m.elnod.push_back(it->second);
    }
  }
  return m;
}

// ---- the deck ---------------------------------------------------------------------------------------------------
struct DeckBC {
  int zoneId = 0, valueType = 0;
  double3 value{0, 0, 0}, start{0, 0, 0}, end{0, 0, 0};
};

struct DeckSummary {
  int dim = 0, nodxelem = 0, n_nodes = 0, n_elems = 0, bc_nodes = 0, sym_nodes = 0, rigid_bodies = 0, rigid_facets = 0;
  int bc_count[3] = {0, 0, 0};  // entries per direction list, as Domain_d::bc_count after AllocateBCs
  bool contact = false, thermal = false;
  double dt = 0, end_time = 0, min_length = 0, out_time = 0;
  std::string material;
};

inline bool readVector(const Json &j, double3 &v) {  // Input.h:54-63
  if (j.is_null()) return false;
  if (j.type != Json::Array || j.arr.size() < 3) throw std::runtime_error("JSON: 3-vector expected");
  v = make_double3(j.arr[0].num, j.arr[1].num, j.arr[2].num);
  return true;
}

// Everything main.C does between reading the deck and calling SolveChungHulbert (main.C:191-975).  `dom` must be a
// fresh Domain_d; `msh` receives the rigid surfaces and must outlive the solve.  With parse_only no engine call that
// needs a GPU is made (mesh and deck are read and checked, counts reported).
inline DeckSummary setup_from_deck(const std::string &deck_path, Domain_d &dom, TriMesh_d &msh, bool parse_only = false,
                                   double hexa_hg = 0.0, bool strict = false) {
  const Json j = Json::parse_file(deck_path);
  const Json &config = j["Configuration"], &material = j["Materials"], &domblock = j["DomainBlocks"];
  const Json &rigbodies = j["RigidBodies"], &contact_ = j["Contact"], &bcs = j["BoundaryConditions"], &ics = j["InitialConditions"];
  std::string dir = deck_path.substr(0, deck_path.find_last_of("/\\") + 1);
  DeckSummary S;

  // loadStabilizationParams, main.C:84-120: with a "Stabilization" block EVERY field is overwritten (absent keys -> 0),
  // without one the constructor defaults stay (all 0, hg_stiff 0.1; Domain_d.h:283-296).  pspg_scale is the exception:
  // main.C never assigns it in its local struct, so the reference passes an INDETERMINATE value on to
  // calcElemPressure (Mechanical.C:796 reads it on the 3D path when div v < 0).  Deliberate deviation: this
  // front-end pins it to 0 unless the deck carries an explicit "pspg_scale" key.
  if (j.contains("Stabilization") && !j["Stabilization"].is_null()) {
    const Json &st = j["Stabilization"];
    StabilizationParams p;
    p.alpha_free = st.value("alpha_free", 0.0); p.alpha_contact = st.value("alpha_contact", 0.0);
    p.hg_coeff_free = st.value("hg_coeff_free", 0.0); p.hg_coeff_contact = st.value("hg_coeff_contact", 0.0);
    p.av_coeff_div = st.value("av_coeff_div", 0.0); p.av_coeff_bulk = st.value("av_coeff_bulk", 0.0);
    p.log_factor = st.value("log_factor", 0.0); p.p_pspg_bulkfac = st.value("p_pspg_bulkfac", 0.0);
    p.J_min = st.value("J_min", 0.0); p.hg_visc = st.value("hg_visc", 0.0); p.hg_stiff = st.value("hg_stiff", 0.0);
    p.pspg_scale = st.value("pspg_scale", 0.0);
    dom.m_stab = p;
  } else {
    dom.m_stab.hg_stiff = 0.1;
  }
  double out_time = 0, sim_time = 0, cflFactor = 0.3;
  bool fixedTS = false;
  readValue(config["outTime"], out_time);
  readValue(config["simTime"], sim_time);
  readValue(config["plHeatFrac"], dom.m_plheatfraction);
  readValue(config["maxStrRate"], dom.m_max_edot);
  std::string plType = "Hardening";
  readValue(config["plasticType"], plType);
  if (plType != "Hardening") throw std::runtime_error("plasticType '" + plType + "' is not supported (Hardening only)");
  readValue(config["fixedTS"], fixedTS);
  readValue(config["cflFactor"], cflFactor);
  std::vector<double> av;
  if (readArray(config["artifViscCoeffs"], av) && av.size() >= 2) { dom.m_artifvisc[0] = av[0]; dom.m_artifvisc[1] = av[1]; }
  std::string dom_type = "3D";
  readValue(config["domType"], dom_type);
  bool xyzsym[3] = {false, false, false};
  double symtol = 1.0e-4;
  readValue(config["symtol"], symtol);
  readValue(config["xSymm"], xyzsym[0]); readValue(config["ySymm"], xyzsym[1]); readValue(config["zSymm"], xyzsym[2]);
  if (dom_type == "AxiSymm" || dom_type == "AxiSym") {
    bool vol_weight = false;
    readValue(config["AxiSymmVol"], vol_weight);
    dom.setAxiSymm(vol_weight);
  } else if (dom_type == "plStrain") {
    dom.setDomType(_Plane_Strain_);
  }
  int press_alg = 0;
  readValue(config["pressAlgorithm"], press_alg);
  // Solver_explicit.C:733-743 dispatches on 0 and 1 only (any other value would skip the pressure update altogether)
  if (press_alg != 0 && press_alg != 1) throw std::runtime_error("pressAlgorithm must be 0 or 1");
  bool dev_elastic = true;
  readValue(config["devElastic"], dev_elastic);
  if (!dev_elastic) throw std::runtime_error("devElastic = false (calcElemPressureRigid) is not supported");
  if (press_alg > 0) dom.m_press_algorithm = press_alg;
  dom.setHexaHourglass(hexa_hg);
  dom.setStrict(strict);

  // domain, main.C:352-422
  std::string domtype = "Box";
  readValue(domblock[0]["type"], domtype);
  KMesh km;
  double3 box_start{0, 0, 0}, box_L{0, 0, 0};
  double box_dx = 0.06;
  bool tritet = false;
  if (domtype == "File") {
    std::string filename;
    readValue(domblock[0]["fileName"], filename);
    if (filename.size() < 2 || filename.substr(filename.find_last_of('.') + 1) != "k")
      throw std::runtime_error("DomainBlocks[0].fileName must be an LS-Dyna .k file");
    km = read_k(filename[0] == '/' ? filename : dir + filename);
    S.dim = 3; S.nodxelem = km.nodxelem; S.n_nodes = km.n_nodes(); S.n_elems = km.n_elems();
    if (!parse_only) dom.SetMesh(3, km.nodxelem, km.n_nodes(), km.n_elems(), km.x.data(), km.elnod.data());
  } else if (domtype == "Box") {
    readVector(domblock[0]["dim"], box_L);
    readVector(domblock[0]["start"], box_start);
    readValue(domblock[0]["elemLength"], box_dx);
    std::string eltype;
    readValue(domblock[0]["elemType"], eltype);
    tritet = eltype == "TriTet";
    // main.C:418: the z extent is dropped -> JSON boxes are 2D
    const double L[3] = {box_L.x, box_L.y, 0.0};
    int d2 = 0, k2 = 0, nn = 0, ne = 0;
    wf_host_box_counts(L, box_dx / 2., tritet ? 1 : 0, &d2, &k2, &nn, &ne);
    S.dim = d2; S.nodxelem = k2; S.n_nodes = nn; S.n_elems = ne;
    if (!parse_only) dom.AddBoxLength(box_start, make_double3(box_L.x, box_L.y, 0.0), box_dx / 2., true, tritet);
  } else {
    throw std::runtime_error("DomainBlocks[0].type must be File or Box");
  }

  // material, main.C:455-581
  double E = 0, nu = 0, rho = 0, Fy = 0.0;
  readValue(material[0]["density0"], rho);
  readValue(material[0]["youngsModulus"], E);
  readValue(material[0]["poissonsRatio"], nu);
  std::vector<double> e_range{0.0, 1.0e10}, er_range{0.0, 1.0e10}, T_range{0.0, 1.0e10}, c(10, 0.0);
  readArray(material[0]["strRange"], e_range);
  readArray(material[0]["strdotRange"], er_range);
  readArray(material[0]["tempRange"], T_range);
  std::string mattype = "Bilinear";
  readValue(material[0]["type"], mattype);
  readValue(material[0]["yieldStress0"], Fy);
  readArray(material[0]["const"], c);
  c.resize(10, 0.0);
  double IniTemp = 20.;
  for (size_t q = 0; q < ics.size(); q++) readValue(ics[q]["Temp"], IniTemp);
  bool thermal = false;
  readValue(config["thermal"], thermal);
  Elastic_ el(E, nu);
  Material_ mat(el);
  const double mat_cs = sqrt(el.BulkMod() / rho);
  if (mattype == "Bilinear") {
    mat.Material_model = BILINEAR;
    mat.Ep = E * c[0] / (E - c[0]);
  } else if (mattype == "Hollomon") {
    mat.InitHollomon(el, Fy, c[0], c[1]);
  } else if (mattype == "JohnsonCook") {  // argument order of main.C:539-541: (A = Fy, B, n, C, eps_0, m, T_m, T_t)
    if (thermal) mat.Init_JohnsonCook(el, Fy, c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
    else mat.Init_JohnsonCook(el, Fy, c[0], c[1], c[2], c[3], 1.0, 1.0e10, 0.0);
  } else if (mattype == "GMT") {  // main.C:549-553: n1 n2 C1 C2 m1 m2 I1 I2 + ranges
    mat.Material_model = GMT;
    mat.n1 = c[0]; mat.n2 = c[1]; mat.C1 = c[2]; mat.C2 = c[3]; mat.m1 = c[4]; mat.m2 = c[5]; mat.I1 = c[6]; mat.I2 = c[7];
    mat.e_min = e_range[0]; mat.e_max = e_range[1]; mat.er_min = er_range[0]; mat.er_max = er_range[1];
    mat.T_min = T_range[0]; mat.T_max = T_range[1];
  } else {
    throw std::runtime_error("material type '" + mattype + "' is not supported");
  }
  mat.cs0 = mat_cs;
  mat.sy0 = Fy;  // main.C:571
  readValue(material[0]["thermalCond"], mat.k_T);
  readValue(material[0]["thermalHeatCap"], mat.cp_T);
  readValue(material[0]["thermalExp"], mat.exp_T);
  dom.setDensity(rho);
  dom.AssignMaterial(&mat);
  dom.setTemp(IniTemp);
  if (thermal) dom.setThermalOn();
  S.material = mattype;
  S.thermal = thermal;

  // boundary conditions, main.C:585-631
  std::vector<DeckBC> bConds;
  for (size_t q = 0; q < bcs.size(); q++) {
    DeckBC b;
    readValue(bcs[q]["zoneId"], b.zoneId);
    readValue(bcs[q]["valueType"], b.valueType);
    readVector(bcs[q]["value"], b.value);
    readVector(bcs[q]["start"], b.start);
    readVector(bcs[q]["end"], b.end);
    bConds.push_back(b);
  }

  // rigid bodies + contact, main.C:636-848
  std::string rb_type;
  const bool contact = readValue(rigbodies[0]["type"], rb_type);
  S.contact = contact;
  auto make_body = [&](const Json &rb, int id, TriMesh_d &m) {
    double3 start{0, 0, 0}, dim_{0, 0, 0};
    bool flipnormals = false;
    int partSide = 1;
    std::string type;
    readVector(rb["start"], start);
    readVector(rb["dim"], dim_);
    readValue(rb["flipnormals"], flipnormals);
    readValue(rb["partSide"], partSide);
    readValue(rb["type"], type);
    if (type == "Plane") {
      m.dimension = 3;
      m.AxisPlaneMesh(id, 2, !flipnormals, start, dim_, partSide);  // p2 = `dim` as given, main.C:676
    } else if (type == "Line") {
      m.dimension = 2;
      if (dim_.x > 0.0) m.AxisPlaneMesh(id, 1, !flipnormals, start, dim_, partSide);
      else if (dim_.y > 0.0) m.AxisPlaneMesh(id, 0, !flipnormals, start, dim_, partSide);
      else throw std::runtime_error("rigid Line has null dimension");
    } else {
      throw std::runtime_error("rigid body type '" + type + "' is not supported (Plane, Line)");
    }
  };
  if (contact) {
    if (rigbodies.size() > 2) throw std::runtime_error("at most two rigid bodies, like main.C");
    int id0 = 0;
    readValue(rigbodies[0]["zoneId"], id0);
    make_body(rigbodies[0], 0, msh);
    for (auto &b : bConds)
      if (b.zoneId == id0) msh.SetNodesVel(b.value);  // main.C:703-709
    if (rigbodies.size() > 1) {
      int id1 = 0;
      readValue(rigbodies[1]["zoneId"], id1);
      TriMesh_d m2;
      make_body(rigbodies[1], 1, m2);
      for (auto &b : bConds)
        if (b.zoneId == id1) m2.SetMeshVel(b.value);  // main.C:813-817
      msh.AddMesh(m2);
    }
    double penaltyfac = -1.0;
    msh.T_const = 20.;
    readValue(contact_[0]["fricCoeffStatic"], msh.mu_sta[0]);
    readValue(contact_[0]["fricCoeffDynamic"], msh.mu_dyn[0]);
    readValue(contact_[0]["heatCondCoeff"], msh.heat_cond);
    readValue(contact_[0]["dieTemp"], msh.T_const);
    readValue(contact_[0]["penaltyFactor"], penaltyfac);
    S.rigid_bodies = msh.mesh_count;
    S.rigid_facets = msh.elemcount;
    if (!parse_only) {
      dom.SearchExtNodes();  // main.C:650
      dom.setTriMesh(&msh);
      if (penaltyfac > -1.0) dom.setContactPF(penaltyfac);
      dom.setContactOn();
    }
  }

  // node coordinates for the zone / symmetry searches
  std::vector<double> X;
  if (!parse_only) X = dom.get("x");
  else if (domtype == "File") X = km.x;
  else {
    const double V[3] = {box_start.x, box_start.y, box_start.z}, L[3] = {box_L.x, box_L.y, 0.0};
    X.resize((size_t)S.n_nodes * S.dim);
    std::vector<unsigned> el((size_t)S.n_elems * S.nodxelem);
    wf_host_gen_box(V, L, box_dx / 2., tritet ? 1 : 0, X.data(), el.data());
  }
  const int dim = S.dim, nn = S.n_nodes;
  auto add_bc = [&](int n, int d, double v) { S.bc_count[d]++; if (!parse_only) dom.AddBCVelNode(n, d, v); };
  if (!contact) {  // Domain_d::AddBCVelZone (Domain_d.C:432-452) for every BC, main.C:737-749
    for (auto &b : bConds)
      for (int a = 0; a < nn; a++) {
        bool in = !(X[(size_t)dim * a] < b.start.x || X[(size_t)dim * a] > b.end.x) &&
                  !(X[(size_t)dim * a + 1] < b.start.y || X[(size_t)dim * a + 1] > b.end.y);
        if (dim > 2) in = in && !(X[(size_t)dim * a + 2] < b.start.z || X[(size_t)dim * a + 2] > b.end.z);
        if (!in) continue;
        add_bc(a, 0, b.value.x); add_bc(a, 1, b.value.y);
        if (dim > 2) add_bc(a, 2, b.value.z);
        S.bc_nodes++;
      }
  }

  // time step, main.C:862-883: dt = cflFactor * min edge length / sqrt(K / rho)
  double dx = 0, mh = 0;
  if (!parse_only) {
    dom.calcMinEdgeLength(&dx, &mh);
    S.min_length = dx;
    S.dt = cflFactor * dx / mat_cs;
    dom.SetDT(S.dt);
  }
  dom.SetEndTime(sim_time);
  S.end_time = sim_time;
  S.out_time = out_time;
  (void)out_time; (void)fixedTS;

  // symmetry planes, main.C:947-963
  for (int i = 0; i < nn; i++)
    for (int d = 0; d < 3; d++)
      if (xyzsym[d]) {
        const double coord = d < dim ? X[(size_t)dim * i + d] : 0.0;  // getPosVec3: z = 0 in 2D
        if (coord < symtol) {
          if (d < dim) add_bc(i, d, 0);  // a z list in 2D is never imposed (ImposeBCV runs over d < m_dim)
          S.sym_nodes++;
        }
      }
  if (!parse_only) dom.AllocateBCs();
  return S;
}

}  // namespace wf_b200
