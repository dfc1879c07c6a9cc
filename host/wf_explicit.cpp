// wf_explicit.cpp — C++ driver of the explicit loop for the synthetic box workloads of BASELINE.json
// (SURVEY.md §8d): the host program a WeldFormFEM user would write against wf_domain.hpp.  Builds the mesh with
// AddBoxLength, clamps the bottom plane, prescribes (0,..,v_top) on the top plane, runs `steps` fused explicit
// steps on the GPU and (optionally) dumps reference-layout arrays for the parity tests.
//
//   wf_explicit --kind hex|tet|quad|tri|axiquad --n N [--steps S] [--vtop V] [--hg C] [--press P] [--strict]
//               [--cfl F] [--dump FILE] [--time] [--contact]
// --contact (tet / quad): instead of prescribing the top plane, a rigid plane / line comes down on it with velocity
// (0.5, .., vtop), friction 0.3 / 0.2 and penalty factor 0.6, like examples/input/Contact_Compression_*.json.
// dump format: for every array  "<name> <count>\n" followed by <count> raw little-endian doubles.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "wf_domain.hpp"

using namespace wf_b200;

int main(int argc, char **argv) {
  std::string kind = "hex", dump;
  int n = 8, steps = 10, press = 0;
  double vtop = -10.0, hg = -1.0, cfl = -1.0;
  bool strict = false, timeit = false, contact = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--kind") kind = val();
    else if (a == "--n") n = atoi(val());
    else if (a == "--steps") steps = atoi(val());
    else if (a == "--vtop") vtop = atof(val());
    else if (a == "--hg") hg = atof(val());
    else if (a == "--press") press = atoi(val());
    else if (a == "--cfl") cfl = atof(val());
    else if (a == "--dump") dump = val();
    else if (a == "--strict") strict = true;
    else if (a == "--time") timeit = true;
    else if (a == "--contact") contact = true;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  bool three_d = kind == "hex" || kind == "tet";
  bool tritet = kind == "tet" || kind == "tri";
  double h = three_d ? 1.0e-3 : 0.5e-3;                       // element edge, weldformfem_b200/cases.py
  if (hg < 0.0) hg = kind == "hex" ? 0.06 : 0.0;
  if (cfl < 0.0) cfl = tritet ? 0.1 : 0.3;
  try {
    Domain_d dom(0);
    if (kind == "axiquad") dom.setAxiSymm(false);
    else if (!three_d) dom.setDomType(_Plane_Strain_);
    const double pad = 1.0 + 1.0e-6;                          // keeps (int)(L / (2 r)) == n under rounding
    double3 L = make_double3(n * h * pad, n * h * pad, three_d ? n * h * pad : 0.0);
    dom.AddBoxLength(make_double3(0, 0, 0), L, 0.5 * h, true, tritet);

    // Hollomon aluminium of examples/input/Compression_hexa_hollomon.json
    const double E = 68.9e9, nu = 0.3, rho = 2700.0;
    Elastic_ el(E, nu);
    Material_ mat(el);
    mat.InitHollomon(el, 190.4e6, 386.796e6, 0.154);
    mat.cs0 = sqrt(el.BulkMod() / rho);
    dom.setDensity(rho);
    dom.AssignMaterial(&mat);
    dom.setHexaHourglass(hg);
    dom.setStrict(strict);
    dom.m_press_algorithm = press;
    if (!three_d && !tritet) dom.m_stab.hg_visc = dom.m_stab.hg_stiff = 0.1;

    int dim = dom.getDim();
    int nplane = (n + 1) * (dim == 3 ? n + 1 : 1);
    for (int nd = 0; nd < nplane; nd++)
      for (int d = 0; d < dim; d++) dom.AddBCVelNode(nd, d, 0.0);
    if (!contact)
      for (int nd = nplane * n; nd < nplane * (n + 1); nd++)
        for (int d = 0; d < dim; d++) dom.AddBCVelNode(nd, d, d == dim - 1 ? vtop : 0.0);
    dom.AllocateBCs();

    double dt = cfl * h / mat.cs0;                            // src/explicit/main.C:862-879
    dom.SetDT(dt);
    TriMesh_d msh;
    if (contact) {                                            // src/explicit/main.C:636-848
      dom.SearchExtNodes();
      const double Lb = n * h, top = Lb + 0.01 * h;
      msh.dimension = dim;
      if (dim == 3) msh.AxisPlaneMesh(0, 2, false, make_double3(-0.5 * Lb, -0.5 * Lb, top), make_double3(1.5 * Lb, 1.5 * Lb, top), 4);
      else msh.AxisPlaneMesh(0, 1, false, make_double3(-0.5 * Lb, top, 0.0), make_double3(1.5 * Lb, top, 0.0), 4);
      msh.SetNodesVel(dim == 3 ? make_double3(0.5, 0.0, vtop) : make_double3(0.5, vtop, 0.0));
      msh.mu_sta[0] = 0.3;
      msh.mu_dyn[0] = 0.2;
      dom.setTriMesh(&msh);
      dom.setContactPF(0.6);
      dom.setContactOn();
      dom.SetEndTime(100 * dt);
    }
    dom.InitSolve();
    auto t0 = std::chrono::steady_clock::now();
    dom.Step(steps);
    double ek = 0, de = 0;
    dom.computeEnergies(&ek, &de);                            // synchronises
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("kind %s n %d elements %d nodes %d steps %ld time %.17g Ekin %.17g\n", kind.c_str(), n, dom.getElemCount(),
           dom.getNodeCount(), dom.getStepCount(), dom.getTime(), ek);
    if (timeit) printf("wall %.6f s  %.4g element-steps/s\n", sec, (double)dom.getElemCount() * steps / sec);
    if (!dump.empty()) {
      FILE *f = fopen(dump.c_str(), "wb");
      if (!f) { perror("dump"); return 1; }
      std::vector<const char *> names = {"x", "v", "a", "u", "prev_a", "m_fi", "m_mdiag", "vol", "p", "pl_strain", "sigma_y", "m_sigma", "m_tau"};
      if (contact) { names.push_back("contforce"); names.push_back("ut_prev"); names.push_back("node_area"); }
      for (const char *nm : names) {
        std::vector<double> q = dom.get(nm);
        fprintf(f, "%s %zu\n", nm, q.size());
        fwrite(q.data(), sizeof(double), q.size(), f);
      }
      fclose(f);
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
