// wf_weldform.cpp — run a WeldFormFEM input deck on the B200 engine: the counterpart of the reference's
// `WeldFormFEM deck.json` (src/explicit/main.C) with the explicit loop executed by libwf_b200.so.
//
//   wf_weldform deck.json [--steps N] [--parse-only] [--dump FILE] [--strict] [--hexa-hg C]
//
// Without --steps the loop runs `while (Time < simTime)` like Domain_d::SolveChungHulbert.  --parse-only reads and
// checks deck + mesh without touching the GPU and prints the summary line.  --dump writes reference-layout arrays:
// "<name> <count>\n" followed by <count> raw little-endian doubles each.  The reference's VTK / CSV output is not
// reproduced; the arrays above are what a writer needs.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "wf_deck.hpp"

using namespace wf_b200;

int main(int argc, char **argv) {
  std::string deck, dump;
  int steps = -1;
  bool parse_only = false, strict = false;
  double hexa_hg = 0.0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--steps") steps = atoi(val());
    else if (a == "--dump") dump = val();
    else if (a == "--parse-only") parse_only = true;
    else if (a == "--strict") strict = true;
    else if (a == "--hexa-hg") hexa_hg = atof(val());
    else if (a[0] != '-') deck = a;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (deck.empty()) { fprintf(stderr, "usage: wf_weldform deck.json [--steps N] [--parse-only] [--dump FILE] [--strict] [--hexa-hg C]\n"); return 2; }
  try {
    Domain_d dom(0);
    TriMesh_d msh;
    DeckSummary S = setup_from_deck(deck, dom, msh, parse_only, hexa_hg, strict);
    printf("{\"deck\": \"%s\", \"dim\": %d, \"nodxelem\": %d, \"nodes\": %d, \"elements\": %d, \"material\": \"%s\", \"bc_nodes\": %d, \"bc_count\": [%d, %d, %d], "
           "\"sym_nodes\": %d, \"contact\": %s, \"rigid_bodies\": %d, \"rigid_facets\": %d, \"thermal\": %s, \"end_time\": %.17g",
           deck.c_str(), S.dim, S.nodxelem, S.n_nodes, S.n_elems, S.material.c_str(), S.bc_nodes, S.bc_count[0], S.bc_count[1], S.bc_count[2], S.sym_nodes,
           S.contact ? "true" : "false", S.rigid_bodies, S.rigid_facets, S.thermal ? "true" : "false", S.end_time);
    if (parse_only) { printf("}\n"); return 0; }
    dom.InitSolve();
    auto t0 = std::chrono::steady_clock::now();
    if (steps >= 0) dom.Step(steps);
    else dom.SolveChungHulbert();
    double ek = 0, de = 0;
    dom.computeEnergies(&ek, &de);  // synchronises
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf(", \"min_length\": %.17g, \"dt\": %.17g, \"steps\": %ld, \"time\": %.17g, \"Ekin\": %.17g, \"wall_s\": %.6f, "
           "\"element_steps_per_s\": %.4g}\n", S.min_length, S.dt, dom.getStepCount(), dom.getTime(), ek, sec,
           (double)S.n_elems * (double)dom.getStepCount() / sec);
    if (!dump.empty()) {
      FILE *f = fopen(dump.c_str(), "wb");
      if (!f) { perror("dump"); return 1; }
      std::vector<const char *> names = {"x", "v", "a", "u", "prev_a", "m_fi", "m_mdiag", "vol", "p", "pl_strain", "sigma_y", "m_sigma", "m_tau"};
      if (S.contact) { names.push_back("contforce"); names.push_back("ut_prev"); names.push_back("node_area"); names.push_back("trimesh.node"); }
      if (S.thermal) { names.push_back("T"); names.push_back("m_q_plheat"); }
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
