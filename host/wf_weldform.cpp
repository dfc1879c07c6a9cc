// wf_weldform.cpp — run a WeldFormFEM input deck on the B200 engine: the counterpart of the reference's
// `WeldFormFEM deck.json` (src/explicit/main.C) with the explicit loop executed by libwf_b200.so.
//
//   wf_weldform deck.json [--steps N] [--parse-only] [--dump FILE] [--vtk FILE] [--vtk-ascii] [--out BASE] [--strict]
//                         [--hexa-hg C]
//
// Without --steps the loop runs `while (Time < simTime)` like Domain_d::SolveChungHulbert.  --parse-only reads and
// checks deck + mesh without touching the GPU and prints the summary line.  --dump writes reference-layout arrays:
// "<name> <count>\n" followed by <count> raw little-endian doubles each.  --vtk writes the final state as a legacy VTK
// file with the array names of the reference's VTKWriter.C, binary by default (wf_vtk.hpp).  --out BASE
// reproduces the reference's output cadence (Solver_explicit.C:305, 1036-1041, 1155, 1167): after every step whose START
// time is >= tout (tout = 0, then += Configuration.outTime) the state is written to BASE_%05d.vtk and listed with that
// start time in BASE_res.json ({"vtk_files": [{"file", "time"}]}, ResultsJson, Solver_explicit.C:48-77); BASE_energy.csv
// gets one "t,Ekin,dEint" row per file.  The steps between two outputs run as one fused batch.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "wf_deck.hpp"
#include "wf_vtk.hpp"

using namespace wf_b200;

int main(int argc, char **argv) {
  std::string deck, dump, vtk, out_base;
  int steps = -1;
  bool parse_only = false, strict = false, vtk_ascii = false;
  double hexa_hg = 0.0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--steps") steps = atoi(val());
    else if (a == "--dump") dump = val();
    else if (a == "--vtk") vtk = val();
    else if (a == "--out") out_base = val();
    else if (a == "--vtk-ascii") vtk_ascii = true;
    else if (a == "--parse-only") parse_only = true;
    else if (a == "--strict") strict = true;
    else if (a == "--hexa-hg") hexa_hg = atof(val());
    else if (a[0] != '-') deck = a;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (deck.empty()) { fprintf(stderr, "usage: wf_weldform deck.json [--steps N] [--parse-only] [--dump FILE] [--vtk FILE] [--vtk-ascii] [--out BASE] [--strict] [--hexa-hg C]\n"); return 2; }
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
    if (!out_base.empty()) {
      if (!(S.out_time > 0.0)) throw std::runtime_error("--out needs Configuration.outTime > 0 in the deck");
      const double end_t = S.end_time, dt = S.dt;
      const long max_steps = steps >= 0 ? steps : -1;
      double tout = 0.0;
      int saved = 0;
      std::vector<std::pair<std::string, double>> files;
      FILE *csv = fopen((out_base + "_energy.csv").c_str(), "w");
      if (!csv) throw std::runtime_error("cannot open " + out_base + "_energy.csv");
      fprintf(csv, "t,Ekin,dEint\n");
      auto left = [&]() { return max_steps < 0 ? (long)1 << 40 : max_steps - dom.getStepCount(); };
      while (dom.getTime() < end_t && left() > 0) {
        // the steps whose start time is still below tout: one fused batch (same accumulation Time += dt as the engine)
        long n = 0;
        for (double t = dom.getTime(); t < tout && t < end_t && n < left(); t += dt) n++;
        if (n > 0) dom.Step((int)std::min<long>(n, 1000000));
        if (!(dom.getTime() < end_t) || left() <= 0) break;
        if (dom.getTime() < tout) continue;
        const double t_label = dom.getTime();
        dom.Step(1);
        char name[32];
        snprintf(name, sizeof name, "_%05d.vtk", saved);
        const std::string file = out_base + name;
        wfvtk::write_vtk(dom.handle(), dom.getDim(), dom.getNodxElem(), file, !vtk_ascii);
        files.emplace_back(file, t_label);
        FILE *jf = fopen((out_base + "_res.json").c_str(), "w");
        if (!jf) throw std::runtime_error("cannot open " + out_base + "_res.json");
        fprintf(jf, "{\n  \"vtk_files\": [\n");
        for (size_t i = 0; i < files.size(); i++)
          fprintf(jf, "    { \"file\": \"%s\", \"time\": %.17g }%s\n", files[i].first.c_str(), files[i].second, i + 1 < files.size() ? "," : "");
        fprintf(jf, "  ]\n}\n");
        fclose(jf);
        double ek_ = 0, de_ = 0;
        dom.computeEnergies(&ek_, &de_);
        fprintf(csv, "%.17g,%.17g,%.17g\n", t_label, ek_, de_);
        fflush(csv);
        saved++;
        tout += S.out_time;
        if (saved > 99999) throw std::runtime_error("--out: more than 99999 output files");
      }
      fclose(csv);
    } else if (steps >= 0) dom.Step(steps);
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
      if (!vtk.empty()) { // everything else the VTK writer reads, so that a test can rebuild the file from the dump
        for (const char *nm : {"vol_0", "rho", "p_node", "m_elem_area", "m_str_rate"})
          if (wf_array_bytes(dom.handle(), nm)) names.push_back(nm);
      }
      for (const char *nm : names) {
        std::vector<double> q = dom.get(nm);
        fprintf(f, "%s %zu\n", nm, q.size());
        fwrite(q.data(), sizeof(double), q.size(), f);
      }
      if (!vtk.empty() && wf_array_bytes(dom.handle(), "ext_nodes")) {
        std::vector<unsigned char> ext(wf_array_bytes(dom.handle(), "ext_nodes"));
        if (wf_get_array(dom.handle(), "ext_nodes", ext.data(), ext.size()) == 0) {
          std::vector<double> q(ext.begin(), ext.end());
          fprintf(f, "ext_nodes %zu\n", q.size());
          fwrite(q.data(), sizeof(double), q.size(), f);
        }
      }
      fclose(f);
    }
    if (!vtk.empty()) wfvtk::write_vtk(dom.handle(), dom.getDim(), dom.getNodxElem(), vtk, !vtk_ascii);
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
