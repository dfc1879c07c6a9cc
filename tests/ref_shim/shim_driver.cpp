// TEST INFRASTRUCTURE.  The reference's own front-end (src/explicit/main.C, unmodified) with the ONE call swapped that
// INTEGRATION.md describes: `dom_d->SolveChungHulbert()` (main.C:995) runs the B200 engine through
// tests/ref_shim/Solver_b200.C — or, with WF_SHIM_CPU=1, the reference's own CPU solver — and the final state is
// written to $WF_SHIM_DUMP so the two runs can be compared (tests/test_ref_shim.py).
//
// Built by `make -C oracle shim` into oracle/_ref/wf_ref_shim from the reference sources where they lie.
// The swap is done by name hiding, the same way oracle/ref_harness.cpp captures main.C's domain: main.C's `Domain_d`
// is spelled ShimDomain, whose SolveChungHulbert() is the dispatch.  On top of the shim only test plumbing is added:
// zero-filling of the malloc'ed state the reference never initialises (Domain_d.C:457-621), a fixed
// time step for the CPU arm, and the dump.
#include <cstring>
#include <string>

#include "Solver_b200.C"

#include "Mesh.h"

namespace MetFEM {

class ShimDomain : public Domain_b200 {
  void zero_state() {
    const size_t nd = (size_t)m_node_count * m_dim, ne = (size_t)m_elem_count, nk = ne * m_nodxelem * m_dim;
    auto z = [](double *q, size_t n) { if (q) memset(q, 0, n * sizeof(double)); };
    z(prev_a, nd); z(m_fe, nd); z(m_fi, nd); z(a, nd); z(v, nd); z(u, nd); z(u_dt, nd); z(contforce, nd); z(ut_prev, nd);
    z(m_tau, 6 * ne); z(m_sigma, 6 * ne); z(m_eps, 6 * ne); z(m_str_rate, 6 * ne); z(m_rot_rate, 6 * ne);
    z(m_strain_pl_incr, 6 * ne);
    z(p, ne); z(pl_strain, ne); z(sigma_y, ne); z(m_radius, ne); z(rho, ne); z(rho_0, ne); z(vol, ne); z(vol_0, ne);
    z(m_detJ, ne); z(m_f_elem, nk); z(m_f_elem_hg, nk);
    z(m_mdiag, m_node_count); z(m_voln, m_node_count); z(p_node, m_node_count);
    z(m_dTedt, ne * m_nodxelem); z(m_q_plheat, ne); z(T, m_node_count); z(node_area, m_node_count);
    z(q_cont_conv, m_node_count); z(m_elem_area, ne); z(m_elem_length, ne);
    if (m_dim == 2 && m_hg_q) z(m_hg_q, nk);
  }

 public:
  void AddBoxLength(double3 const &V, double3 const &L, const double &r, const bool &red_int = true, const bool &tritetra = false) {
    Domain_d::AddBoxLength(V, L, r, red_int, tritetra);
    zero_state();
  }
  void CreateFromLSDyna(LS_Dyna::lsdynaReader &reader) {
    Domain_d::CreateFromLSDyna(reader);
    zero_state();
  }

  // what main.C:995 calls
  void SolveChungHulbert() {
    if (getenv("WF_SHIM_CPU")) {
      setFixedDt(true);
      Domain_d::SolveChungHulbert();                 // the reference's own CPU solver
    } else {
      if (const char *h = getenv("WF_SHIM_HEXA_HG")) hexa_hg_coeff = atof(h);
      if (getenv("WF_SHIM_STRICT")) strict_reference = 1;
      AttachB200(0);
      SolveChungHulbert_b200();
    }
    if (const char *path = getenv("WF_SHIM_DUMP")) dump(path);
    // main.C's main() has no return statement: renamed to an ordinary function, falling off its end is undefined
    // behaviour (g++ -O2 runs on into the next function).  The run is complete here, so leave from here.
    fflush(stdout);
    exit(0);
  }

  void dump(const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); exit(1); }
    const size_t nd = (size_t)m_node_count * m_dim, ne = (size_t)m_elem_count;
    struct { const char *nm; const double *p; size_t n; } arr[] = {
        {"x", x, nd}, {"v", v, nd}, {"u", u, nd}, {"a", a, nd}, {"m_sigma", m_sigma, ne * 6}, {"m_tau", m_tau, ne * 6},
        {"pl_strain", pl_strain, ne}, {"p", p, ne}, {"sigma_y", sigma_y, ne}, {"vol", vol, ne}};
    for (auto &q : arr) {
      fprintf(f, "%s %zu\n", q.nm, q.n);
      fwrite(q.p, sizeof(double), q.n, f);
    }
    double t[2] = {Time, dt};
    fprintf(f, "time_dt 2\n");
    fwrite(t, sizeof(double), 2, f);
    fclose(f);
  }
};

}  // namespace MetFEM

#include "src/common/NastranReader.cpp"  // NastranReader::read is `inline` (src/common/NastranReader.cpp:34), main.C:690 calls it
#define Domain_d ShimDomain
#define main wf_shim_main
#include "src/explicit/main.C"
#undef main
#undef Domain_d

int main(int argc, char **argv) { return wf_shim_main(argc, argv); }
