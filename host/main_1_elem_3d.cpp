// main_1_elem_3d.cpp — configs[0] of BASELINE.json driven from C++ through the C ABI: the same program as the
// reference's src/common/main_1_elem_3d.C:50-192 (0.1 m steel cube, ONE reduced-integration hexa, top face
// v_z = -1, dt = 0.8e-5, while (Time < 1e-3)), with `Domain_d` coming from wf_domain.hpp instead of the CPU class.
// Prints the displacement / velocity / internal-force block the reference prints after the loop
// (validation/1elem_3d_red_int_f_0.06.txt).   usage: main_1_elem_3d [hg_coeff=0.06] [strict=0]
#include <cstdio>
#include <cstdlib>

#include "wf_domain.hpp"

using namespace wf_b200;

int main(int argc, char **argv) {
  double hg = argc > 1 ? atof(argv[1]) : 0.06;
  bool strict = argc > 2 && atoi(argv[2]) != 0;
  try {
    Domain_d *dom_d = new Domain_d(/*device=*/0);
    double3 V = make_double3(0.0, 0.0, 0.0);
    double dx = 0.1;
    double3 L = make_double3(dx, dx, dx);
    double r = 0.05;
    dom_d->AddBoxLength(V, L, r, true);

    double E = 206.0e9, nu = 0.3, rho = 7850.0;
    dom_d->setDensity(rho);
    Elastic_ el(E, nu);
    Material_ *material_h = new Material_(el);
    material_h->cs0 = sqrt(material_h->Elastic().BulkMod() / rho);
    material_h->Material_model = BILINEAR;
    dom_d->AssignMaterial(material_h);
    dom_d->setHexaHourglass(hg);
    dom_d->setStrict(strict);

    double dt = 0.800e-5;
    dom_d->SetDT(dt);
    dom_d->SetEndTime(1.0e-3);

    dom_d->AddBCVelNode(0, 0, 0); dom_d->AddBCVelNode(0, 1, 0); dom_d->AddBCVelNode(0, 2, 0);
    dom_d->AddBCVelNode(1, 1, 0); dom_d->AddBCVelNode(1, 2, 0);
    dom_d->AddBCVelNode(2, 0, 0); dom_d->AddBCVelNode(2, 2, 0);
    dom_d->AddBCVelNode(3, 2, 0);
    for (int i = 0; i < 4; i++) dom_d->AddBCVelNode(i + 4, 2, -1.0);
    dom_d->AllocateBCs();

    printf("Element Count %d\n", dom_d->getElemCount());
    dom_d->SolveChungHulbert();
    printf("steps %ld time %.17g\n", dom_d->getStepCount(), dom_d->getTime());

    const char *names[] = {"u", "v", "a", "m_fi"};
    for (const char *nm : names) {
      std::vector<double> q = dom_d->get(nm);
      printf("%s\n", nm);
      for (int n = 0; n < dom_d->getNodeCount(); n++) printf("%.17g %.17g %.17g\n", q[3 * n], q[3 * n + 1], q[3 * n + 2]);
    }
    std::vector<double> s = dom_d->get("m_sigma");
    printf("m_sigma\n%.17g %.17g %.17g %.17g %.17g %.17g\n", s[0], s[1], s[2], s[3], s[4], s[5]);
    delete material_h;
    delete dom_d;
  } catch (const std::exception &e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  printf("Program ended.\n");
  return 0;
}
