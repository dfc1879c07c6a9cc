// wf_vtk.hpp — legacy-VTK output over the C ABI (a binary writer in place of the reference's ASCII one).
//
// The reference prints an ASCII legacy VTK file with 4-6 significant digits per value (src/common/VTKWriter.C:236-660).
// write_vtk() emits the same data set — UNSTRUCTURED_GRID, the same array names in the same order — as BINARY
// (big-endian float32 / int32) or ASCII.  Arrays (reference name <- engine array; an array the engine does not hold is
// skipped):
//   POINTS <- x; CELLS / CELL_TYPES <- m_elnod (12 hexahedron, 10 tetra, 9 quad, 5 triangle, VTKWriter.C:337-352)
//   POINT_DATA: VECTORS DISP <- u, Acceleration <- a, Velocity <- v; SCALARS Part_ID = 0; Temp <- T; VECTORS ContForce;
//     SCALARS nod_mass <- m_mdiag; stress = von Mises of the nodal average of m_sigma (avgScalar, Domain_d.h:77-87, then
//     sqrt(3 J2), VTKWriter.C:470-490); ext_nodes; nod_area <- node_area; nod_p <- p_node; TENSORS SIGMAT (upper triangle,
//     as the reference prints it, VTKWriter.C:524-530); TENSORS EPSR <- nodal average of m_str_rate (when stored)
//   CELL_DATA: ele_area <- m_elem_area; pressure <- p; pl_strain; TENSORS DDEVT <- m_str_rate; J = vol / vol_0; Vol; Rho;
//     Vol_0; sigy <- sigma_y
// Not written: the duplicate "Position" vector and the rigid tool surfaces the reference appends as extra cells.
// Python twin: weldformfem_b200/vtk.py — identical bytes (tests/test_vtk.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "wf_engine.h"

namespace wfvtk {

inline std::vector<double> get_d(wf_engine *e, const char *name, bool required) {
  const size_t nb = wf_array_bytes(e, name);
  std::vector<double> out(nb / sizeof(double));
  if (nb == 0) {
    if (required) throw std::runtime_error(std::string("wf_vtk: array not available: ") + name);
    return out;
  }
  if (wf_get_array(e, name, out.data(), nb)) {
    if (required) throw std::runtime_error(std::string("wf_vtk: ") + (wf_last_error(e) ? wf_last_error(e) : name));
    out.clear();
  }
  return out;
}
template <class T>
inline std::vector<T> get_i(wf_engine *e, const char *name) {
  const size_t nb = wf_array_bytes(e, name);
  std::vector<T> out(nb / sizeof(T));
  if (nb && wf_get_array(e, name, out.data(), nb)) throw std::runtime_error(std::string("wf_vtk: cannot read ") + name);
  return out;
}

struct Out {
  FILE *f;
  bool binary;
  void text(const std::string &s) { fputs(s.c_str(), f); fputc('\n', f); }
  void floats(const double *a, size_t n, int per_line) {
    if (binary) {
      std::vector<unsigned char> buf(4 * n);
      for (size_t i = 0; i < n; i++) {
        const float v = (float)a[i];
        uint32_t u;
        memcpy(&u, &v, 4);
        buf[4 * i] = (unsigned char)(u >> 24); buf[4 * i + 1] = (unsigned char)(u >> 16);
        buf[4 * i + 2] = (unsigned char)(u >> 8); buf[4 * i + 3] = (unsigned char)u;
      }
      fwrite(buf.data(), 1, buf.size(), f);
      fputc('\n', f);
    } else {
      for (size_t i = 0; i < n; i++) fprintf(f, "%.9g%c", (double)(float)a[i], ((int)(i % per_line) == per_line - 1) ? '\n' : ' ');
    }
  }
  void ints(const long long *a, size_t n, int per_line) {
    if (binary) {
      std::vector<unsigned char> buf(4 * n);
      for (size_t i = 0; i < n; i++) {
        const uint32_t u = (uint32_t)(int32_t)a[i];
        buf[4 * i] = (unsigned char)(u >> 24); buf[4 * i + 1] = (unsigned char)(u >> 16);
        buf[4 * i + 2] = (unsigned char)(u >> 8); buf[4 * i + 3] = (unsigned char)u;
      }
      fwrite(buf.data(), 1, buf.size(), f);
      fputc('\n', f);
    } else {
      for (size_t i = 0; i < n; i++) fprintf(f, "%lld%c", a[i], ((int)(i % per_line) == per_line - 1) ? '\n' : ' ');
    }
  }
};

// avgScalar (Domain_d.h:77-87): per node, the element rows of its list added in list order, divided by the count
inline std::vector<double> nodal_average(const std::vector<double> &ev, int nc, int nn, const std::vector<int> &nodel,
                                         const std::vector<int> &off, const std::vector<int> &cnt) {
  std::vector<double> out((size_t)nn * nc, 0.0);
  for (int n = 0; n < nn; n++) {
    for (int j = 0; j < cnt[n]; j++) {
      const size_t e = (size_t)nodel[off[n] + j];
      for (int c = 0; c < nc; c++) out[(size_t)n * nc + c] += ev[e * nc + c];
    }
    for (int c = 0; c < nc; c++) out[(size_t)n * nc + c] /= (double)cnt[n];
  }
  return out;
}

inline void write_vtk(wf_engine *e, int dim, int k, const std::string &path, bool binary = true) {
  int nn = 0, ne = 0;
  if (wf_get_counts(e, &nn, &ne, nullptr)) throw std::runtime_error("wf_vtk: no mesh");
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("wf_vtk: cannot open " + path);
  Out o{f, binary};
  try {
    auto pad3 = [&](const std::vector<double> &a) {
      std::vector<double> q((size_t)nn * 3, 0.0);
      for (int n = 0; n < nn; n++)
        for (int c = 0; c < dim; c++) q[(size_t)n * 3 + c] = a[(size_t)n * dim + c];
      return q;
    };
    auto scalars = [&](const char *name, const std::vector<double> &a) {
      o.text(std::string("SCALARS ") + name + " float 1");
      o.text("LOOKUP_TABLE default");
      o.floats(a.data(), a.size(), 1);
    };
    auto vectors = [&](const char *name, const std::vector<double> &a) {
      o.text(std::string("VECTORS ") + name + " float");
      const std::vector<double> q = pad3(a);
      o.floats(q.data(), q.size(), 3);
    };
    auto tensors = [&](const char *name, const std::vector<double> &t6) { // xx yy zz xy yz xz -> upper triangle
      o.text(std::string("TENSORS ") + name + " float");
      const size_t n = t6.size() / 6;
      std::vector<double> q(9 * n, 0.0);
      for (size_t i = 0; i < n; i++) {
        const double *t = &t6[6 * i];
        double *r = &q[9 * i];
        r[0] = t[0]; r[1] = t[3]; r[2] = t[5]; r[4] = t[1]; r[5] = t[4]; r[8] = t[2];
      }
      o.floats(q.data(), q.size(), 9);
    };
    o.text("# vtk DataFile Version 3.0");
    o.text("WeldFormFEM explicit step, B200 engine");
    o.text(binary ? "BINARY" : "ASCII");
    o.text("DATASET UNSTRUCTURED_GRID");
    o.text("POINTS " + std::to_string(nn) + " float");
    {
      const std::vector<double> q = pad3(get_d(e, "x", true));
      o.floats(q.data(), q.size(), 3);
    }
    const std::vector<unsigned> el = get_i<unsigned>(e, "m_elnod");
    {
      std::vector<long long> c((size_t)ne * (k + 1));
      for (int i = 0; i < ne; i++) {
        c[(size_t)i * (k + 1)] = k;
        for (int a = 0; a < k; a++) c[(size_t)i * (k + 1) + 1 + a] = el[(size_t)i * k + a];
      }
      o.text("CELLS " + std::to_string(ne) + " " + std::to_string((long long)ne * (k + 1)));
      o.ints(c.data(), c.size(), k + 1);
      const int type = dim == 3 ? (k == 8 ? 12 : 10) : (k == 4 ? 9 : 5);
      std::vector<long long> t((size_t)ne, type);
      o.text("CELL_TYPES " + std::to_string(ne));
      o.ints(t.data(), t.size(), 1);
    }
    o.text("POINT_DATA " + std::to_string(nn));
    vectors("DISP", get_d(e, "u", true));
    vectors("Acceleration", get_d(e, "a", true));
    vectors("Velocity", get_d(e, "v", true));
    scalars("Part_ID", std::vector<double>((size_t)nn, 0.0));
    {
      const std::vector<double> T = get_d(e, "T", false);
      if (!T.empty()) scalars("Temp", T);
      const std::vector<double> cf = get_d(e, "contforce", false);
      if (!cf.empty()) vectors("ContForce", cf);
    }
    scalars("nod_mass", get_d(e, "m_mdiag", true));
    const std::vector<int> nodel = get_i<int>(e, "m_nodel"), off = get_i<int>(e, "m_nodel_offset"), cnt = get_i<int>(e, "m_nodel_count");
    const std::vector<double> sig = get_d(e, "m_sigma", false);
    std::vector<double> sig_n;
    if (!sig.empty()) {
      sig_n = nodal_average(sig, 6, nn, nodel, off, cnt);
      std::vector<double> vm((size_t)nn);
      for (int n = 0; n < nn; n++) {
        const double *s = &sig_n[(size_t)n * 6];
        const double tr3 = (s[0] + s[1] + s[2]) * (1.0 / 3.0);
        const double s0 = s[0] - tr3, s1 = s[1] - tr3, s2 = s[2] - tr3;
        const double j2 = 0.5 * (s0 * s0 + 2.0 * (s[3] * s[3]) + 2.0 * (s[5] * s[5]) + s1 * s1 + 2.0 * (s[4] * s[4]) + s2 * s2);
        vm[n] = sqrt(3.0 * j2);
      }
      scalars("stress", vm);
    }
    {
      const std::vector<unsigned char> ext = get_i<unsigned char>(e, "ext_nodes");
      if (!ext.empty()) {
        std::vector<double> q((size_t)nn);
        for (int n = 0; n < nn; n++) q[n] = ext[n] ? 1.0 : 0.0;
        scalars("ext_nodes", q);
      }
      const std::vector<double> na = get_d(e, "node_area", false);
      if (!na.empty()) scalars("nod_area", na);
      const std::vector<double> pn = get_d(e, "p_node", false);
      if (!pn.empty()) scalars("nod_p", pn);
    }
    if (!sig_n.empty()) tensors("SIGMAT", sig_n);
    const std::vector<double> sr = get_d(e, "m_str_rate", false);
    if (!sr.empty()) tensors("EPSR", nodal_average(sr, 6, nn, nodel, off, cnt));
    o.text("CELL_DATA " + std::to_string(ne));
    {
      const std::vector<double> ea = get_d(e, "m_elem_area", false);
      if (!ea.empty()) scalars("ele_area", ea);
    }
    scalars("pressure", get_d(e, "p", true));
    scalars("pl_strain", get_d(e, "pl_strain", true));
    if (!sr.empty()) tensors("DDEVT", sr);
    const std::vector<double> vol = get_d(e, "vol", true), vol0 = get_d(e, "vol_0", true);
    {
      std::vector<double> J((size_t)ne);
      for (int i = 0; i < ne; i++) J[i] = vol[i] / vol0[i];
      scalars("J", J);
    }
    scalars("Vol", vol);
    scalars("Rho", get_d(e, "rho", true));
    scalars("Vol_0", vol0);
    scalars("sigy", get_d(e, "sigma_y", true));
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
}

} // namespace wfvtk
