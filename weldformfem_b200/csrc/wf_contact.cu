// wf_contact.cu — penalty contact of the deformable body's external nodes with rigid tool surfaces
// (SURVEY.md §8f-2), device kernels + their host side.
//
// Reference: Domain_d::SearchExtNodes / CalcExtFaceAreas (src/common/Domain_d.C:110-315),
// Domain_d::CalcContactForces (src/common/Contact.C:31-336), TriMesh_d::Move / CalcNormals /
// UpdatePlaneCoeff / CalcSpheres (include/common/Mesh.h:228-328) and their place in the solver loop
// (src/explicit/Solver_explicit.C:168-173, 445-450, 769-770, 981-1005).
//
// Device design: the external nodes (a surface, O(N^(2/3)) of the mesh) are kept as a compact ascending list;
// one thread per external node walks the rigid facets in the reference's order (the first facet the node lies
// behind AND projects into wins, Contact.C:311), so the search is deterministic and needs no atomics.  Nodal
// areas are a gather over the node's external faces in faceList order.  The contact force persists between
// steps in `contforce` exactly like the reference's array does; the nodal update (N2) adds it to the
// acceleration and the element pass (E2) reads the per-node "has contact force" flag for calcElemPressure.
// This file is compiled with -fmad=false: the kernels are tiny and bit-level agreement with the CPU path
// (libm sqrt is correctly rounded on both sides) is worth more than the FMAs.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include "wf_engine_priv.h"

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      E->err = std::string(#call) + ": " + cudaGetErrorString(_e);                        \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)
#define FAIL(msg) do { E->err = (msg); return 1; } while (0)
#define NEED(cond, msg) do { if (!(cond)) FAIL(msg); } while (0)

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 vdiv(V3 a, double s) { double inv = 1.0 / s; return a * inv; } // double3_c.h:95-99
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double len(V3 a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 ld3(const double *p, int i) { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(double *p, int i, V3 a) { p[3 * i] = a.x; p[3 * i + 1] = a.y; p[3 * i + 2] = a.z; }
// getPosVec3 / getVelVec / getAccVec (Domain_d.h:447-481): z = 0 in 2D
__device__ __forceinline__ V3 node3(const WfDev &d, const double *q, int n) {
  return mk(q[n], q[d.np + n], d.dim == 3 ? q[2 * d.np + n] : 0.0);
}

// ---- CalcExtFaceAreas (Domain_d.C:210-315) ----------------------------------------------------------
__global__ void k_xf_area(WfDev d, WfContact c) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= c.n_xf) return;
  const int n0 = c.xf_nodes[f], n1 = c.xf_nodes[c.n_xf + f];
  if (d.dim == 2) {
    double dx = d.x[n1] - d.x[n0], dy = d.x[d.np + n1] - d.x[d.np + n0];
    c.xf_area[f] = sqrt(dx * dx + dy * dy);
  } else {
    const int n2 = c.xf_nodes[2 * c.n_xf + f];
    V3 p0 = node3(d, d.x, n0), p1 = node3(d, d.x, n1), p2 = node3(d, d.x, n2);
    V3 cr = cross(p1 - p0, p2 - p0);
    c.xf_area[f] = 0.5 * sqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
  }
}
__global__ void k_xn_area(WfDev d, WfContact c) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c.n_ext) return;
  double s = 0.0;
  for (int q = c.xn_ptr[t]; q < c.xn_ptr[t + 1]; q++) {
    const double a = c.xf_area[c.xn_faces[q]];
    s += (d.dim == 2) ? 0.5 * a : a / 3.0;
  }
  c.node_area[c.ext_nodes[t]] = s;
}
// m_elem_area: 3D = largest external face of the element, 2D = sum of its external edges
__global__ void k_xe_area(WfDev d, WfContact c) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= c.n_xf) return;
  const int e = c.xf_elem[f];
  if (f > 0 && c.xf_elem[f - 1] == e) return;
  double a = (d.dim == 2) ? 0.0 : c.xf_area[f];
  if (d.dim == 2) {
    for (int g = f; g < c.n_xf && c.xf_elem[g] == e; g++) a += c.xf_area[g];
  } else {
    for (int g = f + 1; g < c.n_xf && c.xf_elem[g] == e; g++)
      if (c.xf_area[g] > a) a = c.xf_area[g];
  }
  c.elem_area[e] = a;
}

// mean m_elem_length of the elements around every external node (Contact.C:146-153), nodel order
__global__ void k_nodlen(WfDev d, WfContact c, const double *__restrict__ elem_length) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c.n_ext) return;
  const int n = c.ext_nodes[t];
  const long long base = d.sell_ptr[n >> 5];
  const int width = (int)((d.sell_ptr[(n >> 5) + 1] - base) >> 5);
  double s = 0.0;
  for (int j = 0; j < width; j++) {
    int slot = d.sell_slots[base + ((long long)j << 5) + (n & 31)];
    if (slot >= 0) s += elem_length[slot / d.k];
  }
  c.nodlen[t] = s / d.nodel_count[n];
}

// ---- rigid surfaces: ramp + Move + CalcNormals + UpdatePlaneCoeff (Solver_explicit.C:981-1005, Mesh.h:241-297) ----
__global__ void k_trimesh_update(WfContact c, double f, double dt, int move) {
  if (move) {
    for (int n = threadIdx.x; n < c.tm_nn; n += blockDim.x) {
      V3 v = ld3(c.tm_v_orig, n) * f;
      st3(c.tm_node_v, n, v);
      st3(c.tm_node, n, ld3(c.tm_node, n) + v * dt);
    }
    __syncthreads();
  }
  const int nen = c.tm_dim == 3 ? 3 : 2;
  for (int e = threadIdx.x; e < c.tm_ne; e += blockDim.x) {
    V3 nrm;
    if (move) {
      if (c.tm_dim == 3) {
        V3 u = ld3(c.tm_node, c.tm_elnode[3 * e + 1]) - ld3(c.tm_node, c.tm_elnode[3 * e]);
        V3 v = ld3(c.tm_node, c.tm_elnode[3 * e + 2]) - ld3(c.tm_node, c.tm_elnode[3 * e]);
        V3 w = cross(u, v);
        nrm = vdiv(w, len(w));
      } else {
        V3 u = ld3(c.tm_node, c.tm_elnode[2 * e + 1]) - ld3(c.tm_node, c.tm_elnode[2 * e]);
        V3 v = mk(-u.y, u.x, 0.0);
        nrm = vdiv(v, len(v));
      }
      st3(c.tm_normal, e, nrm);
    } else {
      nrm = ld3(c.tm_normal, e);
    }
    // nfar ends as the element's last node whatever the distances (CalcSpheres, Mesh.h:308-316); the reference
    // indexes elnode with `dimension*e`
    c.tm_pplane[e] = dot(ld3(c.tm_node, c.tm_elnode[nen * e + (c.tm_dim - 1)]), nrm);
  }
}

// ---- CalcContactForces (Contact.C:31-336) ---------------------------------------------------------------
struct Hit {
  V3 nj, cf, v_tan, du;
  double kcont, normFn;
};
__device__ __forceinline__ bool slides(const WfContact &c, const Hit &h, double utx, double uty, double utz) {
  V3 ut_acc = mk(utx, uty, utz) + h.du;
  V3 Ft_trial = ut_acc * (-h.kcont);
  return !(len(Ft_trial) <= c.mu_sta * h.normFn);
}
// friction from the accumulated slip (Contact.C:236-304); returns the friction force, updates (utx, uty, utz)
__device__ __forceinline__ V3 friction(const WfContact &c, const Hit &h, double &utx, double &uty, double &utz, bool &slid) {
  V3 ut_acc = mk(utx, uty, utz) + h.du;
  utx += h.du.x; uty += h.du.y; utz += h.du.z;
  V3 Ft_trial = ut_acc * (-h.kcont);
  const double Ft_mag = len(Ft_trial);
  const double Ft_max_static = c.mu_sta * h.normFn;
  slid = !(Ft_mag <= Ft_max_static);
  if (!slid) return Ft_trial;
  const double Ft_max_dynamic = c.mu_dyn * h.normFn;
  const double nvt = sqrt(h.v_tan.x * h.v_tan.x + h.v_tan.y * h.v_tan.y + h.v_tan.z * h.v_tan.z);
  utx = uty = utz = 0.0;
  return vdiv(h.v_tan * (-Ft_max_dynamic), nvt);
}

// facet test of Contact.C:68-139: the node lies behind facet j's plane and its projection falls inside the facet
__device__ __forceinline__ bool facet_hit(const WfContact &c, const V3 &xi, int j, V3 &nj, double &dist) {
  const int nen = c.tm_dim == 3 ? 3 : 2;
  nj = ld3(c.tm_normal, j);
  dist = dot(nj, xi) - c.tm_pplane[j];
  if (!(dist < 0)) return false;
  const V3 Qj = xi - nj * dist;
  if (c.tm_dim == 3) {
    for (int l = 0; l < 3; l++) {
      const int n = (l + 1 > 2) ? 0 : l + 1;
      const V3 nl = ld3(c.tm_node, c.tm_elnode[nen * j + l]);
      const double crit = dot(cross(ld3(c.tm_node, c.tm_elnode[nen * j + n]) - nl, Qj - nl), nj);
      if (crit < 0.0) return false;
    }
  } else {
    for (int l = 0; l < 2; l++) {
      const int n = (l + 1 > 1) ? 0 : l + 1;
      const V3 nl = ld3(c.tm_node, c.tm_elnode[nen * j + l]);
      const double crit = dot(ld3(c.tm_node, c.tm_elnode[nen * j + n]) - nl, Qj - nl);
      if (crit < 0.0) return false;
    }
  }
  return true;
}

struct Hit;
__device__ void contact_force(const WfDev &d, const WfContact &c, double dt, int t, int i, const V3 &xi, int jhit);

// Contact search in two kernels.
// k_contact_filter — one THREAD per external node walks the facets in ascending order against a single-precision copy
// of the plane coefficients staged per CTA in shared memory as float4 (nx, ny, nz, pplane); all threads of a warp read
// the SAME facet, so every load is one broadcast.  The fp32 distance is only a conservative filter: a facet is ruled
// out when d32 exceeds the rounding bound 1e-6 * ((|nx|+|ny|+|nz|) * max|x_c| + |pplane|) (inputs rounded to fp32,
// relative error 2^-24 each, four products).  Nodes in front of every plane are finished here (no contact); the others
// are appended, with the first facet that could not be ruled out, to a candidate list.
// k_contact_exact — one WARP per candidate: the 32 lanes run the reference's double-precision test (Contact.C:68-139)
// on 32 consecutive facets at a time, ascending; the lowest hit lane of the first chunk with a hit is the reference's
// "first facet that contains the projection" (Contact.C:311: one master facet per slave node).  Candidates are the
// nodes at or behind a tool surface — they sit next to each other in the node numbering, so giving each its own warp
// is what spreads the fp64 work over the whole GPU.
constexpr int CONTACT_TPB = 64;
constexpr int CONTACT_CHUNK = 2048; // facets per shared-memory chunk (32 KB)
__global__ void __launch_bounds__(CONTACT_TPB) k_contact_filter(WfDev d, WfContact c, double dt, int *__restrict__ cand_count,
                                                                int2 *__restrict__ cand) {
  __shared__ float4 pl[CONTACT_CHUNK];
  const int t = blockIdx.x * CONTACT_TPB + threadIdx.x;
  const bool valid = t < c.n_ext;
  const int i = valid ? c.ext_nodes[t] : 0;
  const V3 xi = valid ? node3(d, d.x, i) : mk(0.0, 0.0, 0.0);
  const float xf = (float)xi.x, yf = (float)xi.y, zf = (float)xi.z;
  const float xinf = fmaxf(fabsf(xf), fmaxf(fabsf(yf), fabsf(zf)));
  int first = -1;
  for (int base = 0; base < c.tm_ne; base += CONTACT_CHUNK) {
    const int cnt = min(CONTACT_CHUNK, c.tm_ne - base);
    __syncthreads();
    for (int j = threadIdx.x; j < ((cnt + 7) & ~7); j += CONTACT_TPB)
      pl[j] = j < cnt ? make_float4((float)c.tm_normal[3 * (base + j)], (float)c.tm_normal[3 * (base + j) + 1],
                                    (float)c.tm_normal[3 * (base + j) + 2], (float)c.tm_pplane[base + j])
                      : make_float4(0.f, 0.f, 0.f, -1.0e30f); // padding: infinitely far in front
    __syncthreads();
    if (valid && first < 0) {
      // eight facets per trip, flags first and one branch after: the loads and FMAs of a trip overlap
      for (int jj = 0; jj < cnt && first < 0; jj += 8) {
        unsigned m = 0;
#pragma unroll
        for (int q8 = 0; q8 < 8; q8++) {
          const float4 q = pl[jj + q8];
          const float d32 = fmaf(q.x, xf, fmaf(q.y, yf, q.z * zf)) - q.w;
          const float bound = 1.0e-6f * fmaf(fabsf(q.x) + fabsf(q.y) + fabsf(q.z), xinf, fabsf(q.w));
          if (d32 <= bound) m |= 1u << q8; // cannot be ruled out in single precision
        }
        if (m) first = base + jj + (__ffs(m) - 1);
      }
    }
  }
  if (!valid) return;
  if (first < 0) contact_force(d, c, dt, t, i, xi, -1);
  else cand[atomicAdd(cand_count, 1)] = make_int2(t, first);
}

// single-precision classification of (node, facet): false = the reference's test certainly fails (in front of the
// plane, or the projection certainly outside one edge), true = possible hit, to be decided in double precision.
// Rounding bounds: coordinates carry 2^-24 relative error, the fp32 distance at most `bound`; every edge criterion is
// a sum of products of two coordinate differences (times a normal component), hence the margin 1e-5 * s * (|e| + |r|)
// with s the largest coordinate magnitude involved — several times the worst case.
template <bool D3>
__device__ __forceinline__ bool facet_possible32(const WfContact &c, float xf, float yf, float zf, float xinf, int j) {
  constexpr int NEN = D3 ? 3 : 2;
  // branch-free: all loads are issued up front so that several facets per lane overlap their memory round trips
  int g[NEN];
#pragma unroll
  for (int l = 0; l < NEN; l++) g[l] = c.tm_elnode[NEN * j + l];
  const float nx = (float)c.tm_normal[3 * j], ny = (float)c.tm_normal[3 * j + 1], nz = (float)c.tm_normal[3 * j + 2];
  const float pp = (float)c.tm_pplane[j];
  float vx[NEN], vy[NEN], vz[NEN];
#pragma unroll
  for (int l = 0; l < NEN; l++) {
    vx[l] = (float)c.tm_node[3 * g[l]]; vy[l] = (float)c.tm_node[3 * g[l] + 1]; vz[l] = (float)c.tm_node[3 * g[l] + 2];
  }
  const float nsum = fabsf(nx) + fabsf(ny) + fabsf(nz);
  const float d32 = fmaf(nx, xf, fmaf(ny, yf, nz * zf)) - pp;
  const float bound = 1.0e-6f * fmaf(nsum, xinf, fabsf(pp));
  bool ok = !(d32 > bound);
  const float qx = xf - nx * d32, qy = yf - ny * d32, qz = zf - nz * d32;
  float s = xinf + fabsf(d32);
#pragma unroll
  for (int l = 0; l < NEN; l++) s = fmaxf(s, fmaxf(fabsf(vx[l]), fmaxf(fabsf(vy[l]), fabsf(vz[l]))));
#pragma unroll
  for (int l = 0; l < NEN; l++) {
    const int n = (l + 1 == NEN) ? 0 : l + 1;
    const float ex = vx[n] - vx[l], ey = vy[n] - vy[l], ez = vz[n] - vz[l];
    const float rx = qx - vx[l], ry = qy - vy[l], rz = qz - vz[l];
    const float einf = fmaxf(fabsf(ex), fmaxf(fabsf(ey), fabsf(ez))), rinf = fmaxf(fabsf(rx), fmaxf(fabsf(ry), fabsf(rz)));
    float crit, margin = 1.0e-5f * s * (einf + rinf);
    if (D3) {
      crit = (ey * rz - ez * ry) * nx + (ez * rx - ex * rz) * ny + (ex * ry - ey * rx) * nz;
      margin *= fmaxf(nsum, 1.0f);
    } else {
      crit = ex * rx + ey * ry + ez * rz;
    }
    ok = ok && !(crit < -margin);
  }
  return ok;
}

template <bool D3>
__global__ void __launch_bounds__(256) k_contact_exact(WfDev d, WfContact c, double dt, const int *__restrict__ cand_count,
                                                       const int2 *__restrict__ cand) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= *cand_count) return;
  const int2 cd = cand[w];
  const int t = cd.x, i = c.ext_nodes[t];
  const V3 xi = node3(d, d.x, i);
  const float xf = (float)xi.x, yf = (float)xi.y, zf = (float)xi.z;
  const float xinf = fmaxf(fabsf(xf), fmaxf(fabsf(yf), fabsf(zf)));
  int jhit = -1;
  for (int j0 = cd.y & ~31; j0 < c.tm_ne && jhit < 0; j0 += 128) { // four facets per lane and trip
    bool poss[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int j = min(j0 + 32 * q + lane, c.tm_ne - 1);
      poss[q] = facet_possible32<D3>(c, xf, yf, zf, xinf, j) && (j0 + 32 * q + lane < c.tm_ne);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      V3 nj;
      double dist;
      const bool mine = poss[q] && facet_hit(c, xi, j0 + 32 * q + lane, nj, dist);
      const unsigned any = __ballot_sync(0xffffffffu, mine);
      if (any && jhit < 0) jhit = j0 + 32 * q + (__ffs(any) - 1);
    }
  }
  if (lane == 0) contact_force(d, c, dt, t, i, xi, jhit);
}

// force on external node i (list index t) from facet jhit (< 0: no contact), Contact.C:141-309
__device__ void contact_force(const WfDev &d, const WfContact &c, double dt, int t, int i, const V3 &xi, int jhit) {
  bool hit = false;
  Hit h;
  int mesh = -1;
  if (jhit >= 0) {
    const int j = jhit;
    V3 nj;
    double dist;
    facet_hit(c, xi, j, nj, dist);
    hit = true;
    mesh = c.tm_mesh_id[j];
    const V3 v_rel = node3(d, d.v, i);
    const double mass = d.mdiag[i];
    const double v_reln = dot(v_rel, nj);
    const double kcont_geo = c.young * c.node_area[i] / c.nodlen[t];
    const double kcont_mass = 0.2 * mass / (dt * dt);
    const double kcont = kcont_mass < kcont_geo ? kcont_mass : kcont_geo; // std::min(geo, mass)
    const double omega = sqrt(kcont / mass);
    const double ccrit = 2.0 * mass * omega;
    const double F_damp = 0.2 * ccrit * v_reln;
    const double F_normal = c.contPF * kcont * dist;
    h.nj = nj;
    h.cf = nj * (-(F_normal + F_damp));
    h.kcont = kcont;
    h.v_tan = v_rel - nj * dot(v_rel, nj);
    h.du = h.v_tan * dt;
    if (d.dim == 2) h.du.z = 0.0;
    const V3 Fn = nj * dot(h.cf, nj);
    h.normFn = sqrt(Fn.x * Fn.x + Fn.y * Fn.y + Fn.z * Fn.z);
  }
  c.mesh_in_contact[i] = mesh;
  if (d.dim == 3) {
    if (!hit) return;
    double ux = c.ut_prev[i], uy = c.ut_prev[d.np + i], uz = c.ut_prev[2 * d.np + i];
    bool slid;
    const V3 Ft = friction(c, h, ux, uy, uz, slid);
    c.ut_prev[i] = ux; c.ut_prev[d.np + i] = uy; c.ut_prev[2 * d.np + i] = uz;
    const V3 cf = mk(h.cf.x + Ft.x, h.cf.y + Ft.y, h.cf.z + Ft.z);
    d.contforce[i] = cf.x; d.contforce[d.np + i] = cf.y; d.contforce[2 * d.np + i] = cf.z;
    d.cflag[i] = (dot(cf, cf) > 0) ? 1 : 0;
    if (d.q_cont_conv) d.q_cont_conv[i] = c.heat_cond * c.node_area[i] * (c.T_const - d.T[i]); // Contact.C:309
  } else {
    // 2D: the reference's friction code touches ut_prev[2*i+2] — the x slot of node i+1 (Contact.C:258, 297) — so
    // node i+1 sees a zeroed slip if node i slid in the same pass.  Record what the serial pass needs.
    double *r = c.rec + t;
    const long long S = c.n_ext;
    r[0] = hit ? 1.0 : 0.0;
    if (!hit) return;
    const double ux = c.ut_prev[i], uy = c.ut_prev[d.np + i];
    r[1 * S] = slides(c, h, ux, uy, 0.0) ? 1.0 : 0.0;  // outcome with the stored slip
    r[2 * S] = slides(c, h, 0.0, uy, 0.0) ? 1.0 : 0.0; // outcome if node i-1 zeroed this node's x slot
    r[3 * S] = h.kcont; r[4 * S] = h.normFn;
    r[5 * S] = h.du.x; r[6 * S] = h.du.y;
    r[7 * S] = h.v_tan.x; r[8 * S] = h.v_tan.y;
    d.contforce[i] = h.cf.x; d.contforce[d.np + i] = h.cf.y;
    if (d.q_cont_conv) d.q_cont_conv[i] = c.heat_cond * c.node_area[i] * (c.T_const - d.T[i]); // Contact.C:309
  }
}

// serial propagation of "my x slot was zeroed by the previous node" along the ascending node order: only booleans
// travel, staged through shared memory; r[9] <- zeroed flag
__global__ void __launch_bounds__(1024) k_friction2d_chain(WfContact c) {
  __shared__ unsigned char s_hit[1024], s_s[1024], s_z[1024], s_adj[1024], s_out[1024];
  __shared__ int carry_slid;
  const long long S = c.n_ext;
  if (threadIdx.x == 0) carry_slid = 0;
  for (int base = 0; base < c.n_ext; base += 1024) {
    const int t = base + threadIdx.x;
    if (t < c.n_ext) {
      const bool hit = c.rec[t] != 0.0;
      s_hit[threadIdx.x] = hit;
      s_s[threadIdx.x] = hit && c.rec[1 * S + t] != 0.0;
      s_z[threadIdx.x] = hit && c.rec[2 * S + t] != 0.0;
      s_adj[threadIdx.x] = (t > 0) && (c.ext_nodes[t - 1] + 1 == c.ext_nodes[t]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int prev_slid = carry_slid;
      const int cnt = min(1024, c.n_ext - base);
      for (int q = 0; q < cnt; q++) {
        const bool zeroed = prev_slid && s_adj[q];
        s_out[q] = zeroed;
        prev_slid = s_hit[q] ? (zeroed ? s_z[q] : s_s[q]) : 0;
      }
      carry_slid = prev_slid;
    }
    __syncthreads();
    if (t < c.n_ext) c.rec[9 * S + t] = s_out[threadIdx.x] ? 1.0 : 0.0;
    __syncthreads();
  }
}

__global__ void k_friction2d_apply(WfDev d, WfContact c) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c.n_ext) return;
  const long long S = c.n_ext;
  const double *r = c.rec + t;
  if (r[0] == 0.0) return;
  const int i = c.ext_nodes[t];
  Hit h;
  h.kcont = r[3 * S]; h.normFn = r[4 * S];
  h.du = mk(r[5 * S], r[6 * S], 0.0);
  h.v_tan = mk(r[7 * S], r[8 * S], 0.0);
  double ux = (r[9 * S] != 0.0) ? 0.0 : c.ut_prev[i], uy = c.ut_prev[d.np + i], uz = 0.0;
  bool slid;
  const V3 Ft = friction(c, h, ux, uy, uz, slid);
  c.ut_prev[i] = ux; c.ut_prev[d.np + i] = uy;
  d.contforce[i] += Ft.x;
  d.contforce[d.np + i] += Ft.y;
  if (slid && i + 1 < d.nn) { // Contact.C:297 `for d<3: ut_prev[m_dim*i+d] = 0` reaches the next node's x slot
    const bool next_is_hit = (t + 1 < c.n_ext) && (c.ext_nodes[t + 1] == i + 1) && (c.rec[t + 1] != 0.0);
    if (!next_is_hit) c.ut_prev[i + 1] = 0.0; // a hit node i+1 already started from the zeroed slot and stores its own value
  }
}

__global__ void k_cflag(WfDev d) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
  double s = 0.0;
  for (int c = 0; c < d.dim; c++) { double q = d.contforce[(long long)c * d.np + n]; s += q * q; }
  d.cflag[n] = s > 0 ? 1 : 0;
}
__global__ void k_fill_int(int *p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

template <class T>
int upload(wf_engine *E, T **dst, const std::vector<T> &src) {
  if (dalloc(E, dst, src.size())) return 1;
  if (!src.empty()) CK(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, E->stream));
  return 0;
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------
static int calc_ext_face_areas(wf_engine *E) {
  WfContact &C = E->C;
  if (C.n_xf == 0) return 0;
  CK(cudaMemsetAsync(C.elem_area, 0, sizeof(double) * E->d.ep, E->stream));
  k_xf_area<<<cdiv(C.n_xf, 256), 256, 0, E->stream>>>(E->d, C);
  k_xn_area<<<cdiv(C.n_ext, 256), 256, 0, E->stream>>>(E->d, C);
  k_xe_area<<<cdiv(C.n_xf, 256), 256, 0, E->stream>>>(E->d, C);
  return 0;
}

// Domain_d::SearchExtNodes (Domain_d.C:110-205): external faces / nodes, then CalcExtFaceAreas
extern "C" int wf_SearchExtNodes(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->meshed, "SearchExtNodes needs the mesh");
  NEED(!E->distributed, "contact is not available on a partitioned mesh");
  NEED((E->dim == 3 && E->k == 4) || (E->dim == 2 && E->k == 4),
       "SearchExtNodes: the reference's face tables cover tetrahedra (3D) and quadrilaterals (2D) only (Domain_d.h:172-193, 246-249, 785-791)");
  NEED(!E->ext_searched, "SearchExtNodes already done for this mesh");
  CK(cudaSetDevice(E->device));
  const int facenod = E->dim == 3 ? 3 : 2;
  const size_t cap = (size_t)E->ne * 4;
  std::vector<int> fnodes(cap * facenod), felem(cap);
  E->h_ext.assign(E->nn, 0);
  int n_total = 0, n_xf = 0;
  if (wf_host_ext_faces(E->dim, E->k, E->nn, E->ne, E->h_elnod.data(), E->h_ext.data(), &n_total, &n_xf, fnodes.data(), felem.data()))
    FAIL("wf_host_ext_faces failed");
  E->face_count = n_total;
  WfContact &C = E->C;
  memset(&C, 0, sizeof(C));
  C.facenod = facenod; C.n_xf = n_xf;
  std::vector<int> ext_list, ext_index(E->nn, -1);
  for (int n = 0; n < E->nn; n++)
    if (E->h_ext[n]) { ext_index[n] = (int)ext_list.size(); ext_list.push_back(n); }
  C.n_ext = (int)ext_list.size();
  std::vector<int> xf_soa((size_t)facenod * n_xf), xf_elem(felem.begin(), felem.begin() + n_xf);
  if (!E->iperm.empty()) // device element arrays are in the internal element order
    for (int f = 0; f < n_xf; f++) xf_elem[f] = E->iperm[xf_elem[f]];
  std::vector<int> ptr(C.n_ext + 1, 0);
  for (int f = 0; f < n_xf; f++)
    for (int q = 0; q < facenod; q++) {
      const int n = fnodes[(size_t)f * facenod + q];
      xf_soa[(size_t)q * n_xf + f] = n;
      ptr[ext_index[n] + 1]++;
    }
  for (int t = 0; t < C.n_ext; t++) ptr[t + 1] += ptr[t];
  std::vector<int> fill(ptr.begin(), ptr.end() - 1), xn_faces(ptr[C.n_ext]);
  for (int f = 0; f < n_xf; f++) // ascending face index per node == faceList order of the reference's += loop
    for (int q = 0; q < facenod; q++) xn_faces[fill[ext_index[fnodes[(size_t)f * facenod + q]]]++] = f;
  int *d_ext, *d_xf, *d_xe, *d_ptr, *d_xnf;
  if (upload(E, &d_ext, ext_list) || upload(E, &d_xf, xf_soa) || upload(E, &d_xe, xf_elem) || upload(E, &d_ptr, ptr) ||
      upload(E, &d_xnf, xn_faces))
    return 1;
  C.ext_nodes = d_ext; C.xf_nodes = d_xf; C.xf_elem = d_xe; C.xn_ptr = d_ptr; C.xn_faces = d_xnf;
  WfDev &d = E->d;
  const size_t nv = (size_t)E->dim * d.np;
  if (dalloc(E, &C.nodlen, (size_t)C.n_ext) || dalloc(E, &C.node_area, (size_t)d.np) || dalloc(E, &C.ut_prev, nv) ||
      dalloc(E, &C.mesh_in_contact, (size_t)d.np) || dalloc(E, &C.xf_area, (size_t)n_xf) || dalloc(E, &C.elem_area, (size_t)d.ep) ||
      dalloc(E, &C.rec, (size_t)10 * std::max(C.n_ext, 1)) || dalloc(E, &C.cand_count, 1) || dalloc(E, &C.cand, (size_t)std::max(C.n_ext, 1)))
    return 1;
  k_fill_int<<<cdiv(d.np, 256), 256, 0, E->stream>>>(C.mesh_in_contact, (int)d.np, -1);
  if (calc_ext_face_areas(E)) return 1;
  CK(cudaStreamSynchronize(E->stream));
  E->ext_searched = true;
  return wf_check_launch(E, "wf_SearchExtNodes");
}

extern "C" int wf_CalcExtFaceAreas(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->ext_searched, "CalcExtFaceAreas needs wf_SearchExtNodes");
  NEED(!E->predicted, "engine is mid-batch");
  CK(cudaSetDevice(E->device));
  if (calc_ext_face_areas(E)) return 1;
  return wf_check_launch(E, "wf_CalcExtFaceAreas");
}

// Domain_d::setTriMesh with the flattened TriMesh_d the reference ends up with after AxisPlaneMesh / AddMesh
// (main.C:672-708, 775-828): node / node_v as xyz triples, elnode with 3 (3D) or 2 (2D) ids per facet, the
// initial facet normals and ele_mesh_id.
extern "C" int wf_set_trimesh(wf_engine *E, int dimension, int n_nodes, int n_elems, const double *node, const double *node_v,
                              const int *elnode, const double *normal, const int *ele_mesh_id) { WF_NULLCHK(E);
  NEED(E->ext_searched, "wf_set_trimesh needs wf_SearchExtNodes (main.C:650 comes first)");
  NEED(!E->trimesh_set, "rigid surfaces already set");
  NEED(dimension == E->dim, "TriMesh_d::dimension must equal the domain's dimension");
  NEED(n_nodes > 0 && n_elems > 0 && node && node_v && elnode && normal && ele_mesh_id, "empty rigid surface");
  const int nen = dimension == 3 ? 3 : 2;
  for (long long q = 0; q < (long long)nen * n_elems; q++) NEED(elnode[q] >= 0 && elnode[q] < n_nodes, "rigid surface connectivity out of range");
  CK(cudaSetDevice(E->device));
  WfContact &C = E->C;
  C.tm_dim = dimension; C.tm_nn = n_nodes; C.tm_ne = n_elems;
  std::vector<double> vn(node, node + 3 * (size_t)n_nodes), vv(node_v, node_v + 3 * (size_t)n_nodes), vm(normal, normal + 3 * (size_t)n_elems);
  std::vector<int> ve(elnode, elnode + (size_t)nen * n_elems), vi(ele_mesh_id, ele_mesh_id + n_elems);
  double *vorig; int *de, *di;
  if (upload(E, &C.tm_node, vn) || upload(E, &C.tm_node_v, vv) || upload(E, &vorig, vv) || upload(E, &C.tm_normal, vm) ||
      upload(E, &de, ve) || upload(E, &di, vi) || dalloc(E, &C.tm_pplane, (size_t)n_elems))
    return 1;
  C.tm_v_orig = vorig; C.tm_elnode = de; C.tm_mesh_id = di;
  CK(cudaStreamSynchronize(E->stream));
  E->trimesh_set = true;
  return 0;
}

// friction + penalty factor (main.C:716-725), CalcSpheres + setContactOn (main.C:842-847), SetEndTime (Domain_d.h:636)
extern "C" int wf_set_contact(wf_engine *E, double mu_sta, double mu_dyn, double penalty_factor, double end_time) { WF_NULLCHK(E);
  NEED(E->trimesh_set, "wf_set_contact needs wf_set_trimesh");
  NEED(E->material_set, "wf_set_contact needs the material (contact stiffness uses E)");
  NEED(!E->inited, "contact must be switched on before wf_init");
  CK(cudaSetDevice(E->device));
  WfContact &C = E->C;
  WfDev &d = E->d;
  C.mu_sta = mu_sta; C.mu_dyn = mu_dyn;
  C.contPF = penalty_factor > -1.0 ? penalty_factor : 0.1; // Domain_d.h:256
  C.young = E->mat.E;
  E->end_t = end_time;
  if (!d.contforce && (dalloc(E, &d.contforce, (size_t)E->dim * d.np) || dalloc(E, &d.cflag, (size_t)d.np))) return 1;
  k_trimesh_update<<<1, 1024, 0, E->stream>>>(C, 1.0, 0.0, 0); // CalcSpheres -> UpdatePlaneCoeff with the initial normals
  E->contact = true;
  E->P.alpha_contact = E->stab.alpha_contact;
  E->P.hg_coeff_contact = E->stab.hg_coeff_contact;
  return wf_check_launch(E, "wf_set_contact");
}

// heatCondCoeff / dieTemp of the rigid surfaces (main.C:718-719): contact heat flow into the nodal temperature
extern "C" int wf_set_contact_heat(wf_engine *E, double heat_cond, double T_const) { WF_NULLCHK(E);
  NEED(E->contact, "wf_set_contact_heat needs wf_set_contact");
  NEED(E->d.T, "wf_set_contact_heat needs wf_set_thermal");
  NEED(!E->inited, "contact heat must be set before wf_init");
  CK(cudaSetDevice(E->device));
  if (!E->d.q_cont_conv && dalloc(E, &E->d.q_cont_conv, (size_t)E->d.np)) return 1;
  E->C.heat_cond = heat_cond; E->C.T_const = T_const;
  return 0;
}

int wf_contact_refresh_nodlen(wf_engine *E) {
  if (!E->ext_searched || !E->elem_length_valid || E->C.n_ext == 0) return 0;
  k_nodlen<<<cdiv(E->C.n_ext, 128), 128, 0, E->stream>>>(E->d, E->C, E->elem_length);
  return 0;
}

// Solver_explicit.C:168-173 (m_v_orig) and :286-289 (ut_prev = 0)
int wf_contact_init(wf_engine *E) {
  if (!E->contact) return 0;
  NEED(E->elem_length_valid, "contact needs m_elem_length: call wf_calcMinEdgeLength before wf_init (main.C:862)");
  WfContact &C = E->C;
  CK(cudaMemcpyAsync((void *)C.tm_v_orig, C.tm_node_v, sizeof(double) * 3 * C.tm_nn, cudaMemcpyDeviceToDevice, E->stream));
  CK(cudaMemsetAsync(C.ut_prev, 0, sizeof(double) * E->dim * E->d.np, E->stream));
  E->P.alpha_contact = E->stab.alpha_contact;
  E->P.hg_coeff_contact = E->stab.hg_coeff_contact;
  return wf_contact_refresh_nodlen(E);
}

int wf_contact_step_begin(wf_engine *E) {
  if (E->ext_searched && E->dim > 2 && E->step_count % 10 == 0) return calc_ext_face_areas(E);
  return 0;
}

static int contact_forces(wf_engine *E, const double *acc) {
  WfContact &C = E->C;
  if (C.n_ext == 0) return 0;
  (void)acc; // x_pred = x + v dt + a dt^2/2 only feeds quantities the reference computes and never uses (Contact.C:77-79, 236-238)
  CK(cudaMemsetAsync(C.cand_count, 0, sizeof(int), E->stream));
  k_contact_filter<<<cdiv(C.n_ext, CONTACT_TPB), CONTACT_TPB, 0, E->stream>>>(E->d, C, E->P.dt, C.cand_count, C.cand);
  if (E->dim == 3) k_contact_exact<true><<<cdiv((long long)C.n_ext * 32, 256), 256, 0, E->stream>>>(E->d, C, E->P.dt, C.cand_count, C.cand);
  else k_contact_exact<false><<<cdiv((long long)C.n_ext * 32, 256), 256, 0, E->stream>>>(E->d, C, E->P.dt, C.cand_count, C.cand);
  if (E->dim == 2) {
    k_friction2d_chain<<<1, 1024, 0, E->stream>>>(C);
    k_friction2d_apply<<<cdiv(C.n_ext, 128), 128, 0, E->stream>>>(E->d, C);
  }
  return 0;
}
int wf_contact_forces(wf_engine *E) { return E->contact ? contact_forces(E, E->d.prev_a) : 0; }

int wf_contact_step_end(wf_engine *E) {
  if (!E->contact) return 0;
  const double RAMP_FRACTION = 1.0e-2; // Solver_explicit.C:309
  double f = 1.0;
  if (E->time < RAMP_FRACTION * E->end_t) f = pow(E->time / (RAMP_FRACTION * E->end_t), 0.5);
  k_trimesh_update<<<1, 1024, 0, E->stream>>>(E->C, f, E->P.dt, 1);
  return 0;
}

// unfused entry points
extern "C" int wf_CalcContactForces(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->contact && E->inited, "CalcContactForces needs wf_set_contact and wf_init");
  NEED(!E->predicted, "engine is mid-batch");
  CK(cudaSetDevice(E->device));
  if (contact_forces(E, (E->a_in_dbg && E->d.a) ? E->d.a : E->d.prev_a)) return 1;
  return wf_check_launch(E, "wf_CalcContactForces");
}
extern "C" int wf_MoveTriMesh(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->contact && E->inited, "MoveTriMesh needs wf_set_contact and wf_init");
  CK(cudaSetDevice(E->device));
  if (wf_contact_step_end(E)) return 1;
  return wf_check_launch(E, "wf_MoveTriMesh");
}
extern "C" int wf_get_trimesh_counts(wf_engine *E, int *dimension, int *n_nodes, int *n_elems) { WF_NULLCHK(E);
  if (dimension) *dimension = E->C.tm_dim;
  if (n_nodes) *n_nodes = E->trimesh_set ? E->C.tm_nn : 0;
  if (n_elems) *n_elems = E->trimesh_set ? E->C.tm_ne : 0;
  return 0;
}

// array access by Domain_d / TriMesh_d member name.  kind: 0 node vector [dim][np], 1 node scalar, 2 element scalar,
// 3 raw bytes on the device, 4 host bytes
bool wf_contact_lookup(wf_engine *E, const std::string &nm, void **dev, size_t *bytes, int *kind) {
  WfContact &C = E->C;
  const size_t nn = (size_t)E->nn, ne = (size_t)E->ne;
  if (nm == "ext_nodes" && E->ext_searched) { *dev = (void *)E->h_ext.data(); *bytes = nn; *kind = 4; return true; }
  if (!E->ext_searched) return false;
  if (nm == "node_area") { *dev = C.node_area; *bytes = 8 * nn; *kind = 1; return true; }
  if (nm == "m_elem_area") { *dev = C.elem_area; *bytes = 8 * ne; *kind = 2; return true; }
  if (nm == "ut_prev") { *dev = C.ut_prev; *bytes = 8 * nn * E->dim; *kind = 0; return true; }
  if (nm == "m_mesh_in_contact") { *dev = C.mesh_in_contact; *bytes = 4 * nn; *kind = 3; return true; }
  if (nm == "contforce" && E->d.contforce) { *dev = E->d.contforce; *bytes = 8 * nn * E->dim; *kind = 0; return true; }
  if (!E->trimesh_set) return false;
  if (nm == "trimesh.node") { *dev = C.tm_node; *bytes = 24 * (size_t)C.tm_nn; *kind = 3; return true; }
  if (nm == "trimesh.node_v") { *dev = C.tm_node_v; *bytes = 24 * (size_t)C.tm_nn; *kind = 3; return true; }
  if (nm == "trimesh.normal") { *dev = C.tm_normal; *bytes = 24 * (size_t)C.tm_ne; *kind = 3; return true; }
  if (nm == "trimesh.pplane") { *dev = C.tm_pplane; *bytes = 8 * (size_t)C.tm_ne; *kind = 3; return true; }
  return false;
}

int wf_contact_after_set(wf_engine *E, const std::string &nm) {
  if (nm == "contforce" && E->d.cflag) k_cflag<<<cdiv(E->nn, 256), 256, 0, E->stream>>>(E->d);
  return 0;
}
