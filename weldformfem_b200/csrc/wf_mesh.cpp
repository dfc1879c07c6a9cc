// wf_mesh.cpp — host-side integer work of the engine: box mesher, node->element connectivity,
// element-block partition and halo lists.  No CUDA here, so the CPU test-suite can check every
// integer artefact bit-exactly against the oracle without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/wf_engine.h"
#include "wf_host.h"

// ---- Domain_d::AddBoxLength (src/common/Domain_d.C:1136-1504) -------------------------------------
// nel[i] = (int)(L_i/(2r)); node id = i + (nx+1)(j + (ny+1)k); coordinates by repeated "+= 2r";
// hexa node order :1321-1332, quad :1273-1278, triangle split :1296-1303, 6-tet split :1383-1388.
void wf_box_dims(const double L[3], double r, int tritet, WfBox *b) {
  b->dim = (L[2] > 0.0) ? 3 : 2;
  b->nel[0] = (int)(L[0] / (2.0 * r));
  b->nel[1] = (int)(L[1] / (2.0 * r));
  b->nel[2] = (b->dim == 3) ? (int)(L[2] / (2.0 * r)) : 1;
  b->tritet = tritet;
  b->k = (b->dim == 3) ? (tritet ? 4 : 8) : (tritet ? 3 : 4);
  b->per_cell = tritet ? (b->dim == 3 ? 6 : 2) : 1;
  b->nn = (long long)(b->nel[0] + 1) * (b->nel[1] + 1) * (b->dim == 3 ? b->nel[2] + 1 : 1);
  b->ne = (long long)b->nel[0] * b->nel[1] * b->nel[2] * b->per_cell;
}

static const int kTetSplit[6][4] = {{0, 1, 3, 5}, {1, 2, 3, 5}, {0, 5, 3, 4}, {4, 5, 3, 7}, {5, 6, 3, 7}, {5, 2, 3, 6}};

// connectivity of element e of the box (any e, O(1))
void wf_box_elem_nodes(const WfBox &b, long long e, unsigned *out) {
  const long long cell = e / b.per_cell;
  const int sub = (int)(e - cell * b.per_cell);
  const int nx1 = b.nel[0] + 1;
  if (b.dim == 2) {
    const int ey = (int)(cell / b.nel[0]), ex = (int)(cell - (long long)ey * b.nel[0]);
    const unsigned nb1 = (unsigned)(nx1 * ey + ex), nb2 = (unsigned)(nx1 * (ey + 1) + ex);
    if (!b.tritet) { out[0] = nb1; out[1] = nb1 + 1; out[2] = nb2 + 1; out[3] = nb2; }
    else if (sub == 0) { out[0] = nb1; out[1] = nb1 + 1; out[2] = nb2; }
    else { out[0] = nb1 + 1; out[1] = nb2 + 1; out[2] = nb2; }
    return;
  }
  const long long nxy = (long long)b.nel[0] * b.nel[1];
  const int ez = (int)(cell / nxy);
  const long long rem = cell - (long long)ez * nxy;
  const int ey = (int)(rem / b.nel[0]), ex = (int)(rem - (long long)ey * b.nel[0]);
  const long long nnodz = (long long)nx1 * (b.nel[1] + 1);
  const long long nb1 = nnodz * ez + (long long)nx1 * ey + ex, nb2 = nnodz * ez + (long long)nx1 * (ey + 1) + ex;
  const unsigned nh[8] = {(unsigned)nb1, (unsigned)(nb1 + 1), (unsigned)(nb2 + 1), (unsigned)nb2,
                          (unsigned)(nb1 + nnodz), (unsigned)(nb1 + nnodz + 1), (unsigned)(nb2 + nnodz + 1),
                          (unsigned)(nb2 + nnodz)};
  if (!b.tritet) { for (int i = 0; i < 8; i++) out[i] = nh[i]; }
  else { for (int i = 0; i < 4; i++) out[i] = nh[kTetSplit[sub][i]]; }
}

// per-axis coordinate tables, accumulated exactly like the reference's nested loops (:1205-1234)
void wf_box_axes(const WfBox &b, const double V[3], double r, std::vector<double> ax[3]) {
  for (int a = 0; a < 3; a++) {
    int cnt = (a < b.dim) ? b.nel[a] + 1 : 1;
    ax[a].resize(cnt);
    double X = V[a];
    for (int i = 0; i < cnt; i++) { ax[a][i] = X; X = X + 2.0 * r; }
  }
}

void wf_box_node_xyz(const WfBox &b, const std::vector<double> ax[3], long long n, double *out) {
  const int nx1 = b.nel[0] + 1, ny1 = b.nel[1] + 1;
  const long long kz = n / ((long long)nx1 * ny1);
  const long long rem = n - kz * nx1 * ny1;
  const int j = (int)(rem / nx1), i = (int)(rem - (long long)j * nx1);
  out[0] = ax[0][i]; out[1] = ax[1][j];
  if (b.dim == 3) out[2] = ax[2][kz];
}

extern "C" int wf_host_box_counts(const double L[3], double r, int tritet, int *dim, int *nodxelem, int *n_nodes,
                                  int *n_elems) {
  WfBox b;
  wf_box_dims(L, r, tritet, &b);
  if (b.nn > 2147483647LL || b.ne * b.k > 2147483647LL) return 1;
  *dim = b.dim; *nodxelem = b.k; *n_nodes = (int)b.nn; *n_elems = (int)b.ne;
  return 0;
}

extern "C" int wf_host_gen_box(const double V[3], const double L[3], double r, int tritet, double *x, unsigned *elnod) {
  WfBox b;
  wf_box_dims(L, r, tritet, &b);
  std::vector<double> ax[3];
  wf_box_axes(b, V, r, ax);
  for (long long n = 0; n < b.nn; n++) wf_box_node_xyz(b, ax, n, x + n * b.dim);
  for (long long e = 0; e < b.ne; e++) wf_box_elem_nodes(b, e, elnod + e * b.k);
  return 0;
}

// ---- Domain_d::setNodElem (src/common/Domain_d.C:1508-1611) -----------------------------------------
// count pass, exclusive prefix sum, fill in ascending element id then local node: every node's list is
// sorted by element id.  Returns non-zero if a connectivity entry is out of range.
extern "C" int wf_host_nodel(int n_nodes, int n_elems, int k, const unsigned *elnod, int *offset, int *count,
                             int *nodel, int *nodel_loc) {
  for (int n = 0; n < n_nodes; n++) count[n] = 0;
  const long long tot_entries = (long long)n_elems * k;
  for (long long i = 0; i < tot_entries; i++) {
    if (elnod[i] >= (unsigned)n_nodes) return 1;
    count[elnod[i]]++;
  }
  long long tot = 0;
  for (int n = 0; n < n_nodes; n++) { offset[n] = (int)tot; tot += count[n]; }
  for (int n = 0; n < n_nodes; n++) count[n] = 0;
  for (int e = 0; e < n_elems; e++)
    for (int ln = 0; ln < k; ln++) {
      const int n = (int)elnod[(long long)e * k + ln];
      nodel[offset[n] + count[n]] = e;
      nodel_loc[offset[n] + count[n]] = ln;
      count[n]++;
    }
  return 0;
}

// ---- internal element order ("element order") ---------------------------------------------------
// The element passes give one CTA 128 consecutive elements and one warp 32 (a force tile); how many DISTINCT nodes
// such a run touches decides the staging traffic of E1/E2 and the number of force partials N2 gathers.  In the
// reference's numbering (AddBoxLength: x fastest) a 32-element run is a 1-D row segment with 132 distinct nodes; along
// a Morton (Z-order) curve of the element centroids it is a 4x4x2 brick with 75, and 128 elements are an 8x4x4 brick
// with 225 instead of ~516.  The engine therefore keeps its element arrays in Morton order internally and converts
// at the ABI (wf_get_array / wf_set_array); node->element lists stay in ascending USER element id, so every nodal
// sum keeps the reference's order.  Integer work, restated in numpy by the CPU tests:
//   box = bounding box of the nodes, ext_c = hi_c - lo_c; active dims = those with ext_c > 0 (na of them);
//   cells = max(1, n_elems / per_cell), per_cell = 1 (hexa, quad), 6 (tetra), 2 (triangle);
//   h = pow(prod ext_c / cells, 1/na); n_c = max(1, (int)(ext_c / h + 0.5));
//   q_c = min(n_c - 1, (int)((centroid_c - lo_c) / ext_c * n_c)), centroid = (sum of the k node coordinates) / k;
//   key = bit-interleave(q_0, q_1[, q_2]) with q_0 in the lowest bit; order = ascending (key, user id).
//   Hexahedra: key = (bit-interleave(q_2 >> 1, q_0 >> 2, q_1 >> 2) << 5) | (q_2 & 1) << 4 | (q_1 & 3) << 2 | (q_0 & 3):
//   the same bricks, with the 32 elements of a tile ordered layer by layer.
// mode 0 = identity.  perm[internal] = user.
static inline uint64_t spread_bits(uint32_t v, int ndim) {
  uint64_t r = 0;
  for (int b = 0; b < (ndim == 3 ? 21 : 31); b++) r |= (uint64_t)((v >> b) & 1u) << (ndim * b);
  return r;
}
// keys (optional): the sort key of every element, in the resulting INTERNAL order (mode 1 only; see wf_host_brick_plan)
extern "C" int wf_host_elem_order_keys(int dim, int k, int n_nodes, int n_elems, const double *x, const unsigned *elnod, int mode,
                                       int *perm, unsigned long long *keys_out) {
  if (n_elems <= 0 || n_nodes <= 0 || !x || !elnod || !perm || (dim != 2 && dim != 3)) return 1;
  if (mode == 0) {
    if (keys_out) return 1;
    for (int e = 0; e < n_elems; e++) perm[e] = e;
    return 0;
  }
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int c = 0; c < dim; c++) lo[c] = hi[c] = x[c];
  for (long long n = 1; n < n_nodes; n++)
    for (int c = 0; c < dim; c++) {
      const double v = x[n * dim + c];
      if (v < lo[c]) lo[c] = v;
      if (v > hi[c]) hi[c] = v;
    }
  const int per_cell = (k == 8) ? 1 : (dim == 3 ? 6 : (k == 4 ? 1 : 2));
  const long long cells = std::max<long long>(1, n_elems / per_cell);
  double ext[3] = {0, 0, 0}, vol = 1.0;
  int na = 0;
  for (int c = 0; c < dim; c++) {
    ext[c] = hi[c] - lo[c];
    if (ext[c] > 0.0) { vol *= ext[c]; na++; }
  }
  int ncell[3] = {1, 1, 1};
  if (na > 0) {
    const double h = pow(vol / (double)cells, 1.0 / (double)na);
    const int cap = dim == 3 ? (1 << 21) : (1 << 30);
    for (int c = 0; c < dim; c++)
      if (ext[c] > 0.0) {
        const double r = ext[c] / h + 0.5;
        ncell[c] = r >= (double)cap ? cap : std::max(1, (int)r);
      }
  }
  for (size_t i = 0; i < (size_t)n_elems * k; i++)
    if (elnod[i] >= (unsigned)n_nodes) return 2;
  std::vector<std::pair<uint64_t, int>> keys((size_t)n_elems);
  for (int e = 0; e < n_elems; e++) {
    uint32_t q[3] = {0, 0, 0};
    for (int c = 0; c < dim; c++) {
      if (!(ext[c] > 0.0)) continue;
      double s = 0.0;
      for (int a = 0; a < k; a++) s += x[(size_t)elnod[(size_t)e * k + a] * dim + c];
      const double t = (s / (double)k - lo[c]) / ext[c] * (double)ncell[c];
      int qi = (int)t;
      if (qi < 0) qi = 0;
      if (qi > ncell[c] - 1) qi = ncell[c] - 1;
      q[c] = (uint32_t)qi;
    }
    uint64_t key;
    if (k == 8 && dim == 3) {
      // hexahedra: Morton order of the 4x4x2 tiles (z lowest, then x, y: 128 consecutive elements = 8x4x4), and inside
      // a tile z-layer, then y, then x: lanes 0-15 / 16-31 of a warp are the two 4x4 layers (see wf_host_run_slots)
      const uint32_t tx = q[0] >> 2, ty = q[1] >> 2, tz = q[2] >> 1;
      const uint64_t tkey = spread_bits(tz, 3) | (spread_bits(tx, 3) << 1) | (spread_bits(ty, 3) << 2);
      key = (tkey << 5) | ((q[2] & 1u) << 4) | ((q[1] & 3u) << 2) | (q[0] & 3u);
    } else {
      key = spread_bits(q[0], dim) | (spread_bits(q[1], dim) << 1);
      if (dim == 3) key |= spread_bits(q[2], dim) << 2;
    }
    keys[(size_t)e] = std::make_pair(key, e);
  }
  std::sort(keys.begin(), keys.end());
  for (int e = 0; e < n_elems; e++) perm[e] = keys[(size_t)e].second;
  if (keys_out)
    for (int e = 0; e < n_elems; e++) keys_out[e] = keys[(size_t)e].first;
  return 0;
}
extern "C" int wf_host_elem_order(int dim, int k, int n_nodes, int n_elems, const double *x, const unsigned *elnod, int mode,
                                  int *perm) {
  return wf_host_elem_order_keys(dim, k, n_nodes, n_elems, x, elnod, mode, perm, nullptr);
}

// ---- thread slots of the brick passes (k_elem_vol_brick, k_elem_main_hex_brick) -------------------------------------
// The hexahedron key of wf_host_elem_order is (tile Morton code << 5) | position in the 4x4x2 tile, and the two lowest
// bits of the tile code are (z, x) of the tile inside its 8x4x4 group: key >> 7 names the group = one CTA, key & 127
// the thread (warp = tile, lane = position).  When every element of the mesh has its own key (one element per cell:
// any mapped structured mesh), giving thread key & 127 of CTA rank(key >> 7) to the element makes EVERY CTA a clipped
// 8x4x4 brick, also along ragged mesh boundaries (the compact numbering 128 b + t shifts by the missing elements of
// every clipped brick it has passed); threads whose cell does not exist idle.  keys: ascending, internal order.
// slot_elem[cta * 128 + thread] = internal element or -1; call with slot_elem = NULL for n_cta.  Returns 1 when the
// keys are not strictly ascending (two elements in one cell): the caller keeps the compact numbering.
extern "C" int wf_host_brick_plan(int n_elems, const unsigned long long *keys, int *n_cta, int *slot_elem) {
  if (n_elems <= 0 || !keys || !n_cta) return 1;
  long long nc = 1;
  for (int e = 1; e < n_elems; e++) {
    if (keys[e] <= keys[e - 1]) return 1;
    if ((keys[e] >> 7) != (keys[e - 1] >> 7)) nc++;
  }
  if (nc * 128 > 2147483647LL) return 1;
  *n_cta = (int)nc;
  if (!slot_elem) return 0;
  std::fill(slot_elem, slot_elem + nc * 128, -1);
  long long c = 0;
  for (int e = 0; e < n_elems; e++) {
    if (e > 0 && (keys[e] >> 7) != (keys[e - 1] >> 7)) c++;
    slot_elem[c * 128 + (long long)(keys[e] & 127u)] = e;
  }
  return 0;
}

// ---- bank-aware shared-memory slots of the brick kernel (k_elem_main_hex_brick) -----------------------------------
// A 64-bit shared-memory access of a warp is served per half-warp: 16 lanes are conflict-free when their 8-byte words
// fall into 16 different bank pairs (word index mod 16).  In the engine's element order a half-warp is a 4x4 layer of
// a tile, and corner c of those 16 elements are the nodes (x + dx, y + dy) of one node layer, so word index
// x + P*y is conflict-free for every corner iff P = 4 or 12 (mod 16).  The node lists of a CTA / tile are ascending
// node ids = runs of consecutive ids (one run per mesh row: 9 or 5 nodes); placing run r at the first free slot
// congruent to 12 r (mod 16) gives exactly that lattice with P = 12 on a structured mesh, and some valid layout on
// any other (only the conflict count depends on it).  Returns the number of slots used.
extern "C" int wf_host_run_slots(int n, const int *sorted_ids, int *slots) {
  int cur = 0, r = 0;
  for (int a = 0; a < n;) {
    int b = a;
    while (b + 1 < n && sorted_ids[b + 1] == sorted_ids[b] + 1) b++;
    while ((cur & 15) != ((12 * r) & 15)) cur++;
    for (int q = a; q <= b; q++) slots[q] = cur++;
    r++;
    a = b + 1;
  }
  return cur;
}

// ---- force tiles of the tile-reduced force path (WfDev::ftile) ------------------------------------------------
// Tile w = thread slots [32w, 32w+32) of the main element pass (one warp); slot s holds element slot_elem[s] (-1 = idle
// thread; slot_elem = NULL: slot s = element s, the compact numbering).  Per tile: the ascending list of its unique
// nodes; tidx = position of every element node in that list.  Hexahedra accumulate in conflict-free rounds (one
// per local corner), which requires that no two elements of a tile reference the same node through the same
// corner; tetrahedra pull through an incidence table (CSR by unique node, entries ascending element then corner).
// Node n owns one entry per tile that references it, ascending tile order: slots = tile*3*stride + position.
void wf_force_tiles_build(int nn, int ne, const int *slot_elem, int k, int dim, long long ep, const unsigned *elnod,
                          WfForceTiles &T) {
  T = WfForceTiles();
  T.k = k;
  if (!((dim == 3 && (k == 8 || k == 4)) || (dim == 2 && (k == 4 || k == 3))) || ne <= 0) return;
  if (slot_elem && k != 8) return; // the incidence tables of the pull form are built for the compact numbering only
  const int ntile = (ne + 31) / 32;
  T.n_tiles = ntile;
  T.tidx.assign((size_t)k * ep, 0);
  std::vector<int> toff((size_t)ntile + 1, 0), tnodes, tmp;
  tnodes.reserve((size_t)ne * (k == 8 ? 5 : 1));
  int wmax = 0;
  bool ok = true, conflict = false;
  std::vector<int> stamp(256, -1);
  auto elem = [&](int s) { return slot_elem ? slot_elem[s] : s; };
  for (int w = 0; w < ntile && ok; w++) {
    const int e0 = w * 32, e1 = std::min(ne, e0 + 32);
    tmp.clear();
    for (int s = e0; s < e1; s++)
      if (elem(s) >= 0) tmp.insert(tmp.end(), elnod + (size_t)elem(s) * k, elnod + (size_t)(elem(s) + 1) * k);
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    if (tmp.size() > 256) { ok = false; break; }
    for (int n = 0; n < k && ok; n++)
      for (int e = e0; e < e1; e++) {
        if (elem(e) < 0) continue;
        const int u = (int)(std::lower_bound(tmp.begin(), tmp.end(), (int)elnod[(size_t)elem(e) * k + n]) - tmp.begin());
        if (stamp[u] == w * k + n) {
          conflict = true;
          if (k == 8) { ok = false; break; } // hexahedra have no pull form
        }
        stamp[u] = w * k + n;
        T.tidx[(size_t)n * ep + e] = (unsigned char)u;
      }
    tnodes.insert(tnodes.end(), tmp.begin(), tmp.end());
    toff[w + 1] = (int)tnodes.size();
    wmax = std::max(wmax, (int)tmp.size());
  }
  const int stride = (wmax + 3) / 4 * 4;
  ok = ok && (long long)ntile * dim * stride < 4294967295LL;
  if (!ok) return;
  T.stride = stride;
  std::vector<int> cnt((size_t)nn, 0);
  for (int g : tnodes) cnt[g]++;
  const int nsl = (nn + 31) / 32;
  T.ptr.assign((size_t)nsl + 1, 0);
  for (int sl = 0; sl < nsl; sl++) {
    int wd = 0;
    for (int n = sl * 32; n < std::min(nn, sl * 32 + 32); n++) wd = std::max(wd, cnt[n]);
    T.ptr[sl + 1] = T.ptr[sl] + 32LL * wd;
  }
  T.slots.assign((size_t)T.ptr[nsl], 0xFFFFFFFFu);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int w = 0; w < ntile; w++)
    for (int i = toff[w]; i < toff[w + 1]; i++) {
      const int g = tnodes[i];
      T.slots[(size_t)(T.ptr[g >> 5] + 32LL * cnt[g] + (g & 31))] = (unsigned)((long long)w * dim * stride + (i - toff[w]));
      cnt[g]++;
    }
  if (k != 8) {
    const int tpitch = (stride + 1 + 32 * k + 3) / 4 * 4;
    T.tpitch = tpitch;
    T.tab.assign((size_t)ntile * tpitch, 0);
    std::vector<int> c2((size_t)stride + 1);
    for (int w = 0; w < ntile; w++) {
      const int e0 = w * 32, e1 = std::min(ne, e0 + 32);
      unsigned char *tb = T.tab.data() + (size_t)w * tpitch;
      std::fill(c2.begin(), c2.end(), 0);
      for (int e = e0; e < e1; e++)
        for (int n = 0; n < k; n++) c2[T.tidx[(size_t)n * ep + e] + 1]++;
      for (int u = 0; u < stride; u++) c2[u + 1] += c2[u];
      for (int u = 0; u <= stride; u++) tb[u] = (unsigned char)c2[u];
      for (int e = e0; e < e1; e++)
        for (int n = 0; n < k; n++) {
          const int u = T.tidx[(size_t)n * ep + e];
          tb[stride + 1 + c2[u]++] = (unsigned char)((e - e0) * k + n);
        }
    }
  }
  T.rounds = !conflict;
  T.usable = true;
}

// test-facing copy of the tables (no GPU needed).  First call with NULL buffers for the sizes:
// info[7] = {usable, n_tiles, stride, tpitch, n_slices, n_slot_entries, conflict-free rounds}; tidx comes back as [e*k + ln].
extern "C" int wf_host_force_tiles(int n_nodes, int n_elems, int k, int dim, const unsigned *elnod, long long *info,
                                   unsigned char *tidx, long long *ptr, unsigned *slots, unsigned char *tab) {
  for (long long i = 0; i < (long long)n_elems * k; i++)
    if (elnod[i] >= (unsigned)n_nodes) return 1;
  WfForceTiles T;
  wf_force_tiles_build(n_nodes, n_elems, nullptr, k, dim, n_elems, elnod, T);
  info[0] = T.usable ? 1 : 0; info[1] = T.n_tiles; info[2] = T.stride; info[3] = T.tpitch;
  info[4] = (n_nodes + 31) / 32; info[5] = (long long)T.slots.size(); info[6] = T.rounds ? 1 : 0;
  if (!T.usable) return 0;
  if (tidx)
    for (int e = 0; e < n_elems; e++)
      for (int n = 0; n < k; n++) tidx[(size_t)e * k + n] = T.tidx[(size_t)n * n_elems + e];
  if (ptr) std::copy(T.ptr.begin(), T.ptr.end(), ptr);
  if (slots) std::copy(T.slots.begin(), T.slots.end(), slots);
  if (tab) std::copy(T.tab.begin(), T.tab.end(), tab);
  return 0;
}

// ---- canonical element-block partition + halo lists (the reference has none) ---------
struct wf_partition {
  int nranks, rank, k;
  int elem_begin, elem_end;
  std::vector<int> l2g;           // local node -> global id, ascending
  std::vector<unsigned> elnod;    // local connectivity (local node ids)
  std::vector<int> neigh, halo_offset, halo_nodes;
  bool is_box = false;
  WfBox box;
  double V[3], r;
};

static inline long long block_begin(long long ne, int P, int p) { return (ne * p) / P; }

template <class ConnFn>
static void build_partition(wf_partition *pt, long long n_nodes, long long n_elems, ConnFn conn) {
  const int P = pt->nranks, k = pt->k;
  pt->elem_begin = (int)block_begin(n_elems, P, pt->rank);
  pt->elem_end = (int)block_begin(n_elems, P, pt->rank + 1);
  unsigned tmp[WF_MAXK_HOST];
  // local nodes = nodes referenced by owned elements, ascending global id
  std::vector<int> &l2g = pt->l2g;
  l2g.clear();
  l2g.reserve((size_t)(pt->elem_end - pt->elem_begin) * 2);
  for (long long e = pt->elem_begin; e < pt->elem_end; e++) {
    conn(e, tmp);
    for (int i = 0; i < k; i++) l2g.push_back((int)tmp[i]);
  }
  std::sort(l2g.begin(), l2g.end());
  l2g.erase(std::unique(l2g.begin(), l2g.end()), l2g.end());
  auto g2l = [&](unsigned g) { return (unsigned)(std::lower_bound(l2g.begin(), l2g.end(), (int)g) - l2g.begin()); };
  pt->elnod.resize((size_t)(pt->elem_end - pt->elem_begin) * k);
  for (long long e = pt->elem_begin; e < pt->elem_end; e++) {
    conn(e, tmp);
    for (int i = 0; i < k; i++) pt->elnod[(size_t)(e - pt->elem_begin) * k + i] = g2l(tmp[i]);
  }
  // shared nodes: for every other rank q, the global ids it references that are also local here.
  // Only elements of q can reference a node; scan q's block and keep hits (ascending, unique).
  pt->neigh.clear(); pt->halo_offset.assign(1, 0); pt->halo_nodes.clear();
  (void)n_nodes;
  for (int q = 0; q < P && !l2g.empty(); q++) { // a rank that owns no elements has no shared nodes
    if (q == pt->rank) continue;
    const long long qb = block_begin(n_elems, P, q), qe = block_begin(n_elems, P, q + 1);
    std::vector<int> hits;
    for (long long e = qb; e < qe; e++) {
      conn(e, tmp);
      for (int i = 0; i < k; i++) {
        const int g = (int)tmp[i];
        if (g < l2g.front() || g > l2g.back()) continue;
        auto it = std::lower_bound(l2g.begin(), l2g.end(), g);
        if (it != l2g.end() && *it == g) hits.push_back((int)(it - l2g.begin()));
      }
    }
    if (hits.empty()) continue;
    std::sort(hits.begin(), hits.end());
    hits.erase(std::unique(hits.begin(), hits.end()), hits.end());
    pt->neigh.push_back(q);
    pt->halo_nodes.insert(pt->halo_nodes.end(), hits.begin(), hits.end());
    pt->halo_offset.push_back((int)pt->halo_nodes.size());
  }
}

extern "C" int wf_partition_build(wf_partition **out, int nranks, int rank, int k, int n_nodes, int n_elems,
                                  const unsigned *elnod) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || k < 1 || k > WF_MAXK_HOST) return 1;
  wf_partition *pt = new wf_partition();
  pt->nranks = nranks; pt->rank = rank; pt->k = k;
  build_partition(pt, n_nodes, n_elems, [&](long long e, unsigned *o) {
    for (int i = 0; i < k; i++) o[i] = elnod[e * k + i];
  });
  *out = pt;
  return 0;
}

extern "C" int wf_partition_build_box(wf_partition **out, int nranks, int rank, const double V[3], const double L[3],
                                      double r, int tritet) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks) return 1;
  wf_partition *pt = new wf_partition();
  pt->nranks = nranks; pt->rank = rank;
  pt->is_box = true;
  wf_box_dims(L, r, tritet, &pt->box);
  pt->k = pt->box.k;
  for (int i = 0; i < 3; i++) pt->V[i] = V[i];
  pt->r = r;
  const WfBox b = pt->box;
  build_partition(pt, b.nn, b.ne, [&](long long e, unsigned *o) { wf_box_elem_nodes(b, e, o); });
  *out = pt;
  return 0;
}

extern "C" void wf_partition_free(wf_partition *p) { delete p; }
extern "C" int wf_partition_info(const wf_partition *p, int *eb, int *ee, int *nl, int *nn) {
  if (!p) return 1;
  if (eb) *eb = p->elem_begin;
  if (ee) *ee = p->elem_end;
  if (nl) *nl = (int)p->l2g.size();
  if (nn) *nn = (int)p->neigh.size();
  return 0;
}
extern "C" const int *wf_partition_node_l2g(const wf_partition *p) { return p->l2g.data(); }
extern "C" const unsigned *wf_partition_local_elnod(const wf_partition *p) { return p->elnod.data(); }
extern "C" const int *wf_partition_neigh_ranks(const wf_partition *p) { return p->neigh.data(); }
extern "C" const int *wf_partition_halo_offset(const wf_partition *p) { return p->halo_offset.data(); }
extern "C" const int *wf_partition_halo_nodes(const wf_partition *p) { return p->halo_nodes.data(); }

// accessors used by the engine
int wf_partition_k(const wf_partition *p) { return p->k; }
void wf_partition_ranks(const wf_partition *p, int *rank, int *nranks) { *rank = p->rank; *nranks = p->nranks; }
bool wf_partition_is_box(const wf_partition *p) { return p->is_box; }
void wf_partition_box_coords(const wf_partition *p, std::vector<double> &x) {
  std::vector<double> ax[3];
  wf_box_axes(p->box, p->V, p->r, ax);
  const int dim = p->box.dim;
  x.resize(p->l2g.size() * dim);
  for (size_t i = 0; i < p->l2g.size(); i++) wf_box_node_xyz(p->box, ax, p->l2g[i], x.data() + i * dim);
}
int wf_partition_box_dim(const wf_partition *p) { return p->box.dim; }

// ---- Domain_d::SearchExtNodes (src/common/Domain_d.C:110-205), integer part --------------------------------
// The reference appends every element face to faceList unless a face with the same node set is already there
// (then it counts it), matching by linear search.  What it ends with: every distinct node set once, ordered by
// first occurrence (ascending element, then local face), with its multiplicity; nodes of faces that occur once
// are the external nodes.  Same result here from one sort of the canonical (sorted-node) keys.
// Face tables: tetra_faces / quad_edges (Domain_d.h:172-193); the constructor wires the tetra table for 3D
// (Domain_d.h:246-249) and set2DFacesValues the quad edges for 2D (:785-791).
namespace {
const int kTetraFaces[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};
const int kQuadEdges[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
struct FaceKey {
  int n[3];
  int serial;
  bool same(const FaceKey &o) const { return n[0] == o.n[0] && n[1] == o.n[1] && n[2] == o.n[2]; }
};
}  // namespace

extern "C" int wf_host_ext_faces(int dim, int nodxelem, int n_nodes, int n_elems, const unsigned *elnod, unsigned char *ext_nodes,
                                 int *n_faces_total, int *n_ext_faces, int *ext_face_nodes, int *ext_face_elem) {
  int facenod;
  if (dim == 3 && nodxelem == 4) facenod = 3;
  else if (dim == 2 && nodxelem == 4) facenod = 2;
  else return 1;
  const size_t nf = (size_t)n_elems * 4;
  std::vector<FaceKey> keys(nf);
  for (int e = 0; e < n_elems; e++)
    for (int f = 0; f < 4; f++) {
      FaceKey &k = keys[(size_t)e * 4 + f];
      k.n[2] = -1;
      for (int q = 0; q < facenod; q++) {
        const unsigned g = elnod[(size_t)e * nodxelem + (facenod == 3 ? kTetraFaces[f][q] : kQuadEdges[f][q])];
        if (g >= (unsigned)n_nodes) return 2;
        k.n[q] = (int)g;
      }
      std::sort(k.n, k.n + facenod);
      k.serial = (int)((size_t)e * 4 + f);
    }
  std::sort(keys.begin(), keys.end(), [](const FaceKey &a, const FaceKey &b) {
    for (int q = 0; q < 3; q++)
      if (a.n[q] != b.n[q]) return a.n[q] < b.n[q];
    return a.serial < b.serial;
  });
  std::vector<int> mult(nf, 0); // multiplicity, stored at the serial of the first occurrence
  int total = 0;
  for (size_t i = 0; i < nf;) {
    size_t j = i + 1;
    while (j < nf && keys[j].same(keys[i])) j++;
    mult[keys[i].serial] = (int)(j - i);
    total++;
    i = j;
  }
  memset(ext_nodes, 0, (size_t)n_nodes);
  int nx = 0;
  for (size_t s = 0; s < nf; s++) {
    if (mult[s] != 1) continue;
    const int e = (int)(s / 4), f = (int)(s % 4);
    for (int q = 0; q < facenod; q++) {
      const int g = (int)elnod[(size_t)e * nodxelem + (facenod == 3 ? kTetraFaces[f][q] : kQuadEdges[f][q])];
      ext_face_nodes[(size_t)nx * facenod + q] = g;
      ext_nodes[g] = 1;
    }
    ext_face_elem[nx] = e;
    nx++;
  }
  if (n_faces_total) *n_faces_total = total;
  if (n_ext_faces) *n_ext_faces = nx;
  return 0;
}

// ---- TriMesh_d::AxisPlaneMesh (src/common/Mesh.C:48-283) ----------------------------------------------------
// (dens+1)^2 nodes / 2 dens^2 triangles (3D) or dens+1 nodes / dens segments (2D).  As in the reference the nodes
// start at p1 and advance by dl = (p2 - p1).x / dens along x (and y in 3D) whatever `axis` says — `axis` only
// selects the component of the initial normal (+1 if positaxisorent, else -1) — and the winding follows
// positaxisorent (:156-163, :184-190).
extern "C" int wf_host_axis_plane_counts(int dimension, int dens, int *n_nodes, int *n_elems) {
  if ((dimension != 2 && dimension != 3) || dens < 1) return 1;
  *n_nodes = dimension == 3 ? (dens + 1) * (dens + 1) : dens + 1;
  *n_elems = dimension == 3 ? dens * dens * 2 : dens;
  return 0;
}
extern "C" int wf_host_axis_plane_mesh(int dimension, int mesh_id, int axis, int positaxisorent, const double p1[3], const double p2[3],
                                       int dens, double *node, int *elnode, double *normal, int *ele_mesh_id) {
  int nn, ne;
  if (wf_host_axis_plane_counts(dimension, dens, &nn, &ne)) return 1;
  const double dl = (p2[0] - p1[0]) / dens;
  double x2 = p1[1];
  const double x3 = p1[2];
  int vi = 0;
  const int rows = dimension == 2 ? 1 : dens + 1;
  for (int j = 0; j < rows; j++) {
    double x1 = p1[0];
    for (int i = 0; i < dens + 1; i++) {
      node[3 * vi] = x1; node[3 * vi + 1] = x2; node[3 * vi + 2] = x3;
      vi++;
      x1 += dl;
    }
    x2 += dl;
  }
  int el = 0;
  if (dimension == 3) {
    for (int j = 0; j < dens; j++)
      for (int i = 0; i < dens; i++) {
        const int a = (dens + 1) * j + i, b = a + 1, c = (dens + 1) * (j + 1) + i, d = c + 1;
        const int t[2][2][3] = {{{a, c, b}, {b, c, d}}, {{a, b, c}, {b, d, c}}};
        for (int e = 0; e < 2; e++, el++)
          for (int q = 0; q < 3; q++) elnode[3 * el + q] = t[positaxisorent ? 1 : 0][e][q];
      }
  } else {
    for (int i = 0; i < dens; i++, el++) {
      elnode[2 * el] = positaxisorent ? i : i + 1;
      elnode[2 * el + 1] = positaxisorent ? i + 1 : i;
    }
  }
  const double f = positaxisorent ? 1. : -1.;
  for (int e = 0; e < ne; e++) {
    normal[3 * e] = normal[3 * e + 1] = normal[3 * e + 2] = 0.0;
    if (axis == 0) normal[3 * e] = f;
    else if (axis == 1) normal[3 * e + 1] = f;
    else if (dimension == 3) normal[3 * e + 2] = f;
    ele_mesh_id[e] = mesh_id;
  }
  return 0;
}
