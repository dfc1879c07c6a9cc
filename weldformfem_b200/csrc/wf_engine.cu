// wf_engine.cu — host side of the C ABI (include/wf_engine.h): owns device memory, mirrors the
// call sequence of Domain_d::SolveChungHulbert() (src/explicit/Solver_explicit.C) and converts
// between the reference's array layouts and the private device layouts (wf_dev.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wf_engine.h"
#include "wf_dev.h"
#include "wf_host.h"
#include "wf_launch.h"
#include "wf_math_ids.h"

static std::string g_create_error;

#include "wf_engine_priv.h"

static int check_launch(wf_engine *E, const char *what) { return wf_check_launch(E, what); }
int wf_null_engine(void) { g_create_error = "null engine handle (create the mesh first)"; return 1; }

static int select_flavour(wf_engine *E) {
  E->L = E->strict ? wf_strict_table() : wf_fast_table();
  E->P.strict = E->strict ? 1 : 0;
  return 0;
}

extern "C" const char *wf_version(void) { return "weldformfem_b200 0.1 (sm_100a, fp64)"; }

extern "C" const char *wf_last_error(wf_engine *E) { return E ? E->err.c_str() : g_create_error.c_str(); }

extern "C" int wf_create(wf_engine **out, int dim, int nodxelem, int domtype, int device) {
  if (!out) return 1;
  *out = nullptr;
  int et = -1;
  if (dim == 3 && nodxelem == 8) et = ET_HEX8;
  else if (dim == 3 && nodxelem == 4) et = ET_TET4;
  else if (dim == 2 && nodxelem == 4) et = ET_QUAD4;
  else if (dim == 2 && nodxelem == 3) et = ET_TRI3;
  if (et < 0) { g_create_error = "unsupported element: dim/nodxelem must be 3/8, 3/4, 2/4 or 2/3"; return 1; }
  if ((dim == 3) != (domtype == WF_3D)) { g_create_error = "domtype does not match dim"; return 1; }
  if (domtype == WF_PLANE_STRESS) { g_create_error = "plane stress is not implemented by the reference step either"; return 1; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                     " (this engine has no CPU fallback)";
    return 2;
  }
  if (device < 0 || device >= ndev) { g_create_error = "bad device ordinal"; return 1; }
  ce = cudaSetDevice(device);
  if (ce != cudaSuccess) { g_create_error = cudaGetErrorString(ce); return 1; }
  wf_engine *E = new wf_engine();
  E->device = device; E->dim = dim; E->k = nodxelem; E->et = et; E->domtype = domtype;
  memset(&E->d, 0, sizeof(E->d));
  memset(&E->P, 0, sizeof(E->P));
  memset(&E->mat, 0, sizeof(E->mat));
  memset(&E->stab, 0, sizeof(E->stab));
  memset(&E->C, 0, sizeof(E->C));
  E->stab.hg_stiff = 0.1; // Domain_d.h:294
  E->d.dim = dim; E->d.k = nodxelem; E->d.domtype = domtype;
  E->P.w = (et == ET_HEX8) ? 8.0 : (et == ET_TET4 ? 1.0 / 6.0 : (et == ET_QUAD4 ? 4.0 : 0.5));
  E->P.stab_simple = 1;
  E->P.hg_stiff = 0.1;
  if (const char *t = getenv("WF_ELEM_ORDER")) E->order_mode = atoi(t) != 0 ? 1 : 0;
  if (const char *t = getenv("WF_VARIANT")) sscanf(t, "%d,%d,%d,%d", &E->P.variant[0], &E->P.variant[1], &E->P.variant[2], &E->P.variant[3]); // tuning runs
  select_flavour(E);
  *out = E;
  return 0;
}

extern "C" void wf_destroy(wf_engine *E) {
  if (!E) return;
  cudaSetDevice(E->device);
  cudaStreamSynchronize(E->stream);
  for (int b = 0; b < 2; b++) {
    if (E->bc_stage[b]) cudaFreeHost(E->bc_stage[b]);
    if (E->bc_ev[b]) cudaEventDestroy(E->bc_ev[b]);
  }
  if (E->mon_host) cudaFreeHost(E->mon_host);
  for (int b = 0; b < 2; b++)
    if (E->mon_ev[b]) cudaEventDestroy(E->mon_ev[b]);
  if (E->copy_stream) { cudaStreamSynchronize(E->copy_stream); cudaStreamDestroy(E->copy_stream); }
  for (int b = 0; b < 2; b++)
    if (E->bc_at_ev[b]) cudaEventDestroy(E->bc_at_ev[b]);
  if (E->bc_copy_ev) cudaEventDestroy(E->bc_copy_ev);
  for (void *p : E->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void *p : E->allocs) cudaFree(p);
  if (E->scratch) cudaFree(E->scratch);
  if (E->own_stream) cudaStreamDestroy(E->stream);
  delete E;
}

extern "C" int wf_set_stream(wf_engine *E, void *s) { WF_NULLCHK(E);
  CK(cudaSetDevice(E->device));
  CK(cudaStreamSynchronize(E->stream));
  if (E->own_stream) { cudaStreamDestroy(E->stream); E->own_stream = false; }
  E->stream = (cudaStream_t)s;
  return 0;
}
extern "C" int wf_get_stream(wf_engine *E, void **s) { WF_NULLCHK(E); if (s) *s = (void *)E->stream; return 0; }
// multi-GPU: a halo wait that timed out leaves comm_error set (k_halo_wait); surfaced by every entry point that reads
static int check_halo_error(wf_engine *E) {
  if (!E->distributed || !E->d.comm_error) return 0;
  int e = 0;
  CK(cudaMemcpyAsync(&e, E->d.comm_error, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  if (e) FAIL("halo exchange timed out waiting for neighbour index " + std::to_string(e - 1) + " (shared-node state is stale)");
  return 0;
}
extern "C" int wf_synchronize(wf_engine *E) { WF_NULLCHK(E);
  CK(cudaSetDevice(E->device));
  CK(cudaStreamSynchronize(E->stream));
  return check_halo_error(E);
}

extern "C" int wf_set_axisymm_vol_weight(wf_engine *E, int on) { WF_NULLCHK(E);
  NEED(E->domtype == WF_AXISYMM, "vol_weight only applies to axisymmetric domains");
  E->d.vol_weight = on ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// mesh
// ---------------------------------------------------------------------------------------------------
static int upload_mesh(wf_engine *E, int nn, int ne, const double *x, const unsigned *elnod) {
  NEED(!E->meshed, "mesh already set");
  NEED(nn > 0 && ne > 0, "empty mesh");
  CK(cudaSetDevice(E->device));
  const int k = E->k, dim = E->dim;
  NEED((long long)ne * k <= 2147483647LL, "connectivity exceeds 32-bit slot range");
  E->nn = nn; E->ne = ne;
  WfDev &d = E->d;
  d.nn = nn; d.ne = ne;
  d.np = round_up(nn, 32); d.ep = round_up(ne, 32);
  d.nslices = (nn + 31) / 32;
  // node -> element lists exactly like setNodElem
  E->h_elnod.assign(elnod, elnod + (size_t)ne * k);
  E->h_offset.resize(nn); E->h_count.resize(nn);
  E->h_nodel.resize((size_t)ne * k); E->h_nodel_loc.resize((size_t)ne * k);
  if (wf_host_nodel(nn, ne, k, elnod, E->h_offset.data(), E->h_count.data(), E->h_nodel.data(), E->h_nodel_loc.data()))
    FAIL("connectivity entry out of range");
  // internal element order (wf_host_elem_order): the device arrays, the CTA / tile tables and the slot ids below use
  // it; the node->element LISTS keep the order of setNodElem (ascending user element id)
  E->perm.clear(); E->iperm.clear();
  std::vector<unsigned> el_int;
  const unsigned *eli = elnod; // connectivity in internal order, reference layout [e*k + ln]
  const int order = E->order_mode >= 0 ? E->order_mode : (k == 8 ? 1 : 0);
  std::vector<unsigned long long> okeys; // hexahedra: sort keys in internal order (wf_host_brick_plan)
  if (order != 0 && ne > 1) {
    E->perm.resize(ne); E->iperm.resize(ne);
    if (k == 8 && dim == 3) okeys.resize(ne);
    if (wf_host_elem_order_keys(dim, k, nn, ne, x, elnod, order, E->perm.data(), okeys.empty() ? nullptr : okeys.data()))
      FAIL("element ordering failed");
    bool ident = true;
    for (int e = 0; e < ne; e++) { E->iperm[E->perm[e]] = e; ident = ident && E->perm[e] == e; }
    if (ident) { E->perm.clear(); E->iperm.clear(); }
    else {
      el_int.resize((size_t)ne * k);
      for (int e = 0; e < ne; e++) memcpy(&el_int[(size_t)e * k], elnod + (size_t)E->perm[e] * k, sizeof(unsigned) * k);
      eli = el_int.data();
    }
  }
  const int *ip = E->iperm.empty() ? nullptr : E->iperm.data();
  auto internal = [&](int e_user) { return ip ? ip[e_user] : e_user; };
  // sliced-ELL packing of slot = e*k + ln (e = internal id)
  std::vector<long long> sell_ptr(d.nslices + 1);
  long long tot = 0;
  for (int s = 0; s < d.nslices; s++) {
    sell_ptr[s] = tot;
    int w = 0;
    for (int n = s * 32; n < std::min(nn, s * 32 + 32); n++) w = std::max(w, E->h_count[n]);
    tot += (long long)w * 32;
  }
  sell_ptr[d.nslices] = tot;
  std::vector<int> slots((size_t)tot, -1);
  for (int n = 0; n < nn; n++) {
    const long long base = sell_ptr[n >> 5];
    const int off = E->h_offset[n];
    for (int j = 0; j < E->h_count[n]; j++)
      slots[(size_t)(base + (long long)j * 32 + (n & 31))] = internal(E->h_nodel[off + j]) * k + E->h_nodel_loc[off + j];
  }
  // offset of (e, ln) in the node-ordered force buffer [slice][j][dim][32]: dim*q - (dim-1)*lane
  NEED((long long)dim * tot < 4294967295LL, "node-ordered force buffer exceeds 32-bit offsets");
  E->sell_total = tot;
  const long long ep_ = round_up(ne, 32);
  E->h_pos.assign((size_t)k * ep_, 0u);
  for (int n = 0; n < nn; n++) {
    const long long base = sell_ptr[n >> 5];
    const int off = E->h_offset[n];
    for (int j = 0; j < E->h_count[n]; j++) {
      const long long q = base + (long long)j * 32 + (n & 31);
      E->h_pos[(size_t)E->h_nodel_loc[off + j] * ep_ + E->h_nodel[off + j]] = (unsigned)(dim * q - (long long)(dim - 1) * (n & 31));
    }
  }
  long long *dptr; int *dslots; int *dpos;
  if (dalloc(E, &dptr, sell_ptr.size()) || dalloc(E, &dslots, slots.size()) || dalloc(E, &dpos, E->h_pos.size()) ||
      dalloc(E, &d.fsell, (size_t)dim * tot))
    return 1;
  if (ip) { // device copy indexed by internal id
    std::vector<unsigned> pos_int(E->h_pos.size(), 0u);
    for (int n = 0; n < k; n++)
      for (int e = 0; e < ne; e++) pos_int[(size_t)n * ep_ + e] = E->h_pos[(size_t)n * ep_ + E->perm[e]];
    CK(cudaMemcpyAsync(dpos, pos_int.data(), pos_int.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    int *dperm;
    if (dalloc(E, &dperm, (size_t)ep_) || dalloc(E, &E->iperm_d, (size_t)ep_)) return 1;
    CK(cudaMemcpyAsync(dperm, E->perm.data(), sizeof(int) * ne, cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(E->iperm_d, E->iperm.data(), sizeof(int) * ne, cudaMemcpyHostToDevice, E->stream));
    d.e_user = dperm;
  } else {
    CK(cudaMemcpyAsync(dpos, E->h_pos.data(), E->h_pos.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    d.e_user = nullptr; E->iperm_d = nullptr;
  }
  d.pos = dpos;
  CK(cudaMemcpyAsync(dptr, sell_ptr.data(), sell_ptr.size() * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
  CK(cudaMemcpyAsync(dslots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
  d.sell_ptr = dptr; d.sell_slots = dslots;
  // connectivity, SoA
  {
    std::vector<int> el((size_t)k * d.ep, 0);
    for (int e = 0; e < ne; e++)
      for (int n = 0; n < k; n++) el[(size_t)n * d.ep + e] = (int)eli[(size_t)e * k + n];
    int *del;
    if (dalloc(E, &del, el.size())) return 1;
    CK(cudaMemcpyAsync(del, el.data(), el.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    d.elnod = del;
    CK(cudaStreamSynchronize(E->stream));
  }
  // element-block node tables (see WfDev::blk_off)
  {
    const int nblk = (ne + WF_EBLK - 1) / WF_EBLK;
    std::vector<int> boff(nblk + 1, 0), bnodes;
    std::vector<unsigned short> lidx((size_t)k * d.ep, 0);
    bnodes.reserve((size_t)ne * 2);
    std::vector<int> tmp;
    int umax = 0;
    for (int b = 0; b < nblk; b++) {
      const int e0 = b * WF_EBLK, e1 = std::min(ne, e0 + WF_EBLK);
      tmp.assign(eli + (size_t)e0 * k, eli + (size_t)e1 * k);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      for (int e = e0; e < e1; e++)
        for (int n = 0; n < k; n++)
          lidx[(size_t)n * d.ep + e] =
              (unsigned short)(std::lower_bound(tmp.begin(), tmp.end(), (int)eli[(size_t)e * k + n]) - tmp.begin());
      bnodes.insert(bnodes.end(), tmp.begin(), tmp.end());
      boff[b + 1] = (int)bnodes.size();
      umax = std::max(umax, (int)tmp.size());
    }
    int *doff, *dnodes; unsigned short *dl;
    if (dalloc(E, &doff, boff.size()) || dalloc(E, &dnodes, bnodes.size()) || dalloc(E, &dl, lidx.size())) return 1;
    CK(cudaMemcpyAsync(doff, boff.data(), boff.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(dnodes, bnodes.data(), bnodes.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(dl, lidx.data(), lidx.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    d.blk_off = doff; d.blk_nodes = dnodes; d.lidx = dl; d.blk_umax = umax;
    {
      const int pitch = (umax + 31) / 32 * 32;
      d.blk_pitch = pitch;
      std::vector<int> pad((size_t)nblk * pitch, -1);
      for (int b = 0; b < nblk; b++) std::copy(bnodes.begin() + boff[b], bnodes.begin() + boff[b + 1], pad.begin() + (size_t)b * pitch);
      int *dpad;
      if (dalloc(E, &dpad, std::max<size_t>(pad.size(), 1))) return 1;
      CK(cudaMemcpyAsync(dpad, pad.data(), pad.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
      CK(cudaStreamSynchronize(E->stream));
      d.blk_pad = dpad;
    }
    // tile-reduced force path (WfDev::ftile): tables built on the host (wf_force_tiles_build, wf_mesh.cpp)
    d.ftile = nullptr; d.tf_ptr = nullptr; d.tf_slots = nullptr; d.tf_idx = nullptr; d.tf_tab = nullptr;
    d.lidx_pk = nullptr; d.tile_pk = nullptr; d.blk_pad_b = nullptr; d.brick_elem = nullptr; d.cta_lookahead = 0;
    d.sm_count = 0;
    CK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, E->device));
    d.n_bcta = 0; d.brick_plan = 0;
    d.tf_stride = 0; d.tf_tpitch = 0;
    if (dim == 3) { // 2D (1M quads, measured): rounds form E2 0.080 -> 0.099 ms, pull form 0.105 ms, N2 unchanged: not used
      WfForceTiles T;
      // brick form of the hexa passes (k_elem_vol_brick, k_elem_main_hex_brick): thread slots (wf_host_brick_plan, or the
      // compact numbering), bank-aware shared-memory slots (wf_host_run_slots) of every CTA's node list and of every
      // tile's node list, the slots of an element's eight nodes packed into one record each, and per tile the table
      // rank -> slot used when the partial sums are written out.  All built on the host first: the plan is dropped for
      // the compact numbering when its tables do not fit.
      constexpr int BS = WF_BRICK_STRIDE, BW = WF_BRICK_WS;
      std::vector<int> slot_elem, bpad, belem;
      std::vector<unsigned short> lpk;
      std::vector<unsigned> tpk;
      int n_bcta = 0;
      auto build_brick = [&](bool plan) -> bool {
        n_bcta = nblk;
        slot_elem.clear();
        if (plan) {
          if (wf_host_brick_plan(ne, okeys.data(), &n_bcta, nullptr)) return false;
          slot_elem.resize((size_t)n_bcta * WF_EBLK);
          if (wf_host_brick_plan(ne, okeys.data(), &n_bcta, slot_elem.data())) return false;
        }
        const long long ns = plan ? (long long)n_bcta * WF_EBLK : ne, sp = plan ? ns : d.ep;
        const int *se = plan ? slot_elem.data() : nullptr;
        auto elem = [&](long long s_) { return se ? se[s_] : (s_ < ne ? (int)s_ : -1); };
        wf_force_tiles_build(nn, (int)ns, se, k, dim, sp, eli, T);
        if (!T.usable) return false;
        if (k != 8 || !T.rounds) return !plan;
        const long long nslot = (long long)n_bcta * WF_EBLK;
        bpad.assign((size_t)n_bcta * BS, -1);
        lpk.assign((size_t)8 * nslot, 0);
        tpk.assign((size_t)4 * nslot, 0u);
        belem.assign((size_t)nslot, 0);
        std::vector<int> sl, ids;
        for (int b = 0; b < n_bcta; b++) {
          const long long s0 = (long long)b * WF_EBLK;
          ids.clear();
          int first = -1;
          for (int t = 0; t < WF_EBLK; t++) {
            const int e = elem(s0 + t);
            if (e < 0) continue;
            if (first < 0) first = t;
            ids.insert(ids.end(), eli + (size_t)e * 8, eli + (size_t)e * 8 + 8);
          }
          if (first < 0) return false;
          std::sort(ids.begin(), ids.end());
          ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
          const int U = (int)ids.size();
          if (U > BS) return false;
          sl.resize(U);
          // ragged CTAs (not a whole brick: more, shorter runs) keep the plain ascending layout
          if (wf_host_run_slots(U, ids.data(), sl.data()) > BS)
            for (int i = 0; i < U; i++) sl[i] = i;
          for (int i = 0; i < U; i++) bpad[(size_t)b * BS + sl[i]] = ids[i];
          for (int t = 0; t < WF_EBLK; t++) {
            const int e = elem(s0 + t), es = e >= 0 ? e : elem(s0 + first); // idle threads read the tables of a real element
            belem[(size_t)(s0 + t)] = e >= 0 ? e : ~es;
            for (int n = 0; n < 8; n++)
              lpk[(size_t)(s0 + t) * 8 + n] =
                  (unsigned short)sl[std::lower_bound(ids.begin(), ids.end(), (int)eli[(size_t)es * 8 + n]) - ids.begin()];
          }
        }
        for (long long w = 0; w < (long long)n_bcta * (WF_EBLK / 32); w++) {
          const long long s0 = w * 32;
          ids.clear();
          for (int t = 0; t < 32; t++)
            if (elem(s0 + t) >= 0) ids.insert(ids.end(), eli + (size_t)elem(s0 + t) * 8, eli + (size_t)elem(s0 + t) * 8 + 8);
          std::sort(ids.begin(), ids.end());
          ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
          const int U = (int)ids.size();
          if (U > BW || U > 128) return false;
          sl.resize(U);
          if (wf_host_run_slots(U, ids.data(), sl.data()) > BW)
            for (int i = 0; i < U; i++) sl[i] = i;
          unsigned r2s[32];
          for (int t = 0; t < 32; t++) r2s[t] = 0xFFFFFFFFu;
          for (int i = 0; i < U; i++) r2s[i & 31] = (r2s[i & 31] & ~(0xFFu << (8 * (i >> 5)))) | ((unsigned)sl[i] << (8 * (i >> 5)));
          for (int t = 0; t < 32; t++) {
            unsigned *rec = &tpk[(size_t)(s0 + t) * 4];
            if (elem(s0 + t) >= 0)
              for (int n = 0; n < 8; n++) rec[n >> 2] |= (unsigned)sl[T.tidx[(size_t)n * sp + (s0 + t)]] << (8 * (n & 3));
            rec[2] = r2s[t];
            rec[3] = (unsigned)belem[(size_t)(s0 + t)];
          }
        }
        return true;
      };
      bool plan = k == 8 && !okeys.empty();
      if (const char *t = getenv("WF_BRICK_PLAN")) plan = plan && atoi(t) != 0; // 0: compact thread slots (A/B runs, tests)
      bool brick = build_brick(plan);
      if (!brick && plan) { plan = false; brick = build_brick(false); }
      brick = brick && k == 8 && T.usable && T.rounds && !bpad.empty();
      if (T.usable) {
        long long *dtp; unsigned *dts;
        if (dalloc(E, &dtp, T.ptr.size()) || dalloc(E, &dts, std::max<size_t>(T.slots.size(), 1)) ||
            dalloc(E, &d.ftile, (size_t)T.n_tiles * dim * T.stride))
          return 1;
        CK(cudaMemcpyAsync(dtp, T.ptr.data(), T.ptr.size() * sizeof(long long), cudaMemcpyHostToDevice, E->stream));
        CK(cudaMemcpyAsync(dts, T.slots.data(), T.slots.size() * sizeof(unsigned), cudaMemcpyHostToDevice, E->stream));
        if (k == 8) {
          if (!plan) { // the generic tile kernel addresses tiles by the compact numbering
            unsigned char *dti;
            if (dalloc(E, &dti, T.tidx.size())) return 1;
            CK(cudaMemcpyAsync(dti, T.tidx.data(), T.tidx.size(), cudaMemcpyHostToDevice, E->stream));
            d.tf_idx = dti;
          }
          if (brick) {
            uint4 *dl4, *dt4; int *dbp, *dbe;
            if (dalloc(E, &dl4, belem.size()) || dalloc(E, &dt4, belem.size()) || dalloc(E, &dbp, bpad.size()) || dalloc(E, &dbe, belem.size()))
              return 1;
            CK(cudaMemcpyAsync(dl4, lpk.data(), lpk.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, E->stream));
            CK(cudaMemcpyAsync(dt4, tpk.data(), tpk.size() * sizeof(unsigned), cudaMemcpyHostToDevice, E->stream));
            CK(cudaMemcpyAsync(dbp, bpad.data(), bpad.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
            CK(cudaMemcpyAsync(dbe, belem.data(), belem.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
            CK(cudaStreamSynchronize(E->stream));
            d.lidx_pk = dl4; d.tile_pk = dt4; d.blk_pad_b = dbp; d.brick_elem = dbe;
            d.n_bcta = n_bcta; d.brick_plan = plan ? 1 : 0;
            int sms = 0;
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, E->device));
            d.cta_lookahead = 4 * sms; // four 128-register CTAs per SM
          }
        } else {
          unsigned char *dtab;
          if (dalloc(E, &dtab, T.tab.size())) return 1;
          CK(cudaMemcpyAsync(dtab, T.tab.data(), T.tab.size(), cudaMemcpyHostToDevice, E->stream));
          d.tf_tab = dtab; d.tf_tpitch = T.tpitch;
        }
        CK(cudaStreamSynchronize(E->stream));
        d.tf_ptr = dtp; d.tf_slots = dts; d.tf_stride = T.stride;
      }
    }
  }
  // state
  const size_t nv = (size_t)dim * d.np, e6 = (size_t)6 * d.ep;
  if (dalloc(E, &d.x, nv) || dalloc(E, &d.v, nv) || dalloc(E, &d.prev_a, nv) || dalloc(E, &d.u, nv) ||
      dalloc(E, &d.u_dt, nv) || dalloc(E, &d.voln_sum, d.np) || dalloc(E, &d.voln0_sum, d.np) ||
      dalloc(E, &d.nodal_p, d.np) || dalloc(E, &d.rhobar, d.np) || dalloc(E, &d.nodel_count, d.np) || dalloc(E, &d.bc_index, d.np) ||
      dalloc(E, &d.tau, e6) || dalloc(E, &d.p, d.ep) || dalloc(E, &d.pl_strain, d.ep) ||
      dalloc(E, &d.sigma_y, d.ep) || dalloc(E, &d.vol, d.ep) || dalloc(E, &d.vol_0, d.ep) ||
      dalloc(E, &d.rho, d.ep) || dalloc(E, &d.rho_0, d.ep) ||
      dalloc(E, &d.nonfinite, 1) || dalloc(E, &d.xmin_key, 2) || dalloc(E, &d.red, 8) || dalloc(E, &d.mdiag, d.np))
    return 1;
  if (E->et == ET_QUAD4 && dalloc(E, &d.hg_q, (size_t)2 * d.ep)) return 1;
  {
    std::vector<double> xs(nv, 0.0);
    for (int n = 0; n < nn; n++)
      for (int c = 0; c < dim; c++) xs[(size_t)c * d.np + n] = x[(size_t)n * dim + c];
    CK(cudaMemcpyAsync(d.x, xs.data(), nv * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    std::vector<int> cnt(d.np, 1);
    std::copy(E->h_count.begin(), E->h_count.end(), cnt.begin());
    CK(cudaMemcpyAsync(d.nodel_count, cnt.data(), d.np * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    std::vector<int> bci(d.np, -1);
    CK(cudaMemcpyAsync(d.bc_index, bci.data(), d.np * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
  }
  E->meshed = true;
  if (E->material_set) {
    std::vector<double> r(d.ep, E->mat.rho0);
    CK(cudaMemcpy(d.rho_0, r.data(), d.ep * sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}

extern "C" int wf_set_mesh(wf_engine *E, int nn, int ne, const double *x, const unsigned *elnod) { WF_NULLCHK(E);
  NEED(x && elnod, "null mesh arrays");
  return upload_mesh(E, nn, ne, x, elnod);
}

extern "C" int wf_gen_box(wf_engine *E, const double V[3], const double L[3], double r, int tritet) { WF_NULLCHK(E);
  WfBox b;
  wf_box_dims(L, r, tritet, &b);
  NEED(b.dim == E->dim && b.k == E->k, "box element type does not match the engine's dim/nodxelem");
  NEED(b.nn > 0 && b.ne > 0, "box has no elements");
  NEED(b.nn <= 2147483647LL && b.ne * b.k <= 2147483647LL, "box exceeds 32-bit index range");
  std::vector<double> x((size_t)b.nn * b.dim);
  std::vector<unsigned> el((size_t)b.ne * b.k);
  wf_host_gen_box(V, L, r, tritet, x.data(), el.data());
  return upload_mesh(E, (int)b.nn, (int)b.ne, x.data(), el.data());
}

extern "C" int wf_set_elem_order(wf_engine *E, int mode) { WF_NULLCHK(E);
  NEED(!E->meshed, "wf_set_elem_order before the mesh is set");
  NEED(mode == 0 || mode == 1, "element order mode must be 0 (caller's numbering) or 1 (Morton)");
  E->order_mode = mode;
  return 0;
}

extern "C" int wf_set_axis_xmin(wf_engine *E, double xmin) { WF_NULLCHK(E);
  NEED(!E->meshed, "wf_set_axis_xmin before the mesh is set");
  E->axis_xmin = xmin; E->axis_xmin_set = true;
  return 0;
}

extern "C" int wf_brick_info(wf_engine *E, int *n_cta, int *plan) { WF_NULLCHK(E);
  NEED(E->meshed, "no mesh");
  if (n_cta) *n_cta = E->d.n_bcta;
  if (plan) *plan = E->d.brick_plan;
  return 0;
}

extern "C" int wf_get_counts(wf_engine *E, int *nn, int *ne, int *ntot) { WF_NULLCHK(E);
  NEED(E->meshed, "no mesh");
  if (nn) *nn = E->nn;
  if (ne) *ne = E->ne;
  if (ntot) *ntot = E->ne * E->k;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// material / options / BCs
// ---------------------------------------------------------------------------------------------------
extern "C" int wf_set_material(wf_engine *E, const wf_material *m) { WF_NULLCHK(E);
  NEED(m, "null material");
  NEED(m->model >= WF_BILINEAR && m->model <= WF_GMT, "material model must be Bilinear, Hollomon, JohnsonCook or GMT");
  E->mat = *m;
  WfPar &P = E->P;
  P.model = m->model;
  P.Kbulk = m->E / (3.0 * (1.0 - 2.0 * m->nu)); // Elastic_, Material.cuh:24-28
  P.G = m->E / (2.0 * (1.0 + m->nu));
  P.sy0 = m->sy0;
  if (m->model == WF_HOLLOMON) { // InitHollomon, Material.cuh:90-104
    P.Kh = m->K; P.mh = m->m;
    P.eps0 = m->sy0 / m->E;
    P.eps1 = pow(m->sy0 / m->K, 1. / m->m);
  }
  P.cs0 = sqrt(P.Kbulk / m->rho0); // main.C:574
  P.young = m->E;
  for (int i = 0; i < 14; i++) P.mq[i] = m->q[i];
  P.temp = m->temp;
  P.max_edot = m->max_edot > 0.0 ? m->max_edot : 1.0e6; // Domain_d.h:824
  E->material_set = true;
  if (E->meshed) {
    CK(cudaSetDevice(E->device));
    std::vector<double> r(E->d.ep, m->rho0);
    CK(cudaMemcpy(E->d.rho_0, r.data(), E->d.ep * sizeof(double), cudaMemcpyHostToDevice));
  }
  return 0;
}

static void refresh_stab_simple(wf_engine *E) {
  const wf_stab &s = E->stab;
  E->P.stab_simple = (s.alpha_free == 0.0 && s.alpha_contact == 0.0 && s.hg_coeff_contact == 0.0 && s.hg_coeff_free == 0.0 && s.av_coeff_div == 0.0 && s.av_coeff_bulk == 0.0 &&
                      s.log_factor == 0.0 && s.pspg_scale == 0.0 && s.p_pspg_bulkfac == 0.0) ? 1 : 0;
}

extern "C" int wf_set_stab(wf_engine *E, const wf_stab *s) { WF_NULLCHK(E);
  NEED(s, "null stab");
  E->stab = *s;
  WfPar &P = E->P;
  P.alpha_free = s->alpha_free; P.hg_coeff_free = s->hg_coeff_free; P.av_coeff_div = s->av_coeff_div;
  P.av_coeff_bulk = s->av_coeff_bulk; P.log_factor = s->log_factor; P.pspg_scale = s->pspg_scale;
  P.p_pspg_bulkfac = s->p_pspg_bulkfac; P.J_min = s->J_min; P.hg_visc = s->hg_visc; P.hg_stiff = s->hg_stiff;
  P.hexa_hg = s->hexa_hg_coeff;
  P.alpha_contact = s->alpha_contact; P.hg_coeff_contact = s->hg_coeff_contact;
  refresh_stab_simple(E);
  return 0;
}

extern "C" int wf_set_options(wf_engine *E, int press, double av_alpha, double av_beta, int strict) { WF_NULLCHK(E);
  NEED(press == WF_PRESS_DEFAULT || press == WF_PRESS_ANP_SHIPPED || press == WF_PRESS_ANP_NODAL, "bad pressure algorithm");
  NEED(!E->inited, "options must be set before wf_init");
  E->P.press = press; E->P.av_alpha = av_alpha; E->P.av_beta = av_beta;
  E->strict = strict != 0;
  select_flavour(E);
  return 0;
}

extern "C" int wf_set_tracking(wf_engine *E, int flags) { WF_NULLCHK(E);
  NEED(!E->inited, "tracking must be set before wf_init");
  E->tracking = flags;
  return 0;
}

// thermal coupling: setThermalOn + setTemp(T0) + thermalCond / thermalHeatCap / thermalExp + plHeatFrac
// (main.C:218, 436-441, 567-570; Thermal.C)
extern "C" int wf_set_thermal(wf_engine *E, double k_T, double cp_T, double exp_T, double plheatfrac, double T0) { WF_NULLCHK(E);
  NEED(E->meshed, "wf_set_thermal needs the mesh");
  NEED(!E->inited, "thermal coupling must be switched on before wf_init");
  NEED(!E->distributed, "thermal coupling is not available on a partitioned mesh");
  CK(cudaSetDevice(E->device));
  WfDev &d = E->d;
  if (!d.T && (dalloc(E, &d.T, (size_t)d.np) || dalloc(E, &d.tsell, (size_t)E->sell_total) || dalloc(E, &d.q_plheat, (size_t)d.ep) ||
               dalloc(E, &d.dtedt_low[0], (size_t)d.np) || dalloc(E, &d.dtedt_low[1], (size_t)d.np)))
    return 1;
  std::vector<double> t0(d.np, T0);
  CK(cudaMemcpyAsync(d.T, t0.data(), sizeof(double) * d.np, cudaMemcpyHostToDevice, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  E->P.thermal = 1; E->P.dtedt_cur = 0;
  E->P.k_T = k_T; E->P.cp_T = cp_T; E->P.exp_T = exp_T; E->P.plheatfrac = plheatfrac;
  return 0;
}

extern "C" int wf_add_bc_vel(wf_engine *E, int node, int dim, double val) { WF_NULLCHK(E);
  NEED(dim >= 0 && dim < 3, "bad BC dim");
  E->bc_nod[dim].push_back(node);
  E->bc_val[dim].push_back(val);
  E->bcs_ready = false;
  return 0;
}
extern "C" int wf_add_bc_vel_array(wf_engine *E, int count, const int *node, const int *dim, const double *val) { WF_NULLCHK(E);
  for (int i = 0; i < count; i++)
    if (wf_add_bc_vel(E, node[i], dim[i], val[i])) return 1;
  return 0;
}

extern "C" int wf_allocate_bcs(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->meshed, "AllocateBCs needs the mesh");
  CK(cudaSetDevice(E->device));
  // per node: mask of prescribed dims + values; a later AddBCVelNode on the same (node, dim) wins, which is
  // what the serial scatter of ImposeBCV (Domain_d.C:1109-1121) produces
  std::map<int, int> row_of;
  std::vector<unsigned char> mask;
  std::vector<double> vals;
  for (int dd = 0; dd < E->dim; dd++)
    for (size_t i = 0; i < E->bc_nod[dd].size(); i++) {
      int n = E->bc_nod[dd][i];
      if (i == 0) E->bc_slot[dd].assign(E->bc_nod[dd].size(), -1);
      if (E->distributed) { // ids are GLOBAL node ids; nodes of other ranks are skipped
        auto it = std::lower_bound(E->l2g.begin(), E->l2g.end(), n);
        if (it == E->l2g.end() || *it != n) continue;
        n = (int)(it - E->l2g.begin());
      }
      NEED(n >= 0 && n < E->nn, "BC node out of range");
      auto it = row_of.find(n);
      int row;
      if (it == row_of.end()) {
        row = (int)mask.size();
        row_of[n] = row;
        mask.push_back(0);
        vals.insert(vals.end(), 3, 0.0);
      } else row = it->second;
      mask[row] |= (unsigned char)(1u << dd);
      vals[3 * (size_t)row + dd] = E->bc_val[dd][i];
      E->bc_slot[dd][i] = 3 * row + dd;
    }
  std::vector<int> bci(E->d.np, -1);
  for (auto &kv : row_of) bci[kv.first] = kv.second;
  unsigned char *dm; double *dv;
  if (dalloc(E, &dm, mask.size()) || dalloc(E, &dv, vals.size())) return 1;
  if (!mask.empty()) {
    CK(cudaMemcpyAsync(dm, mask.data(), mask.size(), cudaMemcpyHostToDevice, E->stream));
    CK(cudaMemcpyAsync(dv, vals.data(), vals.size() * sizeof(double), cudaMemcpyHostToDevice, E->stream));
  }
  CK(cudaMemcpyAsync(E->d.bc_index, bci.data(), bci.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  E->d.bc_mask = dm; E->d.bc_vals = dv;
  E->bc_vals_d = dv;
  E->nbc_rows = (int)mask.size();
  E->bc_vals_buf[0] = dv; E->bc_vals_buf[1] = nullptr; E->bc_vals_cur = 0; E->bc_set_calls = 0;
  {
    std::vector<int> row_node(std::max<size_t>(mask.size(), 1), 0);
    for (auto &kv : row_of) row_node[kv.second] = kv.first;
    if (dalloc(E, &E->bc_row_node_d, row_node.size())) return 1;
    CK(cudaMemcpy(E->bc_row_node_d, row_node.data(), row_node.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  for (int b = 0; b < 2; b++) {
    if (E->bc_stage[b]) { cudaFreeHost(E->bc_stage[b]); E->bc_stage[b] = nullptr; }
    if (!vals.empty()) {
      CK(cudaMallocHost((void **)&E->bc_stage[b], vals.size() * sizeof(double)));
      memcpy(E->bc_stage[b], vals.data(), vals.size() * sizeof(double));
    }
    if (!E->bc_ev[b]) CK(cudaEventCreateWithFlags(&E->bc_ev[b], cudaEventDisableTiming));
    for (int dd = 0; dd < 3; dd++) E->bc_stage_version[b][dd] = 0;
  }
  for (int dd = 0; dd < 3; dd++) E->bc_version[dd] = 0;
  E->bcs_ready = true;
  return 0;
}

// New prescribed values for the BCs of one dimension, in insertion order — the engine-side equivalent of writing
// into Domain_d::bcx_val / bcy_val / bcz_val (Domain_d.h:901) between steps (time-dependent velocity BCs).
// Host values are staged in pinned memory and uploaded asynchronously on the engine's stream.
extern "C" int wf_set_bc_values(wf_engine *E, int dim, int count, const double *vals) { WF_NULLCHK(E);
  NEED(E->bcs_ready, "wf_set_bc_values needs wf_allocate_bcs");
  NEED(dim >= 0 && dim < E->dim, "bad BC dim");
  NEED(count == (int)E->bc_slot[dim].size(), "count must equal the number of BCs of this dimension");
  if (E->nbc_rows == 0 || count == 0) return 0;
  CK(cudaSetDevice(E->device));
  for (int i = 0; i < count; i++) E->bc_val[dim][i] = vals[i];
  E->bc_version[dim]++;
  const int b = E->bc_stage_cur;
  CK(cudaEventSynchronize(E->bc_ev[b])); // the copy that last used this staging buffer has finished
  double *st = E->bc_stage[b];
  for (int dd = 0; dd < E->dim; dd++) {   // bring this staging buffer up to date (only dimensions that changed)
    if (E->bc_stage_version[b][dd] == E->bc_version[dd]) continue;
    const std::vector<int> &slot = E->bc_slot[dd];
    const std::vector<double> &v = E->bc_val[dd];
    for (size_t i = 0; i < slot.size(); i++)
      if (slot[i] >= 0) st[slot[i]] = v[i];
    E->bc_stage_version[b][dd] = E->bc_version[dd];
  }
  // Upload on a copy stream into the device copy the running step does NOT read (two copies, used alternately), so the
  // transfer of step k+1's values overlaps step k.  Copy j targets the buffer last read by step j-2: it waits for the
  // event recorded on the engine's stream at the previous call (covers every step enqueued before it).
  const size_t nbytes = (size_t)3 * E->nbc_rows * sizeof(double);
  if (!E->copy_stream) {
    CK(cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking));
    for (int q = 0; q < 2; q++) CK(cudaEventCreateWithFlags(&E->bc_at_ev[q], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&E->bc_copy_ev, cudaEventDisableTiming));
  }
  if (!E->bc_vals_buf[1] && dalloc(E, &E->bc_vals_buf[1], (size_t)3 * E->nbc_rows)) return 1;
  const long j = E->bc_set_calls++;
  const int tgt = E->bc_vals_cur ^ 1;
  if (j > 0) CK(cudaStreamWaitEvent(E->copy_stream, E->bc_at_ev[(j - 1) & 1], 0));
  else CK(cudaStreamSynchronize(E->stream)); // first call: the second copy was just allocated (memset on the engine's stream)
  CK(cudaEventRecord(E->bc_at_ev[j & 1], E->stream));
  CK(cudaMemcpyAsync(E->bc_vals_buf[tgt], st, nbytes, cudaMemcpyHostToDevice, E->copy_stream));
  CK(cudaEventRecord(E->bc_ev[b], E->copy_stream));
  CK(cudaEventRecord(E->bc_copy_ev, E->copy_stream));
  CK(cudaStreamWaitEvent(E->stream, E->bc_copy_ev, 0));
  E->bc_vals_cur = tgt;
  E->bc_vals_d = E->bc_vals_buf[tgt];
  E->d.bc_vals = E->bc_vals_d;
  E->bc_stage_cur ^= 1;
  // engine in predicted state (wf_step_open): the velocities of the prescribed components already hold the values of
  // the previous predictor; replace them (ImposeBCV after UpdatePrediction, Solver_explicit.C:535-540)
  if (E->predicted) {
    E->L->bc_patch_v(E->d, E->bc_row_node_d, E->nbc_rows, E->stream);
    return check_launch(E, "wf_set_bc_values");
  }
  return 0;
}

// Step monitor without draining the stream: wf_monitor_async enqueues the kinetic-energy reduction
// (computeEnergies, Mechanical.C:2145) and a copy of {Ekin, non-finite flag (Solver_explicit.C:779), halo error}
// into pinned host memory; wf_monitor_wait returns the OLDEST pending result.  Up to two may be pending, so a
// host loop can read step i's monitor while step i+1 is already running.
extern "C" int wf_monitor_async(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->inited, "wf_monitor_async before wf_init");
  NEED(E->mon_pending < 2, "two monitors already pending: call wf_monitor_wait");
  CK(cudaSetDevice(E->device));
  const int b = (E->mon_head + E->mon_pending) & 1;
  WfDev d2 = E->d;
  d2.ne = 0;                       // kinetic part only
  d2.red = E->mon_red + wf_engine::MON_NACC * b;
  const bool fused = E->ekin_step == E->step_count && E->ekin_slot == b; // formed by the last node pass already
  NEED(fused || !E->predicted, "no kinetic energy was formed for this step (two monitors were pending during wf_step_open)");
  if (!fused) {
    CK(cudaMemsetAsync(d2.red, 0, wf_engine::MON_NACC * sizeof(double), E->stream));
    E->L->energy(d2, nullptr, E->stream);   // adds into d2.red[0]
  }
  E->mon_seen = true;
  E->ekin_step = -1;
  wf_engine::MonSlot *h = E->mon_host + b;
  CK(cudaMemcpyAsync(h->part, d2.red, wf_engine::MON_NACC * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaMemcpyAsync(&h->nonfinite, E->d.nonfinite, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
  if (E->distributed) CK(cudaMemcpyAsync(&h->halo_error, E->d.comm_error, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
  else h->halo_error = 0;
  CK(cudaEventRecord(E->mon_ev[b], E->stream));
  E->mon_pending++;
  return check_launch(E, "wf_monitor_async");
}

extern "C" int wf_monitor_wait(wf_engine *E, double *Ekin, int *nonfinite) { WF_NULLCHK(E);
  NEED(E->mon_pending > 0, "no monitor pending");
  CK(cudaSetDevice(E->device));
  const int b = E->mon_head;
  CK(cudaEventSynchronize(E->mon_ev[b]));
  const wf_engine::MonSlot &h = E->mon_host[b];
  E->mon_head ^= 1;
  E->mon_pending--;
  double ek = 0.0;
  for (int i = 0; i < wf_engine::MON_NACC; i++) ek += h.part[i];
  if (Ekin) *Ekin = ek;
  if (nonfinite) *nonfinite = h.nonfinite;
  if (h.halo_error) FAIL("halo exchange timed out waiting for neighbour index " + std::to_string(h.halo_error - 1));
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// optional arrays
// ---------------------------------------------------------------------------------------------------
static int ensure_dbg(wf_engine *E) {
  if (E->dbg) return 0;
  WfDev &d = E->d;
  const size_t nv = (size_t)E->dim * d.np, e6 = (size_t)6 * d.ep, ekd = (size_t)E->k * E->dim * d.ep;
  if (dalloc(E, &d.dH, ekd) || dalloc(E, &d.detJ, d.ep) || dalloc(E, &d.radius, d.ep) || dalloc(E, &d.str_rate, e6) ||
      dalloc(E, &d.rot_rate, e6) || dalloc(E, &d.a, nv) || dalloc(E, &d.fi, nv) || dalloc(E, &d.voln, d.np))
    return 1;
  if (!d.sigma && dalloc(E, &d.sigma, e6)) return 1;
  if (!d.f_elem && dalloc(E, &d.f_elem, ekd)) return 1;
  if (!d.f_elem_hg && dalloc(E, &d.f_elem_hg, ekd)) return 1;
  E->dbg = true;
  return 0;
}

int wf_check_launch(wf_engine *E, const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { E->err = std::string(what) + ": " + cudaGetErrorString(e); return 1; }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// init + step
// ---------------------------------------------------------------------------------------------------
static int reset_xmin(wf_engine *E, int slot) {
  // ordered key of 1000.0 (Solver_explicit.C:956 `double xmin = 1000.0`)
  double v = 1000.0;
  unsigned long long b;
  memcpy(&b, &v, 8);
  unsigned long long key = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  CK(cudaMemcpyAsync(E->d.xmin_key + slot, &key, 8, cudaMemcpyHostToDevice, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  return 0;
}

// ---- stages ---------------------------------------------------------------------------------------
// Both wf_init and wf_step are sequences of stages; between some of them a distributed engine exchanges
// partial nodal sums with its neighbours (halo_exchange).  With the peer-memory transport the exchange is
// part of the stream (stores into the neighbour's receive region + flag, then a one-CTA wait kernel), so
// wf_init / wf_step enqueue everything without touching the host; with the host-driven transport the
// caller moves the staging regions (NCCL send/recv) between wf_init_phase / wf_step_phase calls.
static void halo_send(wf_engine *E, int mode) {
  E->seq++;
  E->P.halo_parity = (int)(E->seq & 1ull);
  E->L->halo_send(E->d, E->P, mode, E->strict ? 1 : 0, E->seq, E->max_halo_count, E->stream);
}
static void halo_wait(wf_engine *E) {
  if (E->transport == 0) E->L->halo_wait(E->d, E->seq, E->timeout_ns, E->stream);
}
// Peer transport inside the step: no launches of their own.  The send of an exchange rides in the first CTAs of the
// next node pass (WfPar::send_ctas), the wait at the head of the kernel that consumes the neighbours' partials
// (WfPar::wait_seq).  WF_HALO_FOLD=0 keeps the separate kernels (measurement).
static bool halo_folded(const wf_engine *E) {
  static const bool on = [] { const char *t = getenv("WF_HALO_FOLD"); return !(t && atoi(t) == 0); }();
  return on && E->distributed && E->transport == 0;
}
static void halo_send_arm(wf_engine *E) {   // the NEXT node-pass launch sends
  E->seq++;
  E->P.halo_parity = (int)(E->seq & 1ull);
  E->P.send_chunks = (std::max(E->max_halo_count, 1) + 255) / 256;   // 256 = threads per CTA of the node passes
  E->P.send_ctas = (int)E->neigh.size() * E->P.send_chunks;
  E->P.send_seq = E->seq;
}
static void halo_wait_arm(wf_engine *E) {   // the NEXT consumer launch waits
  E->P.wait_seq = E->seq;
  E->P.wait_timeout_ns = E->timeout_ns;
}
static void halo_disarm(wf_engine *E) { E->P.send_ctas = 0; E->P.wait_seq = 0; }

static int init_stage(wf_engine *E, int stage, double dt) {
  WfDev &d = E->d;
  WfPar &P = E->P;
  const size_t nv = (size_t)E->dim * d.np;
  if (stage == 0) {
    P.dt = dt;
    const double rho_b = 0.818200; // Solver_explicit.C:193-197
    P.alpha = (2.0 * rho_b - 1.0) / (1.0 + rho_b);
    P.beta = (5.0 - 3.0 * rho_b) / ((1.0 + rho_b) * (1.0 + rho_b) * (2.0 - rho_b));
    P.gamma = 1.5 - P.alpha;
    E->L->init_elem(d, P, E->stream); // InitValues
    // Solver_explicit.C:176-190: v, a, u are zeroed inside the per-dimension loop, so only the LAST
    // dimension's prescribed velocities survive initialisation
    CK(cudaMemsetAsync(d.v, 0, nv * sizeof(double), E->stream));
    CK(cudaMemsetAsync(d.u, 0, nv * sizeof(double), E->stream));
    if (d.a) CK(cudaMemsetAsync(d.a, 0, nv * sizeof(double), E->stream));
    E->L->impose_bc(d, E->dim - 1, 0, d.v, E->stream);
    E->L->elem_vol(d, P, E->et, 0, E->stream);      // calcElemJAndDerivatives + CalcElemInitialVol/CalcElemVol
    E->L->vol0_density(d, E->stream);               // vol_0 = vol ; calcElemDensity
    E->L->node_vol(d, P, 0, E->stream);             // sum vol_0 / mean rho per node
    if (E->distributed) halo_send(E, 0);
  } else if (stage == 1) {
    if (E->distributed) E->L->halo_finish(d, P, 0, P.halo_parity, E->stream);
    E->L->node_vol(d, P, 1, E->stream);             // CalcNodalVol
    if (E->distributed) halo_send(E, 1);
  } else {
    if (E->distributed) E->L->halo_finish(d, P, 1, P.halo_parity, E->stream);
    if (E->domtype == WF_AXISYMM) {
      P.xmin_cur = 0;
      if (reset_xmin(E, 0)) return 1;
      E->L->xmin(d, 0, E->stream);
    }
    if (wf_contact_init(E)) return 1;
    E->time = 0.0; E->step_count = 0; E->predicted = false;
    E->a_in_dbg = E->fi_in_dbg = E->sigma_in_dbg = E->rates_in_dbg = E->felem_in_dbg = false;
    E->inited = true;
  }
  return check_launch(E, "wf_init");
}

static int init_prologue(wf_engine *E) {
  NEED(E->meshed && E->material_set, "wf_init needs mesh and material");
  if (!E->bcs_ready && wf_allocate_bcs(E)) return 1;
  CK(cudaSetDevice(E->device));
  { // allocations and kernel loading happen here, before any halo wait can be in flight
    WfDev &d = E->d;
    WfPar &P = E->P;
    P.track_eps = (E->tracking & 1) ? 1 : 0;
    P.store_sigma = ((E->tracking & 2) || P.av_alpha != 0.0 || P.av_beta != 0.0) ? 1 : 0;
    const size_t e6 = (size_t)6 * d.ep;
    if (P.track_eps && !d.eps && dalloc(E, &d.eps, e6)) return 1;
    if (P.store_sigma && !d.sigma && dalloc(E, &d.sigma, e6)) return 1;
    if (E->strict && !d.fsell_hg && dalloc(E, &d.fsell_hg, (size_t)E->dim * E->sell_total)) return 1;
  }
  if (!E->mon_host) { // pinned result ring of wf_monitor_async (page pinning can take milliseconds: do it here)
    CK(cudaMallocHost((void **)&E->mon_host, 2 * sizeof(wf_engine::MonSlot)));
    if (dalloc(E, &E->mon_red, 2 * wf_engine::MON_NACC)) return 1;
    for (int b = 0; b < 2; b++) CK(cudaEventCreateWithFlags(&E->mon_ev[b], cudaEventDisableTiming));
  }
  E->L->preload(E->et, E->dim, E->k);
  CK(cudaStreamSynchronize(E->stream));
  return check_launch(E, "kernel preload");
}

extern "C" int wf_init(wf_engine *E, double dt) { WF_NULLCHK(E);
  if (init_prologue(E)) return 1;
  if (E->distributed) {
    NEED(E->transport == 0, "host-driven halo transport: initialise with wf_init_phase");
    NEED(E->n_connected == (int)E->neigh.size(), "wf_init: connect every neighbour first (wf_halo_connect / wf_connect_all)");
  }
  for (int st = 0; st < 3; st++) {
    if (init_stage(E, st, dt)) return 1;
    if (E->distributed && st < 2) halo_wait(E);
  }
  return 0;
}

extern "C" int wf_init_phase(wf_engine *E, int phase, double dt) { WF_NULLCHK(E);
  NEED(phase >= 0 && phase < 3, "init phase must be 0, 1 or 2");
  NEED(phase == E->init_stage, "init phases must be called in order 0, 1, 2");
  if (phase == 0 && init_prologue(E)) return 1;
  CK(cudaSetDevice(E->device));
  if (init_stage(E, phase, dt)) return 1;
  E->init_stage = (phase + 1) % 3;
  return 0;
}

// flags of the node pass: bit 0 = also run the next step's predictor (every step of a batch but the last);
// WF_FAST only: bit 1 = this step's u_dt was not stored by the previous (fused) step and is recomputed from v and prev_a,
// bit 2 = do not store u_dt because the next step recomputes it (saves 48 B per node and step)
static int fuse_flags(const wf_engine *E, bool last) {
  if (E->open_mode) last = false; // wf_step_open: the call's last node pass runs the next predictor too
  int f = last ? 0 : 1;
  if (!E->strict) {
    // only a predictor fused into the PREVIOUS node pass may drop u_dt: the first predictor of a batch must store it,
    // because prescribed velocities can change between batches (u_dt of a constrained component is dt * OLD value)
    if (E->predicted) f |= 2;
    if (!last) f |= 4;
  }
  return f;
}

// wf_step_timed: an event after the launch(es) just enqueued; tags: 0 predictor, 1 E1, 2 N1 (+ folded halo send), 3 E2,
// 4 N2 (single GPU: the whole node pass; partitioned: the nodes this rank does not share, + folded send), 5 nodal sums of
// the shared nodes (+ folded wait), 6 N2 of the shared nodes (+ folded wait), 7 halo kernels of their own
static inline void tmark(wf_engine *E, int tag) {
  if (!E->tm_ev) return;
  cudaEvent_t x;
  cudaEventCreate(&x);
  cudaEventRecord(x, E->stream);
  E->tm_ev->push_back(x);
  E->tm_tag->push_back(tag);
}

// one explicit step = stages 0..2 (Solver_explicit.C:524-978, rows 1-22)
static int step_stage(wf_engine *E, int stage, bool last) {
  WfDev &d = E->d;
  WfPar &P = E->P;
  const int sep = E->strict ? 1 : 0;
  if (stage == 0) {
    if (wf_contact_step_begin(E)) return 1;      // CalcExtFaceAreas every 10th step (Solver_explicit.C:445-450)
    // the batch's first predictor: its own kernel (strict) or folded into the nodal-sum pass (fast)
    const bool fold = !E->predicted && !E->strict;
    if (!E->predicted && E->strict) { E->L->predict(d, P, 1, E->stream); tmark(E, 0); }
    E->L->elem_vol(d, P, E->et, 0, E->stream);
    tmark(E, 1);
    // the partial volume sums of the shared nodes depend on E1 only: they are sent BEFORE the nodal sums are formed
    // (first CTAs of the N1 launch, or a kernel of their own), so the transfer and the neighbours' flags travel
    // while N1 runs
    const bool fold_halo = halo_folded(E);
    if (fold_halo) halo_send_arm(E);
    else if (E->distributed) { halo_send(E, 1); tmark(E, 7); }
    E->L->node_vol(d, P, fold ? 3 : 1, E->stream);
    tmark(E, 2);
    halo_disarm(E);
  } else if (stage == 1) {
    const bool fold_halo = halo_folded(E);
    if (fold_halo) halo_wait_arm(E);
    if (E->distributed) { E->L->halo_finish(d, P, 1, P.halo_parity, E->stream); tmark(E, 5); }
    halo_disarm(E);
    E->L->elem_main(d, P, E->et, sep, E->stream);
    tmark(E, 3);
    if (E->distributed && !fold_halo) { halo_send(E, 2); tmark(E, 7); }
  } else {
    if (wf_contact_forces(E)) return 1;          // CalcContactForces (Solver_explicit.C:769-770)
    // (wf_step_open always forms it: in predicted state the corrected velocities are gone afterwards)
    const bool fuse_ekin = last && (E->mon_seen || E->open_mode) && E->mon_red && E->mon_pending < 2;
    if (fuse_ekin) { // the node pass of the call's last step also forms the kinetic energy for wf_monitor_async
      const int b = (E->mon_head + E->mon_pending) & 1;
      d.ekin_acc = E->mon_red + wf_engine::MON_NACC * b;
      CK(cudaMemsetAsync(d.ekin_acc, 0, wf_engine::MON_NACC * sizeof(double), E->stream));
      E->ekin_step = E->step_count + 1;
      E->ekin_slot = b;
    }
    if (E->distributed && E->transport == 0) {
      // peer transport: the nodes this rank does not share are integrated while the neighbours' force partials are
      // still travelling; the shared ones follow after the wait (step_once skips its own wait before this stage)
      if (halo_folded(E)) {
        halo_send_arm(E);                        // first CTAs of the phase-3 launch send the force partials
        E->L->node_update(d, P, sep, fuse_flags(E, last), 3, E->stream);
        tmark(E, 4);
        halo_disarm(E);
        halo_wait_arm(E);                        // phase 4 waits for the neighbours itself
        E->L->node_update(d, P, sep, fuse_flags(E, last), 4, E->stream);
        tmark(E, 6);
        halo_disarm(E);
      } else {
        E->L->node_update(d, P, sep, fuse_flags(E, last), 3, E->stream);
        tmark(E, 4);
        halo_wait(E);
        tmark(E, 7);
        E->L->node_update(d, P, sep, fuse_flags(E, last), 4, E->stream);
        tmark(E, 6);
      }
    } else {
      E->L->node_update(d, P, sep, fuse_flags(E, last), 0, E->stream);
    }
    d.ekin_acc = nullptr;
    if (wf_contact_step_end(E)) return 1;        // rigid surfaces: ramp, Move, normals, plane coefficients (:981-1005)
    if (P.thermal) { E->L->node_thermal(d, P, E->stream); P.dtedt_cur ^= 1; }  // ThermalCalcs, node part (:1008-1012)
    if (!(E->distributed && E->transport == 0)) tmark(E, 4);
    E->predicted = !last || E->open_mode;
    E->udt_valid = !E->predicted;
    P.xmin_cur ^= 1;
    E->time += P.dt;
    E->step_count++;
  }
  return 0;
}

static int step_once(wf_engine *E, bool last) {
  for (int st = 0; st < 3; st++) {
    if (step_stage(E, st, last)) return 1;
    // stage 2 waits for the forces itself; with the folded exchange the consumers wait themselves
    if (E->distributed && st < 2 && !(st == 1 && E->transport == 0) && !halo_folded(E)) { halo_wait(E); tmark(E, 7); }
  }
  return 0;
}

static int step_prologue(wf_engine *E) {
  NEED(E->inited, "wf_step before wf_init");
  if (E->distributed) NEED(E->transport == 0, "host-driven halo transport: step with wf_step_phase");
  CK(cudaSetDevice(E->device));
  return 0;
}
static int step_epilogue(wf_engine *E, const char *what) {
  E->a_in_dbg = E->fi_in_dbg = E->sigma_in_dbg = E->rates_in_dbg = E->felem_in_dbg = false;
  return check_launch(E, what);
}

extern "C" int wf_step(wf_engine *E, int nsteps) { WF_NULLCHK(E);
  if (step_prologue(E)) return 1;
  for (int s = 0; s < nsteps; s++)
    if (step_once(E, s == nsteps - 1)) return 1;
  return step_epilogue(E, "wf_step");
}

// wf_step that leaves the engine in PREDICTED state: the last node pass of the call also runs the next step's
// UpdatePrediction + ImposeBCV (as every other step of a batch does), so a host loop that steps once per call — new
// prescribed velocities in, a monitor value out, every step — runs the same fused schedule as one long batch.
// Between calls only wf_set_bc_values (which patches the predicted velocities of the prescribed components),
// wf_monitor_async / wf_monitor_wait, wf_step and wf_step_open are allowed; wf_step_close (or a plain wf_step) returns
// to the state every other entry point expects.  Fast flavour only.
extern "C" int wf_step_open(wf_engine *E, int nsteps) { WF_NULLCHK(E);
  NEED(!E->strict, "wf_step_open needs the fast flavour (the strict flavour runs the reference's unfused predictor)");
  if (step_prologue(E)) return 1;
  E->open_mode = true;
  int rc = 0;
  for (int s = 0; s < nsteps && !rc; s++) rc = step_once(E, s == nsteps - 1);
  E->open_mode = false;
  if (rc) return 1;
  return step_epilogue(E, "wf_step_open");
}

// leave the predicted state without stepping: v_c = v_p - (1 - gamma) dt a.  "u_dt" (the last increment, which the
// fused schedule does not store) is not available until the next closed step.
extern "C" int wf_step_close(wf_engine *E) { WF_NULLCHK(E);
  NEED(E->inited, "wf_step_close before wf_init");
  if (!E->predicted) return 0;
  CK(cudaSetDevice(E->device));
  E->L->unpredict(E->d, E->P, E->stream);
  E->predicted = false;
  E->udt_valid = false;
  return check_launch(E, "wf_step_close");
}

extern "C" int wf_step_phase(wf_engine *E, int phase, int last_step) { WF_NULLCHK(E);
  NEED(E->inited, "wf_step_phase before wf_init");
  NEED(phase >= 0 && phase < 3, "step phase must be 0, 1 or 2");
  NEED(phase == E->step_stage, "step phases must be called in order 0, 1, 2");
  CK(cudaSetDevice(E->device));
  if (step_stage(E, phase, last_step != 0)) return 1;
  E->step_stage = (phase + 1) % 3;
  return step_epilogue(E, "wf_step_phase");
}

// tuning / profiling hooks -----------------------------------------------------------------------------
extern "C" int wf_set_variant(wf_engine *E, int kernel, int variant) { WF_NULLCHK(E);
  NEED(kernel >= 0 && kernel < 4, "kernel id must be 0..3 (E1, N1, E2, N2)");
  E->P.variant[kernel] = variant;
  return 0;
}

// wf_step with a CUDA event after every launch; ms[0..7] += predictor, E1, N1, E2, N2, and on a partitioned mesh: nodal
// sums of the shared nodes, N2 of the shared nodes, halo kernels of their own (see tmark; folded sends / waits are part
// of the launch that carries them, so a neighbour's lateness shows up in slots 5 and 6).  Peer transport or one GPU.
extern "C" int wf_step_timed(wf_engine *E, int nsteps, float *ms) { WF_NULLCHK(E);
  if (step_prologue(E)) return 1;
  NEED(ms, "null output");
  std::vector<cudaEvent_t> ev;
  std::vector<int> tag;
  E->tm_ev = &ev; E->tm_tag = &tag;
  tmark(E, -1);
  int rc = 0;
  for (int s = 0; s < nsteps && !rc; s++) rc = step_once(E, s == nsteps - 1);
  E->tm_ev = nullptr; E->tm_tag = nullptr;
  if (!rc && cudaStreamSynchronize(E->stream) != cudaSuccess) rc = 1;
  for (size_t i = 1; i < ev.size() && !rc; i++) {
    float t = 0.f;
    cudaEventElapsedTime(&t, ev[i - 1], ev[i]);
    ms[tag[i]] += t;
  }
  for (auto x : ev) cudaEventDestroy(x);
  if (rc) { if (E->err.empty()) E->err = "wf_step_timed failed"; return 1; }
  return step_epilogue(E, "wf_step_timed");
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU: local part of a partitioned mesh + halo plumbing
// ---------------------------------------------------------------------------------------------------
extern "C" int wf_set_mesh_partition(wf_engine *E, const wf_partition *p, const double *x_local) { WF_NULLCHK(E);
  NEED(p, "null partition");
  NEED(wf_partition_k(p) == E->k, "partition nodxelem does not match the engine");
  NEED(E->domtype != WF_AXISYMM || E->axis_xmin_set,
       "axisymmetric partition: call wf_set_axis_xmin with the minimum radial coordinate of the whole mesh first");
  int eb = 0, ee = 0, nl = 0, nng = 0;
  wf_partition_info(p, &eb, &ee, &nl, &nng);
  NEED(ee > eb && nl > 0, "this rank owns no elements");
  std::vector<double> xb;
  if (!x_local) {
    NEED(wf_partition_is_box(p), "x_local is required unless the partition was built by wf_partition_build_box");
    NEED(wf_partition_box_dim(p) == E->dim, "box dimension does not match the engine");
    wf_partition_box_coords(p, xb);
    x_local = xb.data();
  }
  if (E->domtype == WF_AXISYMM) {
    // the axis constraint uses the rank-local minimum of x_r: it must be the global one (see wf_set_axis_xmin)
    double lmin = x_local[0];
    for (int n = 1; n < nl; n++) lmin = std::min(lmin, x_local[(size_t)n * E->dim]);
    NEED(lmin <= E->axis_xmin + 1.e-6, "axisymmetric partition: this rank owns no node on the axis (x_r = global minimum); "
                                       "the axis constraint cannot be evaluated locally");
  }
  CK(cudaSetDevice(E->device));
  if (E->stream == 0) { // engines of one process must not serialise on the legacy default stream
    CK(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    E->own_stream = true;
  }
  if (upload_mesh(E, nl, ee - eb, x_local, wf_partition_local_elnod(p))) return 1;
  WfDev &d = E->d;
  E->distributed = true;
  wf_partition_ranks(p, &E->rank, &E->nranks);
  E->l2g.assign(wf_partition_node_l2g(p), wf_partition_node_l2g(p) + nl);
  E->neigh.assign(wf_partition_neigh_ranks(p), wf_partition_neigh_ranks(p) + nng);
  E->halo_offset.assign(wf_partition_halo_offset(p), wf_partition_halo_offset(p) + nng + 1);
  const int nh = E->halo_offset[nng];
  E->halo_nodes_h.assign(wf_partition_halo_nodes(p), wf_partition_halo_nodes(p) + nh);
  if (const char *t = getenv("WF_HALO_TIMEOUT_S")) {
    double sec = atof(t);
    if (sec > 0.0) E->timeout_ns = (unsigned long long)(sec * 1e9);
  }
  // unique shared nodes, each with its sharers in ascending rank order (this rank included)
  std::vector<int> uniq(E->halo_nodes_h);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  const int nu = (int)uniq.size();
  std::vector<int> slot(d.np, -1), ptr(nu + 1, 0);
  for (int u = 0; u < nu; u++) slot[uniq[u]] = u;
  struct Sharer { int rank; int2 e; };
  std::vector<std::vector<Sharer>> ent(nu);
  for (int u = 0; u < nu; u++) ent[u].push_back({E->rank, make_int2(-1, 0)});
  E->max_halo_count = 0;
  for (int i = 0; i < nng; i++) {
    const int off = E->halo_offset[i], cnt = E->halo_offset[i + 1] - off;
    E->max_halo_count = std::max(E->max_halo_count, cnt);
    for (int j = 0; j < cnt; j++)
      ent[slot[E->halo_nodes_h[off + j]]].push_back({E->neigh[i], make_int2(2 * WF_HALO_NC * off + j, cnt)});
  }
  std::vector<int2> flat;
  for (int u = 0; u < nu; u++) {
    std::sort(ent[u].begin(), ent[u].end(), [](const Sharer &a, const Sharer &b) { return a.rank < b.rank; });
    ptr[u] = (int)flat.size();
    for (auto &q : ent[u]) flat.push_back(q.e);
  }
  ptr[nu] = (int)flat.size();
  // device copies
  int *d_halo = nullptr, *d_slot = nullptr, *d_unode = nullptr, *d_ptr = nullptr;
  int2 *d_ent = nullptr;
  if (dalloc(E, &d_halo, (size_t)nh) || dalloc(E, &d_slot, (size_t)d.np) || dalloc(E, &d_unode, (size_t)nu) ||
      dalloc(E, &d_ptr, (size_t)nu + 1) || dalloc(E, &d_ent, flat.size()) || dalloc(E, &E->counters_d, (size_t)nng) ||
      dalloc(E, &E->nb_d, (size_t)nng) || dalloc(E, &d.comm_error, 1))
    return 1;
  E->flag_bytes = (size_t)round_up(8LL * std::max(nng, 1), 256);
  E->comm_bytes = E->flag_bytes + sizeof(double) * 2 * WF_HALO_NC * (size_t)nh;
  if (dalloc(E, &E->comm, E->comm_bytes)) return 1;
  auto up = [&](void *dst, const void *src, size_t bytes) {
    return bytes == 0 ? cudaSuccess : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, E->stream);
  };
  CK(up(d_halo, E->halo_nodes_h.data(), sizeof(int) * nh));
  CK(up(d_slot, slot.data(), sizeof(int) * d.np));
  CK(up(d_unode, uniq.data(), sizeof(int) * nu));
  CK(up(d_ptr, ptr.data(), sizeof(int) * (nu + 1)));
  CK(up(d_ent, flat.data(), sizeof(int2) * flat.size()));
  CK(cudaStreamSynchronize(E->stream));
  d.halo_nodes = d_halo; d.halo_slot = d_slot; d.hu_node = d_unode; d.hu_ptr = d_ptr; d.hu_ent = d_ent;
  d.n_halo = nh; d.n_uniq = nu; d.n_neigh = nng;
  d.flags = (unsigned long long *)E->comm;
  d.recv = (const double *)(E->comm + E->flag_bytes);
  d.nb = E->nb_d;
  E->nb_h.assign(nng, WfHaloNb());
  for (int i = 0; i < nng; i++) {
    E->nb_h[i].offset = E->halo_offset[i];
    E->nb_h[i].count = E->halo_offset[i + 1] - E->halo_offset[i];
    E->nb_h[i].dst = nullptr; E->nb_h[i].flag = nullptr;
    E->nb_h[i].counter = E->counters_d + i;
  }
  E->n_connected = 0;
  return 0;
}

extern "C" int wf_halo_info(wf_engine *E, int *rank, int *nranks, int *n_neigh, const int **neigh_ranks, const int **halo_offset,
                            const int **node_l2g) { WF_NULLCHK(E);
  NEED(E->distributed, "not a partitioned engine");
  if (rank) *rank = E->rank;
  if (nranks) *nranks = E->nranks;
  if (n_neigh) *n_neigh = (int)E->neigh.size();
  if (neigh_ranks) *neigh_ranks = E->neigh.data();
  if (halo_offset) *halo_offset = E->halo_offset.data();
  if (node_l2g) *node_l2g = E->l2g.data();
  return 0;
}

extern "C" int wf_halo_comm_block(wf_engine *E, void **base, size_t *bytes) { WF_NULLCHK(E);
  NEED(E->distributed, "not a partitioned engine");
  if (base) *base = E->comm;
  if (bytes) *bytes = E->comm_bytes;
  return 0;
}

extern "C" int wf_halo_slot_offsets(wf_engine *E, int i, size_t *flag_off, size_t *region_off, size_t *region_bytes) { WF_NULLCHK(E);
  NEED(E->distributed, "not a partitioned engine");
  NEED(i >= 0 && i < (int)E->neigh.size(), "bad neighbour index");
  if (flag_off) *flag_off = 8 * (size_t)i;
  if (region_off) *region_off = E->flag_bytes + sizeof(double) * 2 * WF_HALO_NC * (size_t)E->halo_offset[i];
  if (region_bytes) *region_bytes = sizeof(double) * 2 * WF_HALO_NC * (size_t)(E->halo_offset[i + 1] - E->halo_offset[i]);
  return 0;
}

static int push_nb(wf_engine *E) {
  CK(cudaSetDevice(E->device));
  if (!E->nb_h.empty())
    CK(cudaMemcpy(E->nb_d, E->nb_h.data(), sizeof(WfHaloNb) * E->nb_h.size(), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int wf_halo_connect(wf_engine *E, int i, void *peer_base, size_t flag_off, size_t region_off) { WF_NULLCHK(E);
  NEED(E->distributed, "not a partitioned engine");
  NEED(E->transport == 0, "wf_halo_connect belongs to the peer-memory transport");
  NEED(i >= 0 && i < (int)E->neigh.size(), "bad neighbour index");
  NEED(peer_base, "null peer pointer");
  if (!E->nb_h[i].dst) E->n_connected++;
  E->nb_h[i].dst = (double *)((char *)peer_base + region_off);
  E->nb_h[i].flag = (unsigned long long *)((char *)peer_base + flag_off);
  return push_nb(E);
}

extern "C" int wf_halo_set_transport(wf_engine *E, int host_driven) { WF_NULLCHK(E);
  NEED(E->distributed, "not a partitioned engine");
  NEED(!E->inited, "choose the halo transport before wf_init");
  E->transport = host_driven ? 1 : 0;
  if (E->transport == 1) { // sends land in a local staging block with the layout of the comm block
    CK(cudaSetDevice(E->device));
    if (!E->staging && dalloc(E, &E->staging, E->comm_bytes)) return 1;
    for (size_t i = 0; i < E->nb_h.size(); i++) {
      E->nb_h[i].dst = (double *)(E->staging + E->flag_bytes + sizeof(double) * 2 * WF_HALO_NC * (size_t)E->halo_offset[i]);
      E->nb_h[i].flag = (unsigned long long *)(E->staging + 8 * i);
    }
    E->n_connected = (int)E->nb_h.size();
    return push_nb(E);
  }
  return 0;
}

// host-driven transport: what to send to / receive from neighbour i for the exchange just packed
extern "C" int wf_halo_exchange_ptrs(wf_engine *E, int i, void **send_ptr, void **recv_ptr, size_t *n_doubles) { WF_NULLCHK(E);
  NEED(E->distributed && E->transport == 1, "host-driven halo transport is not selected");
  NEED(i >= 0 && i < (int)E->neigh.size(), "bad neighbour index");
  const size_t cnt = (size_t)(E->halo_offset[i + 1] - E->halo_offset[i]);
  const size_t off = E->flag_bytes + sizeof(double) * (2 * WF_HALO_NC * (size_t)E->halo_offset[i] + (E->seq & 1ull) * WF_HALO_NC * cnt);
  if (send_ptr) *send_ptr = E->staging + off;
  if (recv_ptr) *recv_ptr = E->comm + off;
  if (n_doubles) *n_doubles = WF_HALO_NC * cnt;
  return 0;
}

extern "C" int wf_halo_ipc_export(wf_engine *E, void *handle64) { WF_NULLCHK(E);
  NEED(E->distributed && E->comm, "not a partitioned engine");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CK(cudaSetDevice(E->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, E->comm));
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int wf_halo_ipc_open(wf_engine *E, const void *handle64, void **mapped) { WF_NULLCHK(E);
  NEED(handle64 && mapped, "null argument");
  CK(cudaSetDevice(E->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  E->ipc_opened.push_back(p);
  *mapped = p;
  return 0;
}

extern "C" int wf_halo_status(wf_engine *E, int *error) { WF_NULLCHK(E);
  int e = 0;
  if (E->distributed) {
    CK(cudaSetDevice(E->device));
    CK(cudaMemcpyAsync(&e, E->d.comm_error, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
  }
  if (error) *error = e;
  if (e) FAIL("halo exchange timed out waiting for neighbour index " + std::to_string(e - 1));
  return 0;
}

// ---- all ranks of a job inside ONE process (one host thread drives every GPU) -------------------------
extern "C" int wf_connect_all(wf_engine **R, int n) {
  for (int p = 0; p < n; p++) {
    wf_engine *E = R[p];
    NEED(E->distributed && E->rank == p && E->nranks == n, "wf_connect_all: engines must be given in rank order");
    for (size_t i = 0; i < E->neigh.size(); i++) {
      wf_engine *Q = R[E->neigh[i]];
      int me = -1;
      for (size_t j = 0; j < Q->neigh.size(); j++)
        if (Q->neigh[j] == p) me = (int)j;
      NEED(me >= 0, "wf_connect_all: halo lists are not symmetric");
      if (Q->device != E->device) {
        CK(cudaSetDevice(E->device));
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, E->device, Q->device));
        NEED(can, "no peer access between the GPUs of two neighbouring ranks");
        cudaError_t ce = cudaDeviceEnablePeerAccess(Q->device, 0);
        if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) CK(ce);
        cudaGetLastError();
      }
      size_t fo, ro;
      if (wf_halo_slot_offsets(Q, me, &fo, &ro, nullptr)) { E->err = Q->err; return 1; }
      if (wf_halo_connect(E, (int)i, Q->comm, fo, ro)) return 1;
    }
  }
  return 0;
}

extern "C" int wf_init_all(wf_engine **R, int n, double dt) {
  for (int p = 0; p < n; p++)
    if (init_prologue(R[p])) return 1;
  for (int st = 0; st < 3; st++)
    for (int p = 0; p < n; p++) {
      wf_engine *E = R[p];
      CK(cudaSetDevice(E->device));
      if (init_stage(E, st, dt)) return 1;
      if (st < 2) halo_wait(E);
    }
  return 0;
}

extern "C" int wf_step_all(wf_engine **R, int n, int nsteps) {
  for (int p = 0; p < n; p++)
    if (step_prologue(R[p])) return 1;
  // interleave the ranks step by step so that no stream's launch queue fills up while it waits for a
  // neighbour whose work has not been enqueued yet
  for (int s = 0; s < nsteps; s++)
    for (int p = 0; p < n; p++) {
      wf_engine *E = R[p];
      CK(cudaSetDevice(E->device));
      if (step_once(E, s == nsteps - 1)) return 1;
    }
  for (int p = 0; p < n; p++)
    if (step_epilogue(R[p], "wf_step_all")) return 1;
  return 0;
}

extern "C" int wf_nonfinite_flag(wf_engine *E, int *flag) { WF_NULLCHK(E);
  CK(cudaSetDevice(E->device));
  int f = 0;
  CK(cudaMemcpyAsync(&f, E->d.nonfinite, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  if (f) CK(cudaMemsetAsync(E->d.nonfinite, 0, sizeof(int), E->stream));
  if (flag) *flag = f;
  return check_halo_error(E);
}

extern "C" int wf_set_time(wf_engine *E, double t, long steps) { WF_NULLCHK(E);
  NEED(E->inited, "wf_set_time after wf_init");
  NEED(!E->predicted, "engine is mid-batch");
  NEED(t >= 0.0 && steps >= 0, "negative time or step count");
  E->time = t;
  E->step_count = steps;
  return 0;
}

extern "C" int wf_get_time(wf_engine *E, double *t, long *steps) { WF_NULLCHK(E);
  if (t) *t = E->time;
  if (steps) *steps = E->step_count;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// unfused entry points (names = Domain_d members)
// ---------------------------------------------------------------------------------------------------
#define UNFUSED_PROLOGUE()                                        \
  NEED(E->inited, "call wf_init first");                          \
  NEED(!E->predicted, "engine is mid-batch");                     \
  NEED(!E->P.thermal, "the unfused entry points do not cover thermal coupling: step with wf_step"); \
  E->ekin_step = -1;                                              \
  CK(cudaSetDevice(E->device));                                   \
  if (ensure_dbg(E)) return 1;                                    \
  WfDev &d = E->d; WfPar &P = E->P; (void)d; (void)P;

extern "C" int wf_UpdatePrediction(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->predict(d, P, 0, E->stream); return check_launch(E, __func__); }
extern "C" int wf_ImposeBCV(wf_engine *E, int dd) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); NEED(dd >= 0 && dd < E->dim, "bad dim"); E->L->impose_bc(d, dd, 0, d.v, E->stream); return check_launch(E, __func__); }
extern "C" int wf_ImposeBCA(wf_engine *E, int dd) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); NEED(dd >= 0 && dd < E->dim, "bad dim"); E->L->impose_bc(d, dd, 1, d.a, E->stream); return check_launch(E, __func__); }
extern "C" int wf_calcElemJAndDerivatives(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->elem_vol(d, P, E->et, 1, E->stream); return check_launch(E, __func__); }
extern "C" int wf_Calc_Element_Radius(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->elem_vol(d, P, E->et, 2, E->stream); return check_launch(E, __func__); }
extern "C" int wf_CalcElemVol(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->vol_from_detj(d, E->et, E->stream); return check_launch(E, __func__); }
extern "C" int wf_CalcNodalVol(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_nodal_vol(d, E->stream); return check_launch(E, __func__); }
extern "C" int wf_CalcNodalMassFromVol(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->node_mass(d, P, 1, E->stream); return check_launch(E, __func__); }
extern "C" int wf_calcElemStrainRates(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_strain_rates(d, P, E->et, E->stream); E->rates_in_dbg = true; return check_launch(E, __func__); }
extern "C" int wf_calcElemPressure(wf_engine *E) { WF_NULLCHK(E);
  UNFUSED_PROLOGUE();
  E->L->node_vol(d, P, 1, E->stream);
  E->L->u_pressure(d, P, E->et, E->stream);
  return check_launch(E, __func__);
}
extern "C" int wf_CalcStressStrain(wf_engine *E, double dt) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_stress(d, P, dt, E->stream); E->sigma_in_dbg = true; return check_launch(E, __func__); }
extern "C" int wf_calcArtificialViscosity(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_artvisc(d, P, E->stream); return check_launch(E, __func__); }
extern "C" int wf_calcElemForces(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_forces(d, P, E->et, E->stream); E->felem_in_dbg = true; return check_launch(E, __func__); }
extern "C" int wf_calcElemHourglassForces(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_hourglass(d, P, E->et, E->stream); return check_launch(E, __func__); }
extern "C" int wf_assemblyForces(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_assembly(d, E->stream); E->fi_in_dbg = true; return check_launch(E, __func__); }
extern "C" int wf_calcAccel(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_accel(d, E->stream); E->a_in_dbg = true; return check_launch(E, __func__); }
extern "C" int wf_UpdateCorrectionAccVel(wf_engine *E) { WF_NULLCHK(E); UNFUSED_PROLOGUE(); E->L->u_corr_accvel(d, P, E->stream); return check_launch(E, __func__); }
extern "C" int wf_AxisConstraint(wf_engine *E) { WF_NULLCHK(E);
  UNFUSED_PROLOGUE();
  if (E->domtype != WF_AXISYMM) return 0;
  if (reset_xmin(E, P.xmin_cur)) return 1;
  E->L->xmin(d, P.xmin_cur, E->stream);
  E->L->u_axis(d, P, E->stream);
  return check_launch(E, __func__);
}
extern "C" int wf_UpdateCorrectionPos(wf_engine *E) { WF_NULLCHK(E);
  UNFUSED_PROLOGUE();
  E->L->u_corr_pos(d, P, E->stream);
  E->time += P.dt; E->step_count++;
  if (E->domtype == WF_AXISYMM) { // keep the fused path's running minimum consistent
    if (reset_xmin(E, P.xmin_cur)) return 1;
    E->L->xmin(d, P.xmin_cur, E->stream);
  }
  return check_launch(E, __func__);
}

// ---------------------------------------------------------------------------------------------------
// state access
// ---------------------------------------------------------------------------------------------------
enum Kind { K_NODEVEC, K_NODESCAL, K_ELEMSCAL, K_ELEM6, K_ELEMNODE, K_ELEMNODEVEC, K_HGQ, K_INT_HOST, K_RAW_DEV };

struct ArrayRef {
  Kind kind;
  double *dev = nullptr;
  const void *host = nullptr;
  size_t bytes = 0;
  int comp = 0; // for dH: which dimension
  bool lazy_dtedt = false;
  bool lazy_sigma = false, lazy_fi = false, lazy_voln = false, lazy_pnode = false, lazy_felem = false, hg = false;
};

static bool lookup(wf_engine *E, const std::string &nm, ArrayRef &r, bool for_write) {
  WfDev &d = E->d;
  const size_t nd = sizeof(double) * (size_t)E->nn * E->dim, nnb = sizeof(double) * (size_t)E->nn;
  const size_t neb = sizeof(double) * (size_t)E->ne, nk = neb * E->k;
  auto nodevec = [&](double *p) { r.kind = K_NODEVEC; r.dev = p; r.bytes = nd; return p != nullptr; };
  auto elems = [&](double *p) { r.kind = K_ELEMSCAL; r.dev = p; r.bytes = neb; return p != nullptr; };
  auto elem6 = [&](double *p) { r.kind = K_ELEM6; r.dev = p; r.bytes = 6 * neb; return p != nullptr; };
  if (nm == "x") return nodevec(d.x);
  if (nm == "v") return nodevec(d.v);
  if (nm == "u") return nodevec(d.u);
  if (nm == "u_dt") return E->udt_valid || for_write ? nodevec(d.u_dt) : false;
  if (nm == "prev_a") return nodevec(d.prev_a);
  if (nm == "a") return nodevec((E->a_in_dbg || for_write) && d.a ? d.a : d.prev_a);
  if (nm == "m_fe") return nodevec(d.fe);
  if (nm == "m_fi") {
    if (E->fi_in_dbg && d.fi) return nodevec(d.fi);
    r.kind = K_NODEVEC; r.bytes = nd; r.lazy_fi = true; return !for_write;
  }
  if (nm == "m_mdiag") { r.kind = K_NODESCAL; r.dev = d.mdiag; r.bytes = nnb; return true; }
  if (nm == "m_voln") { r.kind = K_NODESCAL; r.dev = d.voln_sum; r.bytes = nnb; r.lazy_voln = true; return !for_write; }
  if (nm == "p_node") { r.kind = K_NODESCAL; r.bytes = nnb; r.lazy_pnode = true; return !for_write; }
  if (nm == "vol") return elems(d.vol);
  if (nm == "vol_0") return elems(d.vol_0);
  if (nm == "rho") return elems(d.rho);
  if (nm == "rho_0") return elems(d.rho_0);
  if (nm == "p") return elems(d.p);
  if (nm == "pl_strain") return elems(d.pl_strain);
  if (nm == "sigma_y") return elems(d.sigma_y);
  if (nm == "m_detJ") return elems(d.detJ);
  if (nm == "m_radius") return elems(d.radius);
  if (nm == "m_elem_length") return E->elem_length_valid ? elems(E->elem_length) : false;
  if (nm == "T") { r.kind = K_NODESCAL; r.dev = d.T; r.bytes = nnb; return d.T != nullptr; }
  if (nm == "m_q_plheat") return elems(d.q_plheat);
  if (nm == "q_cont_conv") { r.kind = K_NODESCAL; r.dev = d.q_cont_conv; r.bytes = nnb; return d.q_cont_conv != nullptr; }
  if (nm == "m_dTedt") { r.kind = K_ELEMNODE; r.bytes = nk; r.lazy_dtedt = true; return d.tsell != nullptr && !for_write; }
  if (nm == "m_tau") return elem6(d.tau);
  if (nm == "m_eps") return elem6(d.eps);
  if (nm == "m_str_rate") return elem6(d.str_rate);
  if (nm == "m_rot_rate") return elem6(d.rot_rate);
  if (nm == "m_sigma") {
    if (d.sigma && (E->P.store_sigma || E->sigma_in_dbg || for_write)) return elem6(d.sigma);
    r.kind = K_ELEM6; r.bytes = 6 * neb; r.lazy_sigma = true; return !for_write;
  }
  if (nm == "m_dH_detJ_dx" || nm == "m_dH_detJ_dy" || nm == "m_dH_detJ_dz") {
    r.kind = K_ELEMNODE; r.dev = d.dH; r.bytes = nk; r.comp = nm.back() - 'x';
    return d.dH != nullptr && r.comp < E->dim;
  }
  if (nm == "m_f_elem" || nm == "m_f_elem_hg") {
    r.kind = K_ELEMNODEVEC; r.bytes = nk * E->dim; r.hg = (nm == "m_f_elem_hg");
    if ((E->felem_in_dbg || for_write) && d.f_elem) { r.dev = r.hg ? d.f_elem_hg : d.f_elem; return r.dev != nullptr; }
    if (for_write) return false;
    r.lazy_felem = true; // rebuilt from the node-ordered buffer written by the fused step
    return r.hg ? d.fsell_hg != nullptr : true;
  }
  if (nm == "m_hg_q") { r.kind = K_HGQ; r.dev = d.hg_q; r.bytes = nk * E->dim; return d.hg_q != nullptr; }
  { // contact / rigid-surface arrays (wf_contact.cu)
    void *p = nullptr; size_t b = 0; int kind = 0;
    if (wf_contact_lookup(E, nm, &p, &b, &kind)) {
      r.bytes = b;
      if (kind == 4) { r.kind = K_INT_HOST; r.host = p; return !for_write; }
      r.dev = (double *)p;
      r.kind = kind == 0 ? K_NODEVEC : (kind == 1 ? K_NODESCAL : (kind == 2 ? K_ELEMSCAL : K_RAW_DEV));
      return !(for_write && r.kind == K_RAW_DEV);
    }
  }
  auto hosti = [&](const void *p, size_t b) { r.kind = K_INT_HOST; r.host = p; r.bytes = b; return !for_write; };
  if (nm == "m_elnod") return hosti(E->h_elnod.data(), E->h_elnod.size() * sizeof(unsigned));
  if (nm == "elem_perm") { // perm[internal] = user (identity when the mesh was not reordered)
    if (E->perm_out.size() != (size_t)E->ne) {
      E->perm_out.resize(E->ne);
      for (int e = 0; e < E->ne; e++) E->perm_out[e] = E->perm.empty() ? e : E->perm[e];
    }
    return hosti(E->perm_out.data(), E->perm_out.size() * sizeof(int));
  }
  if (nm == "m_nodel") return hosti(E->h_nodel.data(), E->h_nodel.size() * sizeof(int));
  if (nm == "m_nodel_loc") return hosti(E->h_nodel_loc.data(), E->h_nodel_loc.size() * sizeof(int));
  if (nm == "m_nodel_offset") return hosti(E->h_offset.data(), E->h_offset.size() * sizeof(int));
  if (nm == "m_nodel_count") return hosti(E->h_count.data(), E->h_count.size() * sizeof(int));
  return false;
}

extern "C" size_t wf_array_bytes(wf_engine *E, const char *name) {
  if (!E || !E->meshed || !name) return 0;
  ArrayRef r;
  if (!lookup(E, name, r, false)) return 0;
  return r.bytes;
}

static int download(wf_engine *E, const double *dev, size_t count, std::vector<double> &h) {
  h.resize(count);
  CK(cudaMemcpyAsync(h.data(), dev, count * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  return 0;
}

extern "C" int wf_get_array(wf_engine *E, const char *name, void *dst, size_t bytes) { WF_NULLCHK(E);
  NEED(E->meshed, "no mesh");
  NEED(name && dst, "null argument");
  NEED(!E->predicted, "engine is mid-batch");
  CK(cudaSetDevice(E->device));
  if (check_halo_error(E)) return 1;
  ArrayRef r;
  if (!lookup(E, name, r, false))
    FAIL(std::string("array '") + name + "' is unknown or only produced by the unfused entry points / tracking options");
  NEED(bytes == r.bytes, std::string("size mismatch for '") + name + "'");
  WfDev &d = E->d;
  const int nn = E->nn, ne = E->ne, k = E->k, dim = E->dim;
  double *out = (double *)dst;
  std::vector<double> h;
  if (r.kind == K_INT_HOST) { memcpy(dst, r.host, bytes); return 0; }
  if (r.kind == K_RAW_DEV) {
    CK(cudaMemcpyAsync(dst, r.dev, bytes, cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return 0;
  }
  if (r.lazy_pnode) { // calcNodalPressureFromElemental (Mechanical.C:1187-1212) on the device
    if (need_scratch(E, (size_t)d.np)) return 1;
    E->L->p_node(d, E->scratch, E->stream);
    CK(cudaMemcpyAsync(dst, E->scratch, bytes, cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return check_launch(E, "p_node");
  }
  if (r.lazy_sigma) { // sigma = -p I + tau rebuilt, converted and copied without touching host scratch
    const size_t c6 = (size_t)6 * d.ep;
    if (need_scratch(E, 2 * c6)) return 1;
    E->L->rebuild_sigma(d, E->scratch, E->stream);
    E->L->soa_to_aos(E->scratch, d.ep, 6, ne, 1.0, E->scratch + c6, E->iperm_d, E->stream);
    CK(cudaMemcpyAsync(dst, E->scratch + c6, bytes, cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return check_launch(E, "m_sigma");
  } else if (r.lazy_dtedt) { // m_dTedt[e*k+ln] back from the node-ordered buffer
    std::vector<double> ts;
    if (download(E, d.tsell, (size_t)E->sell_total, ts)) return 1;
    for (int e2 = 0; e2 < ne; e2++)
      for (int n = 0; n < k; n++) {
        const size_t o = E->h_pos[(size_t)n * d.ep + e2];
        const size_t lane = o & 31;
        out[(size_t)e2 * k + n] = ts[(o + (size_t)(dim - 1) * lane) / dim];
      }
    return 0;
  } else if (r.lazy_felem) {
    NEED(E->step_count > 0, "m_f_elem is available after a step");
    NEED(!E->L->tile_forces(d, E->P, E->strict ? 1 : 0),
         "m_f_elem is not kept by the tile-reduced force path; use the strict engine or wf_set_variant(2, 9)");
    std::vector<double> fs;
    if (download(E, r.hg ? d.fsell_hg : d.fsell, (size_t)dim * E->sell_total, fs)) return 1;
    for (int e2 = 0; e2 < ne; e2++)
      for (int n = 0; n < k; n++) {
        const size_t o = E->h_pos[(size_t)n * d.ep + e2];
        for (int c = 0; c < dim; c++) out[((size_t)e2 * k + n) * dim + c] = fs[o + 32 * (size_t)c];
      }
    return 0;
  } else if (r.lazy_fi) {
    NEED(E->step_count > 0 || E->dbg, "m_fi is available after a step");
    const size_t cv = (size_t)dim * d.np;
    if (need_scratch(E, 2 * cv)) return 1;
    double *save_fi = d.fi;
    d.fi = E->scratch;
    E->L->node_update(d, E->P, E->strict ? 1 : 0, 0, 1, E->stream); // sums only, exactly as the step forms them
    d.fi = save_fi;
    E->L->soa_to_aos(E->scratch, d.np, dim, nn, 1.0, E->scratch + cv, nullptr, E->stream);
    CK(cudaMemcpyAsync(dst, E->scratch + cv, bytes, cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    return check_launch(E, "m_fi");
  } else {
    NEED(r.dev, std::string("array '") + name + "' is not allocated");
    if (r.kind != K_HGQ) {
      // convert to the reference layout on the device, then ONE device->host copy straight into the caller's buffer
      long long cnt = 0, pitch = 0; int nc = 1; const double *src = r.dev; double scale = 1.0;
      const int *map = nullptr; // element arrays live in the internal element order
      switch (r.kind) {
        case K_NODEVEC: cnt = nn; pitch = d.np; nc = dim; break;
        case K_NODESCAL: cnt = nn; pitch = d.np; nc = 1; if (r.lazy_voln) scale = 1.0 / (double)k; break;
        case K_ELEMSCAL: cnt = ne; pitch = d.ep; nc = 1; map = E->iperm_d; break;
        case K_ELEM6: cnt = ne; pitch = d.ep; nc = 6; map = E->iperm_d; break;
        case K_ELEMNODE: cnt = ne; pitch = d.ep; nc = k; src = r.dev + (size_t)r.comp * k * d.ep; map = E->iperm_d; break;
        case K_ELEMNODEVEC: cnt = ne; pitch = d.ep; nc = k * dim; map = E->iperm_d; break;
        default: break;
      }
      if (nc == 1 && scale == 1.0 && !map) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, E->stream));
      } else {
        if (need_scratch(E, (size_t)cnt * nc)) return 1;
        if (r.lazy_voln) E->L->soa_to_aos(src, pitch, nc, cnt, 1.0, E->scratch, map, E->stream); // then divide exactly like the host did
        else E->L->soa_to_aos(src, pitch, nc, cnt, scale, E->scratch, map, E->stream);
        CK(cudaMemcpyAsync(dst, E->scratch, bytes, cudaMemcpyDeviceToHost, E->stream));
      }
      CK(cudaStreamSynchronize(E->stream));
      if (r.lazy_voln) for (int n = 0; n < nn; n++) out[n] = out[n] / (double)k;
      return check_launch(E, "wf_get_array");
    }
    if (download(E, r.dev, (size_t)2 * d.ep, h)) return 1;
  }
  // host-side conversions of the rebuilt / rare arrays
  switch (r.kind) {
    case K_NODEVEC:
      for (int n = 0; n < nn; n++)
        for (int c = 0; c < dim; c++) out[(size_t)n * dim + c] = h[(size_t)c * d.np + n];
      break;
    case K_ELEM6:
      for (int e = 0; e < ne; e++)
        for (int c = 0; c < 6; c++) out[(size_t)e * 6 + c] = h[(size_t)c * d.ep + e];
      break;
    case K_HGQ:
      memset(out, 0, bytes);
      for (int e = 0; e < ne; e++)
        for (int c = 0; c < 2; c++) out[(size_t)e * 2 + c] = h[(size_t)c * d.ep + (E->iperm.empty() ? e : E->iperm[e])];
      break;
    default: break;
  }
  return 0;
}

extern "C" int wf_set_array(wf_engine *E, const char *name, const void *src, size_t bytes) { WF_NULLCHK(E);
  NEED(E->meshed, "no mesh");
  NEED(name && src, "null argument");
  NEED(!E->predicted, "engine is mid-batch");
  E->ekin_step = -1; // the state may change: a later monitor recomputes
  CK(cudaSetDevice(E->device));
  WfDev &d = E->d;
  const int nn = E->nn, ne = E->ne, k = E->k, dim = E->dim;
  std::string nm(name);
  if (nm == "m_fe" && !d.fe && dalloc(E, &d.fe, (size_t)dim * d.np)) return 1;
  if (nm == "m_eps" && !d.eps) FAIL("m_eps needs wf_set_tracking(bit0) before wf_init");
  if (nm == "m_sigma" && !d.sigma && dalloc(E, &d.sigma, (size_t)6 * d.ep)) return 1;
  if (nm == "a") { if (ensure_dbg(E)) return 1; E->a_in_dbg = true; }
  ArrayRef r;
  if (!lookup(E, nm, r, true)) FAIL(std::string("array '") + name + "' cannot be set");
  NEED(bytes == r.bytes, std::string("size mismatch for '") + name + "'");
  const double *in = (const double *)src;
  if (r.kind == K_HGQ) {
    std::vector<double> h((size_t)2 * d.ep, 0.0);
    for (int e = 0; e < ne; e++)
      for (int c = 0; c < 2; c++) h[(size_t)c * d.ep + (E->iperm.empty() ? e : E->iperm[e])] = in[(size_t)e * 2 + c];
    CK(cudaMemcpyAsync(r.dev, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
  } else {
    long long cnt = 0, pitch = 0; int nc = 1;
    const int *map = nullptr;
    switch (r.kind) {
      case K_NODEVEC: cnt = nn; pitch = d.np; nc = dim; break;
      case K_NODESCAL: cnt = nn; pitch = d.np; nc = 1; break;
      case K_ELEMSCAL: cnt = ne; pitch = d.ep; nc = 1; map = E->iperm_d; break;
      case K_ELEM6: cnt = ne; pitch = d.ep; nc = 6; map = E->iperm_d; break;
      case K_ELEMNODEVEC: cnt = ne; pitch = d.ep; nc = k * dim; map = E->iperm_d; break;
      default: FAIL(std::string("array '") + name + "' cannot be set");
    }
    if (nc == 1 && !map) {
      CK(cudaMemcpyAsync(r.dev, in, bytes, cudaMemcpyHostToDevice, E->stream));
    } else { // one host->device copy of the caller's buffer, layout conversion on the device
      if (need_scratch(E, (size_t)cnt * nc)) return 1;
      CK(cudaMemcpyAsync(E->scratch, in, bytes, cudaMemcpyHostToDevice, E->stream));
      E->L->aos_to_soa(E->scratch, pitch, nc, cnt, r.dev, map, E->stream);
    }
    CK(cudaStreamSynchronize(E->stream));
    if (check_launch(E, "wf_set_array")) return 1;
  }
  if (wf_contact_after_set(E, nm)) return 1;
  if ((nm == "vol_0" || nm == "rho") && E->inited) {
    // remesh hand-off (ReMesher::WriteDomain maps vol_0 and rho onto the new mesh): the nodal reference sums
    // (sum vol_0, mean rho, Solver_explicit.C:262) follow the uploaded element values
    NEED(!E->distributed, "vol_0 / rho can only be replaced on a single-GPU engine");
    E->L->node_vol(d, E->P, 0, E->stream);
    CK(cudaStreamSynchronize(E->stream));
    if (check_launch(E, "nodal reference sums")) return 1;
  }
  if (nm == "x" && E->domtype == WF_AXISYMM && E->inited) {
    if (reset_xmin(E, E->P.xmin_cur)) return 1;
    E->L->xmin(d, E->P.xmin_cur, E->stream);
  }
  return 0;
}

extern "C" void *wf_device_ptr(wf_engine *E, const char *name, size_t *pitch) {
  if (!E || !E->meshed) return nullptr;
  ArrayRef r;
  if (!lookup(E, name, r, true) || r.kind == K_INT_HOST) return nullptr;
  if (pitch) *pitch = (r.kind == K_NODEVEC || r.kind == K_NODESCAL) ? (size_t)E->d.np : (size_t)E->d.ep;
  return r.dev;
}

// ---- diagnostics on the device -----------------------------------------------------
static unsigned long long host_key(double v) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
static double host_unkey(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
  memcpy(&v, &b, 8);
  return v;
}
static int diag_reduce(wf_engine *E, bool edges, bool vel, double out[3]) {
  CK(cudaSetDevice(E->device));
  WfDev &d = E->d;
  if (!E->diag_keys && dalloc(E, &E->diag_keys, 3)) return 1;
  if (edges && !E->elem_length && dalloc(E, &E->elem_length, (size_t)d.ep)) return 1;
  const unsigned long long init[3] = {host_key(1.0e6), host_key(1.0e6), host_key(0.0)};
  CK(cudaMemcpyAsync(E->diag_keys, init, sizeof(init), cudaMemcpyHostToDevice, E->stream));
  if (edges) { E->L->min_edge(d, E->elem_length, E->diag_keys, E->stream); E->elem_length_valid = true; }
  if (vel) E->L->max_vel(d, E->diag_keys, E->stream);
  unsigned long long k[3];
  CK(cudaMemcpyAsync(k, E->diag_keys, sizeof(k), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  for (int i = 0; i < 3; i++) out[i] = host_unkey(k[i]);
  return check_launch(E, "diagnostics");
}

// Domain_d::calcMinEdgeLength (Domain_d.C:2224-2468): m_min_length, m_min_height, m_elem_length
extern "C" int wf_calcMinEdgeLength(wf_engine *E, double *min_length, double *min_height) { WF_NULLCHK(E);
  NEED(E->meshed, "no mesh");
  NEED(!E->predicted, "engine is mid-batch");
  NEED(E->dim == 3 || E->k == 4, "calcMinEdgeLength reads four nodes per element (Domain_d.C:2381-2384): not defined for triangles");
  double o[3];
  if (diag_reduce(E, true, false, o)) return 1;
  if (wf_contact_refresh_nodlen(E)) return 1;
  if (min_length) *min_length = o[0];
  if (min_height) *min_height = o[1];
  return 0;
}

// max |v| over the nodes (Solver_explicit.C:583-587)
extern "C" int wf_max_velocity(wf_engine *E, double *vmax) { WF_NULLCHK(E);
  NEED(E->meshed, "no mesh");
  NEED(!E->predicted, "engine is mid-batch");
  double o[3];
  if (diag_reduce(E, false, true, o)) return 1;
  if (vmax) *vmax = o[2];
  return 0;
}

// variable time step of the explicit loop (Solver_explicit.C:579-598): dt = cfl * min_length / (cs + max|v|),
// cs = sqrt(K / rho[0])
extern "C" int wf_cfl_dt(wf_engine *E, double cfl_factor, double *dt) { WF_NULLCHK(E);
  NEED(E->meshed && E->material_set, "wf_cfl_dt needs mesh and material");
  NEED(!E->predicted, "engine is mid-batch");
  NEED(E->dim == 3 || E->k == 4, "calcMinEdgeLength is not defined for triangles");
  double o[3], rho0 = E->mat.rho0;
  if (diag_reduce(E, true, true, o)) return 1;
  if (E->inited) {
    CK(cudaMemcpyAsync(&rho0, E->d.rho + (E->iperm.empty() ? 0 : E->iperm[0]), sizeof(double), cudaMemcpyDeviceToHost, E->stream)); // rho[0] of the caller's numbering
    CK(cudaStreamSynchronize(E->stream));
  }
  const double cs = sqrt(E->P.Kbulk / rho0);
  if (dt) *dt = cfl_factor * o[0] / (cs + o[2]);
  return 0;
}

// change the step size between batches (variable-dt loop); takes effect at the next wf_step
extern "C" int wf_set_dt(wf_engine *E, double dt) { WF_NULLCHK(E);
  NEED(E->inited, "wf_set_dt after wf_init");
  NEED(!E->predicted, "engine is mid-batch");
  NEED(dt > 0.0, "dt must be positive");
  E->P.dt = dt;
  return 0;
}

// computeEnergies (Mechanical.C:2145-2185).  Ekin from the current velocities and the nodal mass of the last
// step; dEint needs the strain rates of the last step, which only the unfused path keeps.
extern "C" int wf_energies(wf_engine *E, double *Ekin, double *dEint) { WF_NULLCHK(E);
  NEED(E->inited, "wf_energies before wf_init");
  NEED(!E->predicted, "engine is mid-batch");
  CK(cudaSetDevice(E->device));
  WfDev &d = E->d;
  CK(cudaMemsetAsync(d.red, 0, 8 * sizeof(double), E->stream));
  double *sig = d.sigma, *tmp = nullptr;
  bool have_rates = d.str_rate != nullptr && E->rates_in_dbg;
  if (have_rates && !(d.sigma && (E->P.store_sigma || E->sigma_in_dbg))) {
    CK(cudaMalloc((void **)&tmp, (size_t)6 * d.ep * sizeof(double)));
    E->L->rebuild_sigma(d, tmp, E->stream);
    sig = tmp;
  }
  if (have_rates) E->L->energy(d, sig, E->stream);
  else { // kinetic part only
    WfDev d2 = d;
    d2.ne = 0;
    E->L->energy(d2, sig, E->stream);
  }
  double red[8];
  CK(cudaMemcpyAsync(red, d.red, sizeof(red), cudaMemcpyDeviceToHost, E->stream));
  CK(cudaStreamSynchronize(E->stream));
  if (tmp) cudaFree(tmp);
  if (Ekin) *Ekin = red[0];
  if (dEint) *dEint = have_rates ? red[1] * E->P.dt : NAN;
  return check_launch(E, "wf_energies");
}
