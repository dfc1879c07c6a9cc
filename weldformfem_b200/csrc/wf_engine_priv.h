// wf_engine_priv.h — the engine object behind the opaque wf_engine handle (private to csrc/).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/wf_engine.h"
#include "wf_dev.h"
#include "wf_host.h"
#include "wf_launch.h"
#include "wf_math_ids.h"

int wf_check_launch(struct wf_engine *E, const char *what);

struct wf_engine {
  int device = 0;
  cudaStream_t stream = 0;
  std::string err;
  int dim = 3, k = 8, et = ET_HEX8, domtype = WF_3D;
  int nn = 0, ne = 0;
  WfDev d;
  WfPar P;
  const WfLaunch *L = nullptr;
  bool strict = false;
  int tracking = 0;
  bool meshed = false, material_set = false, bcs_ready = false, inited = false, dbg = false;
  bool predicted = false;  // v / u_dt currently hold next-step predictor values (only inside wf_step)
  // which unfused-path products are current (cleared by wf_step)
  bool a_in_dbg = false, fi_in_dbg = false, sigma_in_dbg = false, rates_in_dbg = false, felem_in_dbg = false;
  double time = 0.0;
  long step_count = 0;
  wf_material mat;
  wf_stab stab;
  std::vector<void *> allocs;
  // host copies of integer artefacts (reference layouts)
  std::vector<unsigned> h_elnod;
  std::vector<int> h_nodel, h_nodel_loc, h_offset, h_count;
  std::vector<unsigned> h_pos; // [k][ep], see WfDev::pos; indexed by USER element id
  // internal element order (wf_host_elem_order): perm[internal] = user, iperm[user] = internal; empty = identity
  bool axis_xmin_set = false; double axis_xmin = 0.0; // wf_set_axis_xmin (partitioned axisymmetric domains)
  int order_mode = -1;  // -1 = automatic: reorder hexahedra only (measured: tets 0.863 -> 0.968 ms, quads 0.173 -> 0.176 ms when reordered)
  std::vector<int> perm, iperm, perm_out;
  int *iperm_d = nullptr;
  long long sell_total = 0;
  std::vector<int> bc_nod[3];
  std::vector<double> bc_val[3];
  int nbc_rows = 0;
  std::vector<int> bc_slot[3];             // index into bc_vals of BC i of each dimension (-1: node of another rank)
  double *bc_stage[2] = {nullptr, nullptr}; // pinned staging of bc_vals for wf_set_bc_values
  cudaEvent_t bc_ev[2] = {nullptr, nullptr};
  int bc_stage_cur = 0;
  double *bc_vals_d = nullptr;
  // time-dependent prescribed values (wf_set_bc_values): two device copies of bc_vals, written alternately by a copy
  // stream while the step that reads the other one is still running
  double *bc_vals_buf[2] = {nullptr, nullptr};
  int bc_vals_cur = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t bc_at_ev[2] = {nullptr, nullptr}, bc_copy_ev = nullptr;
  long bc_set_calls = 0;
  int *bc_row_node_d = nullptr;            // node of every BC row (k_bc_patch_v)
  // wf_step_timed: when set, step_stage records a CUDA event after every launch (tag = slot of the caller's ms array)
  std::vector<cudaEvent_t> *tm_ev = nullptr;
  std::vector<int> *tm_tag = nullptr;
  bool open_mode = false;                  // wf_step_open: the call's last node pass also runs the next predictor
  bool udt_valid = true;                   // "u_dt" holds the last step's increment (not after wf_step_close)
  std::vector<double> bc_master;           // host copy of bc_vals
  int bc_version[3] = {0, 0, 0}, bc_stage_version[2][3] = {{0, 0, 0}, {0, 0, 0}};
  // asynchronous step monitor (wf_monitor_async / wf_monitor_wait): 2-deep ring of pinned results
  static constexpr int MON_NACC = 256;      // kinetic-energy accumulators per monitor slot (spreads the atomics of the node pass)
  struct MonSlot { double part[MON_NACC]; int nonfinite; int halo_error; };
  MonSlot *mon_host = nullptr;
  double *mon_red = nullptr;               // [2][MON_NACC] device partial sums
  cudaEvent_t mon_ev[2] = {nullptr, nullptr};
  int mon_head = 0, mon_pending = 0;
  // once the caller has asked for a monitor, the last node pass of every wf_step call accumulates the kinetic
  // energy itself (WfDev::ekin_acc) and wf_monitor_async only copies it; ekin_step = step count that sum belongs to
  bool mon_seen = false;
  long ekin_step = -1;
  int ekin_slot = -1;
  // partition / halo (multi-GPU); see wf_set_mesh_partition
  bool distributed = false, own_stream = false;
  int rank = 0, nranks = 1;
  int transport = 0;                       // 0 = stores into the neighbour's memory, 1 = host-driven (NCCL) via the staging block
  std::vector<int> l2g, neigh, halo_offset, halo_nodes_h;
  std::vector<WfHaloNb> nb_h;
  WfHaloNb *nb_d = nullptr;
  unsigned *counters_d = nullptr;
  char *comm = nullptr, *staging = nullptr; // [flags | receive regions]
  size_t comm_bytes = 0, flag_bytes = 0;
  int max_halo_count = 0, n_connected = 0;
  unsigned long long seq = 0, timeout_ns = 30000000000ull;
  std::vector<void *> ipc_opened;
  int init_stage = 0, step_stage = 0;
  // scratch for device-side layout conversion (wf_get_array / wf_set_array), diagnostics
  double *scratch = nullptr;
  size_t scratch_count = 0;
  double *elem_length = nullptr;           // m_elem_length (calcMinEdgeLength)
  unsigned long long *diag_keys = nullptr; // [3] ordered keys: min length, min height, max |v|
  bool elem_length_valid = false;
  // contact with rigid surfaces (wf_contact.cu)
  bool ext_searched = false, trimesh_set = false, contact = false;
  WfContact C;                             // device view, zero until wf_SearchExtNodes / wf_set_trimesh
  std::vector<unsigned char> h_ext;        // ext_nodes (Domain_d.C:137-156), host copy
  int face_count = 0;                      // m_faceCount
  double end_t = 0.0;                      // Domain_d::end_t (velocity ramp of the rigid surfaces)
  double *tm_stage = nullptr;
};

// wf_contact.cu
int wf_contact_step_begin(wf_engine *E);           // CalcExtFaceAreas cadence, Solver_explicit.C:445-450
int wf_contact_forces(wf_engine *E);               // CalcContactForces, before the nodal update
int wf_contact_step_end(wf_engine *E);             // ramp + Move + normals + plane coefficients, :981-1005
int wf_contact_init(wf_engine *E);                 // m_v_orig, ut_prev = 0
int wf_contact_refresh_nodlen(wf_engine *E);
int wf_contact_after_set(wf_engine *E, const std::string &nm);
bool wf_contact_lookup(wf_engine *E, const std::string &nm, void **dev, size_t *bytes, int *kind);

int wf_null_engine(void); // records "null engine handle" for wf_last_error(NULL), returns 1
#define WF_NULLCHK(E) do { if (!(E)) return wf_null_engine(); } while (0)
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

int wf_check_launch(wf_engine *E, const char *what);
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

template <class T>
static inline int dalloc(wf_engine *E, T **p, size_t count) {
  void *q = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  CK(cudaMalloc(&q, bytes));
  CK(cudaMemsetAsync(q, 0, bytes, E->stream));
  E->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}

static inline int need_scratch(wf_engine *E, size_t count) {
  if (count <= E->scratch_count) return 0;
  if (E->scratch) { CK(cudaStreamSynchronize(E->stream)); cudaFree(E->scratch); E->scratch = nullptr; E->scratch_count = 0; }
  CK(cudaMalloc((void **)&E->scratch, count * sizeof(double)));
  E->scratch_count = count;
  return 0;
}

