// wf_host.h — host-side helpers shared by wf_mesh.cpp and wf_engine.cu (private).
#pragma once
#include <vector>

#define WF_MAXK_HOST 8

struct WfBox {
  int dim, k, tritet, per_cell;
  int nel[3];
  long long nn, ne;
};

void wf_box_dims(const double L[3], double r, int tritet, WfBox *b);
void wf_box_elem_nodes(const WfBox &b, long long e, unsigned *out);
void wf_box_axes(const WfBox &b, const double V[3], double r, std::vector<double> ax[3]);
void wf_box_node_xyz(const WfBox &b, const std::vector<double> ax[3], long long n, double *out);

struct wf_partition;
int wf_partition_k(const wf_partition *p);
void wf_partition_ranks(const wf_partition *p, int *rank, int *nranks);
bool wf_partition_is_box(const wf_partition *p);
void wf_partition_box_coords(const wf_partition *p, std::vector<double> &x);
int wf_partition_box_dim(const wf_partition *p);

// Force tiles of the tile-reduced force path (WfDev::ftile; see wf_dev.h): tile w = thread slots [32w, 32w+32) of the
// main element pass; slot s holds element slot_elem[s] (-1 = idle), or element s when slot_elem is NULL.
struct WfForceTiles {
  bool usable = false;               // hexahedra: false when two elements of a tile share a node at the same corner
  bool rounds = false;               // no tile has two elements sharing a node at the same corner (conflict-free rounds)
  int k = 0, n_tiles = 0, stride = 0, tpitch = 0;
  std::vector<unsigned char> tidx;   // [k][ep] index of element node (slot, ln) in its tile's ascending unique-node list
  std::vector<long long> ptr;        // [nslices+1] sliced-ELL of the tile entries of each node
  std::vector<unsigned> slots;       // offset of component 0 in ftile (tile*dim*stride + position), ascending tile order, ~0u = padding
  std::vector<unsigned char> tab;    // all but hexahedra: [n_tiles][tpitch] incidence tables (ptr[stride+1], inc[32k])
};
// n_slots = thread slots (= n_elems when slot_elem is NULL); ep = pitch of tidx (>= n_slots); elnod is the reference
// layout [e*k + ln]
void wf_force_tiles_build(int n_nodes, int n_slots, const int *slot_elem, int k, int dim, long long ep, const unsigned *elnod,
                          WfForceTiles &out);
