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
