// wf_launch.h — launcher table exported by each numerics flavour of wf_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include "wf_dev.h"

struct WfLaunch {
  void (*predict)(const WfDev &, const WfPar &, int with_bc, cudaStream_t);
  void (*impose_bc)(const WfDev &, int dim, int is_acc, double *arr, cudaStream_t);
  void (*elem_vol)(const WfDev &, const WfPar &, int et, int store_jac, cudaStream_t);
  void (*vol_from_detj)(const WfDev &, int et, cudaStream_t);
  void (*node_vol)(const WfDev &, const WfPar &, int mode, cudaStream_t);
  void (*elem_main)(const WfDev &, const WfPar &, int et, int separate_hg, cudaStream_t);
  void (*node_update)(const WfDev &, const WfPar &, int separate_hg, int fuse_predictor, int phase, cudaStream_t);
  void (*node_mass)(const WfDev &, const WfPar &, int use_stored_voln, cudaStream_t);
  void (*init_elem)(const WfDev &, const WfPar &, cudaStream_t);
  void (*vol0_density)(const WfDev &, cudaStream_t);
  void (*density)(const WfDev &, cudaStream_t);
  void (*xmin)(const WfDev &, int slot, cudaStream_t);
  void (*rebuild_sigma)(const WfDev &, double *out, cudaStream_t);
  void (*energy)(const WfDev &, const double *sig, cudaStream_t);
  void (*u_strain_rates)(const WfDev &, const WfPar &, int et, cudaStream_t);
  void (*u_pressure)(const WfDev &, const WfPar &, int et, cudaStream_t);
  void (*u_stress)(const WfDev &, const WfPar &, double dt, cudaStream_t);
  void (*u_artvisc)(const WfDev &, const WfPar &, cudaStream_t);
  void (*u_forces)(const WfDev &, const WfPar &, int et, cudaStream_t);
  void (*u_hourglass)(const WfDev &, const WfPar &, int et, cudaStream_t);
  void (*u_nodal_vol)(const WfDev &, cudaStream_t);
  void (*u_assembly)(const WfDev &, cudaStream_t);
  void (*u_accel)(const WfDev &, cudaStream_t);
  void (*u_corr_accvel)(const WfDev &, const WfPar &, cudaStream_t);
  void (*u_axis)(const WfDev &, const WfPar &, cudaStream_t);
  void (*u_corr_pos)(const WfDev &, const WfPar &, cudaStream_t);
  void (*halo_send)(const WfDev &, const WfPar &, int mode, int separate_hg, unsigned long long seq, int max_count, cudaStream_t);
  void (*halo_wait)(const WfDev &, unsigned long long seq, unsigned long long timeout_ns, cudaStream_t);
  void (*halo_finish)(const WfDev &, const WfPar &, int mode, int parity, cudaStream_t);
  void (*preload)(int et, int dim, int k);
  void (*p_node)(const WfDev &, double *out, cudaStream_t);
  void (*min_edge)(const WfDev &, double *elem_length, unsigned long long *keys3, cudaStream_t);
  void (*max_vel)(const WfDev &, unsigned long long *keys3, cudaStream_t);
  void (*soa_to_aos)(const double *soa, long long pitch, int nc, long long n, double scale, double *aos, const int *map, cudaStream_t);
  void (*aos_to_soa)(const double *aos, long long pitch, int nc, long long n, double *soa, const int *map, cudaStream_t);
  void (*node_thermal)(const WfDev &, const WfPar &, cudaStream_t);
  int (*tile_forces)(const WfDev &, const WfPar &, int separate_hg); /* does the step use WfDev::ftile instead of fsell? */
  void (*bc_patch_v)(const WfDev &, const int *row_node, int nrows, cudaStream_t);
  void (*unpredict)(const WfDev &, const WfPar &, cudaStream_t);
};

extern "C" const WfLaunch *wf_strict_table();
extern "C" const WfLaunch *wf_fast_table();
