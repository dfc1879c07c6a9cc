// wf_kernels.cu — the CUDA kernels of the explicit step (sm_100a, fp64).
//
// Compiled TWICE into the same library:
//   -DWF_NS=wf_strict -fmad=false : operation-for-operation with the reference CPU path
//   -DWF_NS=wf_fast   -fmad=true  : same source, FMA contraction on; plus the regrouped hexa kernel
// Each flavour exports one launcher table (wf_launch.h).
//
// Per step the fused schedule is four passes:
//   E1 k_elem_vol    : x -> vol                                  (calcElemJAndDerivatives + CalcElemVol)
//   N1 k_node_vol    : vol gather -> nodal sums / ratios         (CalcNodalVol, node part of calcElemPressure*)
//   E2 k_elem_main   : J, dH, D, W, pressure, Jaumann + J2 return, element + hourglass forces
//   N2 k_node_update : per-node force sum, accel, BCs, corrector, position, next-step predictor
// Element forces travel from E2 to N2 in one of two ways, both deterministic gathers without atomics:
//   * strict flavour and 2D: the node-ordered buffer fsell — the contribution of (element e, local node ln) is written
//     straight into the entry of node n's nodel list that the reference's assemblyForces would read for it
//     (pos[ln][e]), so N2 streams its list with coalesced loads and sums it in nodel order;
//   * fast flavour, hexahedra / tetrahedra: tile partials ftile — the forces of the 32 elements of a warp are summed
//     per unique node in shared memory (fixed order) and one partial per (tile, node) is written; N2 gathers the
//     partials of a node in ascending tile order (WfDev::ftile, wf_dev.h).
#include <cuda_runtime.h>
#include <math.h>

#include "wf_launch.h"
#include "wf_math.cuh"

#ifndef WF_NS
#error "compile with -DWF_NS=wf_strict or -DWF_NS=wf_fast"
#endif

namespace WF_NS {

WF_DI void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Programmatic dependent launch (sm_90+): the four kernels of the step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the CTAs of the next pass become resident while the last wave
// of the current one drains, and do their CONSTANT-table loads (connectivity, slot tables, node lists) meanwhile.
// Rule kept by every kernel below: pdl_trigger() first; nothing a previous kernel may have written is read, and nothing
// is written, before pdl_wait() (= the prerequisite grids have completed and their memory operations are visible).
// L2 prefetches are allowed earlier (L2 is the coherence point).  Launched without the attribute both are no-ops.
WF_DI void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::); }
WF_DI void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#include "wf_hex_fast.cuh"

constexpr int TPB_E = 128; // element kernels: register-heavy
constexpr int TPB_N = 256; // node kernels (multiple of 32: one warp == one SELL slice)

// ---------------------------------------------------------------------------------------------
// gathers of element-local node data
// ---------------------------------------------------------------------------------------------
template <int ET>
WF_DI void load_conn(const WfDev &d, int e, int (&nid)[Elem<ET>::K]) {
#pragma unroll
  for (int n = 0; n < Elem<ET>::K; n++) nid[n] = __ldg(d.elnod + (long long)n * d.ep + e);
}
template <int ET>
WF_DI void gather_nodal(const double *__restrict__ q, long long np, const int (&nid)[Elem<ET>::K],
                        double (&out)[Elem<ET>::K][Elem<ET>::D]) {
#pragma unroll
  for (int n = 0; n < Elem<ET>::K; n++)
#pragma unroll
    for (int c = 0; c < Elem<ET>::D; c++) out[n][c] = q[(long long)c * np + nid[n]];
}

// ---------------------------------------------------------------------------------------------
// predictor (UpdatePrediction, Domain_d.C:961-974) + ImposeBCV for all dims (:1109-1121)
// ---------------------------------------------------------------------------------------------
template <int D>
WF_DI void apply_bcv(const WfDev &d, int n, double (&v)[D]) {
  int bi = d.bc_index[n];
  if (bi >= 0) {
    unsigned m = d.bc_mask[bi];
#pragma unroll
    for (int c = 0; c < D; c++)
      if (m & (1u << c)) v[c] = d.bc_vals[3 * bi + c];
  }
}

// UpdatePrediction + ImposeBCV of one node (same arithmetic as k_predict)
template <int D>
WF_DI void predict_node(const WfDev &d, const WfPar &P, int n) {
  double v[D];
#pragma unroll
  for (int c = 0; c < D; c++) {
    const long long i = (long long)c * d.np + n;
    const double pa = d.prev_a[i], vv = d.v[i];
    d.u_dt[i] = P.dt * (vv + (0.5 - P.beta) * P.dt * pa);
    v[c] = vv + (1.0 - P.gamma) * P.dt * pa;
  }
  apply_bcv<D>(d, n, v);
#pragma unroll
  for (int c = 0; c < D; c++) d.v[(long long)c * d.np + n] = v[c];
}

template <int D>
__global__ void __launch_bounds__(TPB_N) k_predict(WfDev d, WfPar P, int with_bc) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
  double v[D];
#pragma unroll
  for (int c = 0; c < D; c++) {
    long long i = (long long)c * d.np + n;
    double pa = d.prev_a[i], vv = d.v[i];
    d.u_dt[i] = P.dt * (vv + (0.5 - P.beta) * P.dt * pa);
    v[c] = vv + (1.0 - P.gamma) * P.dt * pa;
  }
  if (with_bc) apply_bcv<D>(d, n, v);
#pragma unroll
  for (int c = 0; c < D; c++) d.v[(long long)c * d.np + n] = v[c];
}

// ImposeBCV(d) / ImposeBCA(d) as standalone kernels (unfused path)
__global__ void k_impose_bc(WfDev d, int dim, int is_acc, double *a_or_v) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
  int bi = d.bc_index[n];
  if (bi < 0) return;
  if (d.bc_mask[bi] & (1u << dim)) a_or_v[(long long)dim * d.np + n] = is_acc ? 0.0 : d.bc_vals[3 * bi + dim];
}

// wf_set_bc_values on an engine that is in predicted state (wf_step_open): the velocities of the prescribed components
// were set by the fused predictor of the previous node pass; overwrite them with the new values (= the ImposeBCV the
// reference runs after UpdatePrediction, Solver_explicit.C:535-540).  One thread per BC row.
__global__ void k_bc_patch_v(WfDev d, const int *__restrict__ row_node, int nrows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int n = row_node[r];
  const unsigned m = d.bc_mask[r];
  for (int c = 0; c < d.dim; c++)
    if (m & (1u << c)) d.v[(long long)c * d.np + n] = d.bc_vals[3 * r + c];
}
// wf_step_close: undo the fused predictor, v_c = v_p - (1 - gamma) dt a (prescribed components have a = 0)
template <int D>
__global__ void k_unpredict(WfDev d, WfPar P) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
#pragma unroll
  for (int c = 0; c < D; c++) {
    const long long i = (long long)c * d.np + n;
    d.v[i] = d.v[i] - (1.0 - P.gamma) * P.dt * d.prev_a[i];
  }
}

// ---------------------------------------------------------------------------------------------
// E1: element volume from current coordinates
// ---------------------------------------------------------------------------------------------
// (the bodies of the four step kernels are device functions of a virtual block index; the __global__ wrappers pass
// blockIdx.x.  A persistent cooperative kernel that walked them in a loop with grid barriers was measured on the
// 105 k-tet mesh and rejected: 42 us per step against 24.5 us for the four launches with programmatic dependent launch)
template <int ET>
WF_DI void elem_vol_body(const WfDev &d, const WfPar &P, int store_jac, int vbx) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  pdl_trigger();
  int e = vbx * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  int nid[K];
  load_conn<ET>(d, e, nid);
  // (an L2 look-ahead of the connectivity as in elem_main_body was measured here: no difference, 0.0933 vs 0.0936 ms)
  pdl_wait();
  double xl[K][D], A[D][D], detJ;
  gather_nodal<ET>(d.x, d.np, nid, xl);
  jac_adj_det<ET>(xl, A, detJ);
  double radius = 0.0;
  if (D == 2 && d.domtype == 2) radius = elem_radius<ET>(xl);
  if (store_jac) { // unfused calcElemJAndDerivatives / Calc_Element_Radius products
    double dH[D][K];
    shape_derivs<ET>(A, dH);
#pragma unroll
    for (int c = 0; c < D; c++)
#pragma unroll
      for (int n = 0; n < K; n++) d.dH[((long long)c * K + n) * d.ep + e] = dH[c][n];
    d.detJ[e] = detJ;
    if (store_jac == 2) { d.radius[e] = radius; return; }
    if (store_jac == 1) return;
  }
  d.vol[e] = elem_volume<ET>(detJ, radius, d.domtype, d.vol_weight);
}
template <int ET>
__global__ void __launch_bounds__(TPB_E) k_elem_vol(WfDev d, WfPar P, int store_jac) {
  elem_vol_body<ET>(d, P, store_jac, blockIdx.x);
}

// E1 with the CTA's unique nodes staged once in shared memory (WfDev::blk_off / lidx)
template <int ET>
__global__ void __launch_bounds__(WF_EBLK) k_elem_vol_staged(WfDev d, WfPar P, int stride) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  extern __shared__ double sm[];
  const int t = threadIdx.x, b = blockIdx.x;
  const int u0 = __ldg(d.blk_off + b), U = __ldg(d.blk_off + b + 1) - u0;
  for (int i = t; i < U; i += WF_EBLK) {
    const int g = __ldg(d.blk_nodes + u0 + i);
#pragma unroll
    for (int c = 0; c < D; c++) sm[c * stride + i] = d.x[(long long)c * d.np + g];
  }
  const int e = b * WF_EBLK + t;
  unsigned li[K];
  if (e < d.ne) {
#pragma unroll
    for (int n = 0; n < K; n++) li[n] = d.lidx[(long long)n * d.ep + e];
  }
  __syncthreads();
  if (e >= d.ne) return;
  double xl[K][D], A[D][D], detJ;
#pragma unroll
  for (int n = 0; n < K; n++)
#pragma unroll
    for (int c = 0; c < D; c++) xl[n][c] = sm[c * stride + li[n]];
  jac_adj_det<ET>(xl, A, detJ);
  double radius = 0.0;
  if (D == 2 && d.domtype == 2) radius = elem_radius<ET>(xl);
  d.vol[e] = elem_volume<ET>(detJ, radius, d.domtype, d.vol_weight);
}

// E1, brick form (hexahedra whose CTA node lists fit the bank-aware layout, WfDev::blk_pad_b): the CTA stages the
// coordinates of its unique nodes once (coalesced runs of node ids instead of 24 scattered 8-byte loads per element:
// in the engine's element order the per-element gathers are L1-request-bound) and every element reads its eight
// nodes conflict-free through the packed slots.  Same arithmetic as k_elem_vol: identical volumes.
template <int STRIDE>
__global__ void __launch_bounds__(WF_EBLK) k_elem_vol_brick(WfDev d, WfPar P) {
  extern __shared__ double sm[];
  pdl_trigger();
  const int t = threadIdx.x, b = blockIdx.x;
  constexpr int NQ = (STRIDE + WF_EBLK - 1) / WF_EBLK;
  const int *__restrict__ ids = d.blk_pad_b + (long long)b * STRIDE;
  int gid[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int i = q * WF_EBLK + t;
    gid[q] = (i < STRIDE) ? __ldg(ids + i) : -1;
  }
  const int e = __ldg(d.brick_elem + (long long)b * WF_EBLK + t); // < 0: idle thread slot
  const bool active = e >= 0;
  const uint4 lpk = __ldg(d.lidx_pk + (long long)b * WF_EBLK + t);
  if (t < (STRIDE * 4 + 127) / 128) { // node list of the CTA that follows on this SM (see k_elem_main_hex_brick)
    const long long nb = (long long)b + 4LL * d.cta_lookahead;
    if (nb < d.n_bcta) prefetch_l2(reinterpret_cast<const char *>(d.blk_pad_b + nb * STRIDE) + t * 128);
  }
  pdl_wait();
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int i = q * WF_EBLK + t, gq = gid[q];
    if (gq >= 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) hexfast::cp_async8(sm + c * STRIDE + i, d.x + (long long)c * d.np + gq);
    }
  }
  hexfast::cp_async_commit();
  unsigned li[8];
  li[0] = lpk.x & 0xffffu; li[1] = lpk.x >> 16; li[2] = lpk.y & 0xffffu; li[3] = lpk.y >> 16;
  li[4] = lpk.z & 0xffffu; li[5] = lpk.z >> 16; li[6] = lpk.w & 0xffffu; li[7] = lpk.w >> 16;
  hexfast::cp_async_wait_all();
  __syncthreads();
  if (!active) return;
  double xl[8][3], A[3][3], detJ;
#pragma unroll
  for (int n = 0; n < 8; n++)
#pragma unroll
    for (int c = 0; c < 3; c++) xl[n][c] = sm[c * STRIDE + li[n]];
  jac_adj_det<ET_HEX8>(xl, A, detJ);
  d.vol[e] = elem_volume<ET_HEX8>(detJ, 0.0, d.domtype, d.vol_weight);
}

// CalcElemVol on stored detJ / radius (unfused)
template <int ET>
__global__ void k_vol_from_detj(WfDev d) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  d.vol[e] = elem_volume<ET>(d.detJ[e], d.radius ? d.radius[e] : 0.0, d.domtype, d.vol_weight);
}

// ---------------------------------------------------------------------------------------------
// multi-GPU halo primitives used INSIDE the node passes (peer transport; the stand-alone kernels are further below)
// ---------------------------------------------------------------------------------------------
WF_DI void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
WF_DI unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
WF_DI unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Every CTA of a consumer kernel: thread 0 waits (acquire, system scope) until all neighbours have published exchange
// `seq`, the rest of the CTA waits at the barrier.  A timeout is sticky: comm_error stays set, later waits return at
// once (the shared-node state is stale from the first missed exchange on), the non-finite flag is raised, and every
// host entry point that reads the engine reports the error.
WF_DI void halo_wait_cta(const WfDev &d, unsigned long long seq, unsigned long long timeout_ns) {
  if (threadIdx.x == 0 && *(volatile int *)d.comm_error == 0) {
    const unsigned long long t0 = global_ns();
    for (int i = 0; i < d.n_neigh; i++)
      while (ld_acquire_sys(d.flags + i) < seq) {
        if (global_ns() - t0 > timeout_ns) { atomicExch(d.comm_error, 1 + i); atomicExch(d.nonfinite, 1); i = d.n_neigh; break; }
        __nanosleep(100);
      }
  }
  __syncthreads();
}
// partial sums over the local nodel list of node n, in list order
// WITH_RHO: also the density sum and the count (init exchange only)
template <bool WITH_RHO = true>
WF_DI void halo_node_sums(const WfDev &d, int n, const double *__restrict__ src, double &s, double &sq, double &rs, int &cnt) {
  const long long base = d.sell_ptr[n >> 5];
  const int width = (int)((d.sell_ptr[(n >> 5) + 1] - base) >> 5);
  const int lane = n & 31;
  s = 0.0; sq = 0.0; rs = 0.0; cnt = 0;
  for (int j = 0; j < width; j++) {
    int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
    if (slot >= 0) {
      const int e = slot / d.k;
      const double ve = src[e];
      s += ve;
      sq += ve / 4.0;
      if (WITH_RHO) { rs += d.rho[e]; cnt++; }
    }
  }
}
// sum of the tile partials of node n (tile-reduced force path, WfDev::ftile), ascending tile order
WF_DI void tile_node_force(const WfDev &d, int n, double (&fi)[3]) {
  const long long base = d.tf_ptr[n >> 5];
  const int width = (int)((d.tf_ptr[(n >> 5) + 1] - base) >> 5);
  fi[0] = fi[1] = fi[2] = 0.0;
  for (int j = 0; j < width; j++) {
    const unsigned o = __ldg(d.tf_slots + base + ((long long)j << 5) + (n & 31));
    if (o == 0xFFFFFFFFu) continue;
#pragma unroll
    for (int c = 0; c < 3; c++)
      if (c < d.dim) fi[c] += d.ftile[(long long)o + (long long)c * d.tf_stride];
  }
}
WF_DI void halo_node_force(const WfDev &d, int n, int sep, double (&fi)[3]) {
  if (sep == 2) { tile_node_force(d, n, fi); return; }
  const long long base = d.sell_ptr[n >> 5];
  const int width = (int)((d.sell_ptr[(n >> 5) + 1] - base) >> 5);
  const int D = d.dim;
  const double *__restrict__ row = d.fsell + base * D + (n & 31);
  fi[0] = fi[1] = fi[2] = 0.0;
  for (int j = 0; j < width; j++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      if (c < D) fi[c] += row[((long long)j * D + c) * 32];
  if (sep) {
    const double *__restrict__ rowh = d.fsell_hg + base * D + (n & 31);
    for (int j = 0; j < width; j++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        if (c < D) fi[c] -= rowh[((long long)j * D + c) * 32];
  }
}
// One CTA of a halo send: chunk `chunk` of neighbour `inb` (any block size).  Recomputes this rank's partial sums for
// its shared nodes (same list order as the node kernels, so the values are the ones those kernels form) and stores them
// straight into the neighbour's receive region over NVLink; the last of the neighbour's `chunks` CTAs publishes the
// exchange number in the neighbour's flag slot (system-scope release).  MODE 0 = init triple (sum vol_0, sum rho,
// count), 1 = sum vol (+ sum vol/4 for ANP_Nodal), 2 = internal-force partial.
template <int MODE>
WF_DI void halo_send_cta(const WfDev &d, const WfPar &P, int sep, unsigned long long seq, int inb, int chunk, int chunks) {
  const WfHaloNb nb = d.nb[inb];
  const int j = chunk * blockDim.x + threadIdx.x;
  if (j < nb.count) {
    const int n = d.halo_nodes[nb.offset + j];
    double vals[WF_HALO_NC] = {0.0, 0.0, 0.0};
    int nc = WF_HALO_NC;
    if (MODE == 2) {
      halo_node_force(d, n, sep, vals);
      nc = d.dim;
    } else {
      double s, sq, rs;
      int cnt;
      halo_node_sums<MODE == 0>(d, n, MODE == 0 ? d.vol_0 : d.vol, s, sq, rs, cnt);
      if (MODE == 0) { vals[0] = (P.press == 3) ? sq : s; vals[1] = rs; vals[2] = (double)cnt; }
      else { vals[0] = s; vals[1] = sq; nc = 2; }
    }
    double *dst = nb.dst + (long long)(seq & 1ull) * WF_HALO_NC * nb.count + j;
#pragma unroll
    for (int c = 0; c < WF_HALO_NC; c++)
      if (c < nc) dst[(long long)c * nb.count] = vals[c];
  }
  // the CTA's stores are ordered before thread 0's system-scope fence by the barrier (fence cumulativity: the pattern
  // of a barrier followed by ONE posting thread's fence + flag store); a fence per storing thread is not needed and kept
  // every warp of the send CTAs waiting for its own NVLink round trip
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned done = atomicAdd(nb.counter, 1u);
    if (done == (unsigned)chunks - 1u) {
      *nb.counter = 0u;
      __threadfence_system();
      st_release_sys(nb.flag, seq);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// N1: nodal gathers of element volume (one warp == one SELL slice)
//   mode 0 (init)  : voln0_sum = sum vol_0 (or sum vol_0/4.0 for press 3)
//   mode 1 (step)  : voln_sum = sum vol ; nodal_p = ratio (press 0) or nodal pressure (press 1/3)
// CalcNodalVol (Mechanical.C:1555-1572) and the node loops of calcElemPressure (:697-705),
// calcElemPressureANP (:1228-1240), calcElemPressureANP_Nodal (:1262-1284).
// ---------------------------------------------------------------------------------------------
//   mode 3 (step)  : mode 1 + UpdatePrediction and ImposeBCV of the node (Domain_d.C:961-974): the first step of a
//                    batch has no previous node pass to carry its predictor (WF_FAST; strict runs k_predict)
// HALO: instantiation for partitioned meshes (carries the folded halo send; kept out of the single-GPU kernel, which is
// register-capped)
template <int K, bool HALO>
WF_DI void node_vol_body(const WfDev &d, const WfPar &P, int mode_in, int vbx) {
  const bool with_predict = mode_in == 3;
  const int mode = with_predict ? 1 : mode_in;
  pdl_trigger();
  int bx = vbx;
  if (HALO && mode == 1 && P.send_ctas > 0) {
    // multi-GPU: the first CTAs of the launch send this rank's partial volume sums of the shared nodes (they depend on
    // the element volumes only), so the transfer travels while the rest of the grid forms the nodal sums
    if (bx < P.send_ctas) {
      pdl_wait();
      halo_send_cta<1>(d, P, 0, P.send_seq, bx / P.send_chunks, bx % P.send_chunks, P.send_chunks);
      return;
    }
    bx -= P.send_ctas;
  }
  int n = bx * blockDim.x + threadIdx.x;
  int slice = n >> 5;
  if (slice >= d.nslices) { if (n == 0) pdl_wait(); return; }
  const long long base = d.sell_ptr[slice];
  const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
  const int lane = n & 31;
  pdl_wait();
  if (n == 0 && mode == 1 && d.xmin_key) d.xmin_key[P.xmin_cur ^ 1] = dbl_key(1000.0);
  double s = 0.0, sq = 0.0;
  const double *src = (mode == 0) ? d.vol_0 : d.vol;
  const bool quarter = (P.press == 3);
  // mode 1: the node's reference sums do not depend on the gathers; load them first
  double v0n = 1.0, rbn = 0.0;
  if (mode == 1 && n < d.nn) { v0n = d.voln0_sum[n]; if (!P.strict) rbn = d.rhobar[n]; }
  // eight list entries per trip: the slot loads, then the volume gathers, are independent and in flight together
  for (int j0 = 0; j0 < width; j0 += 8) {
    int sl[8];
    double ve[8];
#pragma unroll
    for (int q = 0; q < 8; q++) sl[q] = (j0 + q < width) ? __ldg(d.sell_slots + base + ((long long)(j0 + q) << 5) + lane) : -1;
#pragma unroll
    for (int q = 0; q < 8; q++) ve[q] = (sl[q] >= 0) ? src[sl[q] / K] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (sl[q] >= 0) {
        s += ve[q];
        if (quarter) sq += ve[q] / 4.0;
      }
  }
  if (n >= d.nn) return;
  if (mode == 0) {
    d.voln0_sum[n] = quarter ? sq : s;
    // mean density of the elements around the node (rho is frozen after init on the reference's CPU
    // path, Solver_explicit.C:262 vs :601-608), used by the WF_FAST nodal mass
    double rs = 0.0;
    for (int j = 0; j < width; j++) {
      int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
      if (slot >= 0) rs += d.rho[slot / K];
    }
    d.rhobar[n] = rs / (double)d.nodel_count[n];
    return;
  }
  d.voln_sum[n] = s;
  if (P.press == 0) d.nodal_p[n] = s / v0n;
  else if (P.press == 1) d.nodal_p[n] = P.Kbulk * (1.0 - s / v0n);
  else {
    double v0 = v0n, pn = 0.0;
    if (v0 > 1e-12) { double Jn = sq / v0; pn = P.Kbulk * (1.0 - Jn); }
    d.nodal_p[n] = pn;
  }
  if (with_predict) {
    if (d.dim == 3) predict_node<3>(d, P, n); else predict_node<2>(d, P, n);
  }
  // CalcNodalMassFromVol (Mechanical.C:1576-1601): mass = sum_e rho[e] * voln / count, voln = sum/k
  const double voln = s / (double)K;
  if (!P.strict) {
    // sum_e (rho_e voln / count) regrouped as voln * mean(rho_e): same value up to rounding
    d.mdiag[n] = rbn * voln;
  } else {
    const int cnt = d.nodel_count[n];
    double mass = 0.0;
    if ((cnt & (cnt - 1)) == 0) { // x / 2^m == x * 2^-m exactly: bit-identical to the division
      const double inv = 1.0 / (double)cnt;
      for (int j = 0; j < width; j++) {
        int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
        if (slot >= 0) mass += 1.0 * d.rho[slot / K] * voln * inv;
      }
    } else {
      for (int j = 0; j < width; j++) {
        int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
        if (slot >= 0) mass += 1.0 * d.rho[slot / K] * voln / cnt;
      }
    }
    d.mdiag[n] = mass;
  }
}

template <int K, int MINB = 1, bool HALO = false>
__global__ void __launch_bounds__(TPB_N, MINB) k_node_vol(WfDev d, WfPar P, int mode_in) {
  node_vol_body<K, HALO>(d, P, mode_in, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// E2: the main element pass
// ---------------------------------------------------------------------------------------------
template <int ET>
WF_DI double elem_pressure(const WfDev &d, const WfPar &P, int e, const double (&np_)[Elem<ET>::K], double vol,
                           double vol0, double rho_e, const double (&dH)[Elem<ET>::D][Elem<ET>::K],
                           const double (&vl)[Elem<ET>::K][Elem<ET>::D], bool is_contact = false) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  if (P.press == 0) {
    if constexpr (D == 3) {
      double J_avg = 0.0;
#pragma unroll
      for (int a = 0; a < K; a++) J_avg += np_[a];
      J_avg /= (double)K;
      double div_v = 0.0;
      if (!P.stab_simple) {
#pragma unroll
        for (int a = 0; a < K; a++) div_v += dH[0][a] * vl[a][0] + dH[1][a] * vl[a][1] + dH[2][a] * vl[a][2];
      }
      return pressure_default3d(P, J_avg, vol0, vol, rho_e, div_v, is_contact);
    } else {
      return P.Kbulk * (1.0 - vol / vol0); // calcElemPressureLocal, Mechanical.C:1165-1170
    }
  } else if (P.press == 1) { // as shipped: p += sum pn ; p *= 0.25 k  (Mechanical.C:1243-1247)
    double pe = d.p[e];
#pragma unroll
    for (int a = 0; a < K; a++) pe += np_[a];
    pe *= 0.25 * K;
    return pe;
  } else { // ANP_Nodal (Mechanical.C:1287-1294)
    double pe = 0.0;
#pragma unroll
    for (int a = 0; a < K; a++) pe += np_[a];
    pe /= (double)K;
    return pe;
  }
}
// Mechanical.C:729-747: the element counts as "in contact" when any of its nodes carries a non-zero contact force
template <int ET>
WF_DI bool elem_in_contact(const WfDev &d, int e) {
  bool c = false;
  if (Elem<ET>::D == 3 && d.cflag) {
#pragma unroll
    for (int a = 0; a < Elem<ET>::K; a++) c = c || d.cflag[__ldg(d.elnod + (long long)a * d.ep + e)] != 0;
  }
  return c;
}
template <int ET>
WF_DI void gather_nodal_p(const WfDev &d, const int (&nid)[Elem<ET>::K], double (&np_)[Elem<ET>::K]) {
#pragma unroll
  for (int a = 0; a < Elem<ET>::K; a++) np_[a] = d.nodal_p[nid[a]];
}

// mode bits: 1 = hourglass force kept separate in f_elem_hg (strict two-pass assembly)
// STAGED: the CTA first loads x, v and the nodal ratio of its UNIQUE nodes into shared memory (WfDev::blk_off),
// then every element reads its nodes through 16-bit block-local indices; otherwise every element gathers its own.
// TILE: tile-reduced forces (WfDev::ftile, pull form: see WfDev::tf_tab) instead of one record per element node
template <int ET, bool SEPARATE_HG, bool STAGED, bool THERMAL, bool TILE, bool LOOKAHEAD = TILE>
WF_DI void elem_main_body(const WfDev &d, const WfPar &P, int stride, int vbx) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  static_assert(!(STAGED && THERMAL), "the thermal terms gather by global node id");
  extern __shared__ double sm[];
  pdl_trigger();
  int e = vbx * blockDim.x + threadIdx.x;
  double xl[K][D], vl[K][D], npn[K], A[D][D], dH[D][K], detJ;
  int nid[K];
  if constexpr (STAGED) pdl_wait();
  if constexpr (STAGED) {
    static_assert(TPB_E == WF_EBLK, "block node tables are built for WF_EBLK elements per CTA");
    const int b = vbx;
    const int u0 = __ldg(d.blk_off + b), U = __ldg(d.blk_off + b + 1) - u0;
    for (int i = threadIdx.x; i < U; i += TPB_E) {
      const int g = __ldg(d.blk_nodes + u0 + i);
#pragma unroll
      for (int c = 0; c < D; c++) {
        sm[c * stride + i] = d.x[(long long)c * d.np + g];
        sm[(D + c) * stride + i] = d.v[(long long)c * d.np + g];
      }
      sm[2 * D * stride + i] = d.nodal_p[g];
    }
    unsigned li[K];
    const int ee = e < d.ne ? e : d.ne - 1;
#pragma unroll
    for (int n = 0; n < K; n++) li[n] = d.lidx[(long long)n * d.ep + ee];
    __syncthreads();
    if (e >= d.ne) return;
#pragma unroll
    for (int n = 0; n < K; n++) {
#pragma unroll
      for (int c = 0; c < D; c++) {
        xl[n][c] = sm[c * stride + li[n]];
        vl[n][c] = sm[(D + c) * stride + li[n]];
      }
      npn[n] = sm[2 * D * stride + li[n]];
    }
  } else {
    if (e >= d.ne) return;
    if constexpr (TILE) { // the tile's incidence table is read at the very end: ask L2 for it now (one lane per 32 B sector)
      const int lane = threadIdx.x & 31;
      if (lane * 32 < d.tf_tpitch) prefetch_l2(d.tf_tab + (long long)(e >> 5) * d.tf_tpitch + lane * 32);
    }
    load_conn<ET>(d, e, nid);
    if constexpr (LOOKAHEAD) {
      // connectivity of the CTA that will follow this one on the SM (about five resident CTAs): the gathers of a CTA begin
      // with its node ids, so ask L2 for them one CTA lifetime ahead (K rows of 128 ints = 4 lines each).  Only in the
      // instantiations that were measured with it: in the quadrilateral kernel the two extra registers (128 -> 130)
      // cost the fourth resident CTA (E2 0.081 -> 0.102 ms on 1 M quads), and capped at four CTAs it is 0.086 ms
      if (threadIdx.x < 4 * K) {
        const long long ne0 = ((long long)vbx + 5LL * d.sm_count) * TPB_E + (threadIdx.x & 3) * 32;
        if (ne0 < d.ne) prefetch_l2(d.elnod + (long long)(threadIdx.x >> 2) * d.ep + ne0);
      }
    }
    pdl_wait();
    gather_nodal<ET>(d.x, d.np, nid, xl);
    gather_nodal<ET>(d.v, d.np, nid, vl);
    gather_nodal_p<ET>(d, nid, npn);
  }
  // independent loads issued early
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; i++) tau[i] = d.tau[(long long)i * d.ep + e];
  double pl = d.pl_strain[e];
  const double rho_e = d.rho[e];
  const double vol0 = d.vol_0[e];
  const double sy_prev = d.sigma_y[e];

  jac_adj_det<ET>(xl, A, detJ);
  double radius = 0.0;
  if (D == 2 && d.domtype == 2) radius = elem_radius<ET>(xl);
  const double vol = elem_volume<ET>(detJ, radius, d.domtype, d.vol_weight);
  shape_derivs<ET>(A, dH);
  double Dr[6], Wr[3];
  strain_rates<ET>(dH, detJ, vl, radius, d.domtype, Dr, Wr);
  double temp_e = P.temp;
  if constexpr (THERMAL) {
    // calcThermalExpansion (Thermal.C:151-166): D -= exp_T * dTdt_gp * I with dTdt_gp read from the previous step's
    // flat m_dTedt at the NODE ids, as the reference does
    const double *__restrict__ prev = d.dtedt_low[P.dtedt_cur ^ 1];
    double dTdt_gp = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++) dTdt_gp += 1.0 / K * prev[nid[i]];
    const double f = P.exp_T * dTdt_gp;
    Dr[0] = Dr[0] - f * 1.; Dr[1] = Dr[1] - f * 1.; Dr[2] = Dr[2] - f * 1.;
    Dr[3] = Dr[3] - f * 0.; Dr[4] = Dr[4] - f * 0.; Dr[5] = Dr[5] - f * 0.;
    const int eu = d.e_user ? d.e_user[e] : e; // the reference's element id
    if (eu < d.nn) temp_e = d.T[eu]; // T[e]: nodal array read with the element id (Mechanical.C:1731)
  }
  const double p = elem_pressure<ET>(d, P, e, npn, vol, vol0, rho_e, dH, vl, elem_in_contact<ET>(d, e));
  StressOut so;
  stress_update(P, P.dt, p, Dr, Wr, tau, pl, sy_prev, so, temp_e);
  if constexpr (THERMAL) {
    // m_q_plheat (Mechanical.C:1784-1818) and ThermalCalcs' element part (Thermal.C:41-93): conduction Kt T + plastic
    // heating, one value per element node, stored where the node's nodel list expects it (tsell, same order as fsell)
    const double qpl = plastic_heat(P, P.dt, so);
    d.q_plheat[e] = qpl;
    double Te[K], md[K];
#pragma unroll
    for (int i = 0; i < K; i++) { Te[i] = d.T[nid[i]]; md[i] = d.mdiag[nid[i]]; }
    const double w = gauss_w<ET>();
    const double heat = 0.9 * qpl;
    const double elem_pow = heat * vol;
    const double pow_per_node = elem_pow / K;
    double *__restrict__ cur = d.dtedt_low[P.dtedt_cur];
#pragma unroll
    for (int i = 0; i < K; i++) {
      double dTde = 0.0;
#pragma unroll
      for (int j = 0; j < K; j++) {
        double kk = 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) kk += dH[c][i] * dH[c][j];
        dTde += (kk * P.k_T / detJ * w) * Te[j];
      }
      const double m_inv = 1.0 / md[i];
      double val = -m_inv * dTde;
      val += pow_per_node / (md[i] * P.cp_T);
      const long long o = (long long)__ldg(d.pos + (long long)i * d.ep + e);
      const long long lane = o & 31;
      d.tsell[(o + (D - 1) * lane) / D] = val; // fsell offset D*q - (D-1)*lane  ->  q
      const long long flat = (long long)(d.e_user ? d.e_user[e] : e) * K + i; // index in the reference's flat m_dTedt
      if (flat < d.nn) cur[flat] = val;
    }
  }
  if (P.track_eps) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      long long o = (long long)i * d.ep + e;
      d.eps[o] = d.eps[o] + P.dt * Dr[i];
    }
  }
  if (P.av_alpha != 0.0 || P.av_beta != 0.0) artificial_viscosity(P, Dr, rho_e, vol, so.sig);
#pragma unroll
  for (int i = 0; i < 6; i++) d.tau[(long long)i * d.ep + e] = tau[i];
  if (P.store_sigma) {
#pragma unroll
    for (int i = 0; i < 6; i++) d.sigma[(long long)i * d.ep + e] = so.sig[i];
  }
  d.pl_strain[e] = pl;
  d.sigma_y[e] = so.sy;
  d.p[e] = p;

  double f[K][D];
  elem_forces<ET>(dH, so.sig, detJ, radius, d.domtype, d.vol_weight, f);
  // hourglass control
  double fh[K][D];
  bool have_hg = false;
  if constexpr (ET == ET_HEX8) {
    if (P.hexa_hg != 0.0) { hexa_hourglass(P, vl, vol, rho_e, fh); have_hg = true; }
  } else if constexpr (ET == ET_QUAD4) {
    double q[2] = {d.hg_q[e], d.hg_q[d.ep + e]};
    quad_hourglass(P, vl, vol, rho_e, q, fh);
    d.hg_q[e] = q[0];
    d.hg_q[d.ep + e] = q[1];
    have_hg = true;
  }
  if (have_hg && !SEPARATE_HG) {
#pragma unroll
    for (int n = 0; n < K; n++)
#pragma unroll
      for (int c = 0; c < D; c++) f[n][c] -= fh[n][c];
  }
  if constexpr (TILE) {
    static_assert(!SEPARATE_HG && !STAGED && !THERMAL, "tile-reduced forces: fused hourglass, per-element gathers only");
    // the warp's 32 elements drop their nodal forces in shared memory; lane u then adds up, in the fixed order of
    // the tile's incidence table, the entries of unique node u and writes ONE partial per (tile, node)
    constexpr int KD = K * D;
    const unsigned amask = __activemask(); // a tail tile has fewer than 32 elements (lanes 0 .. nact-1)
    const int nact = __popc(amask), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ws = d.tf_stride, tp = d.tf_tpitch;
    const long long tile = e >> 5;
    double *fb = sm + warp * (KD * 32 + (tp + 7) / 8);
    unsigned char *tab = reinterpret_cast<unsigned char *>(fb + KD * 32);
    const unsigned *__restrict__ gt = reinterpret_cast<const unsigned *>(d.tf_tab + tile * tp);
    for (int i = lane; i < tp / 4; i += nact) reinterpret_cast<unsigned *>(tab)[i] = __ldg(gt + i);
#pragma unroll
    for (int n = 0; n < K; n++)
#pragma unroll
      for (int c = 0; c < D; c++) fb[(n * D + c) * 32 + lane] = f[n][c];
    __syncwarp(amask);
    double *__restrict__ out = d.ftile + tile * D * ws;
    for (int u = lane; u < ws; u += nact) {
      const int q0 = tab[u], q1 = tab[u + 1];
      double sacc[D];
#pragma unroll
      for (int c = 0; c < D; c++) sacc[c] = 0.0;
      for (int q = q0; q < q1; q++) {
        const int slot = tab[ws + 1 + q];
        const double *src = fb + (slot % K) * D * 32 + slot / K;
#pragma unroll
        for (int c = 0; c < D; c++) sacc[c] += src[c * 32];
      }
#pragma unroll
      for (int c = 0; c < D; c++) out[(long long)c * ws + u] = sacc[c];
    }
    return;
  }
  // node-ordered stores: entry of (e, n) in the nodel list of its node
#pragma unroll
  for (int n = 0; n < K; n++) {
    const long long o = (long long)__ldg(d.pos + (long long)n * d.ep + e);
#pragma unroll
    for (int c = 0; c < D; c++) d.fsell[o + 32 * c] = f[n][c];
    if (SEPARATE_HG && have_hg) {
#pragma unroll
      for (int c = 0; c < D; c++) d.fsell_hg[o + 32 * c] = fh[n][c];
    }
  }
}

template <int ET, bool SEPARATE_HG, bool STAGED, bool THERMAL = false, bool TILE = false, int MINB = 1, bool LOOKAHEAD = TILE>
__global__ void __launch_bounds__(TPB_E, MINB) k_elem_main(WfDev d, WfPar P, int stride) {
  elem_main_body<ET, SEPARATE_HG, STAGED, THERMAL, TILE, LOOKAHEAD>(d, P, stride, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// device helpers of the multi-GPU halo exchange (the primitives are defined ahead of the node passes)
// ---------------------------------------------------------------------------------------------
// sum of the sharers' partials of unique shared node u, component comp, ascending rank order
WF_DI double halo_total(const WfDev &d, int u, int comp, int parity, double own) {
  double acc = 0.0;
  const int q1 = d.hu_ptr[u + 1];
  for (int q = d.hu_ptr[u]; q < q1; q++) {
    const int2 en = d.hu_ent[q];
    acc += (en.x < 0) ? own : __ldcg(d.recv + en.x + (long long)(parity * WF_HALO_NC + comp) * en.y);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// N2: per-node force sum in nodel order (assemblyForces, Matrices.C:42-87: element forces first,
// then hourglass forces subtracted), calcAccel (Mechanical.C:321-341), ImposeBCA,
// UpdateCorrectionAccVel (Domain_d.C:981-997), ImposeBCV, axis constraint
// (Solver_explicit.C:953-969), UpdateCorrectionPos (Domain_d.C:1005-1025) and, unless this is the
// last step of the batch, the next step's UpdatePrediction + ImposeBCV.  The nodal mass was formed
// by N1.  One warp == one slice of 32 nodes; every load is a contiguous 256 B row.
//   phase 0 = everything;  phase 1 = sums only, to d.fi (lazy m_fi);  phase 2 = integrate from d.fi;
//   phase 3 = everything, nodes shared with another rank skipped;  phase 4 = everything, shared nodes only (multi-GPU:
//   the bulk of the pass runs while the force partials of the neighbours are still in flight).
// On a partitioned mesh the sums of shared nodes are completed with the neighbours' partials (halo_total).
// ---------------------------------------------------------------------------------------------
// sum of x over the threads of the warp that are executing this call (any subset: the node pass retires lanes early),
// added to *dst with one atomic; the order of the atomics is not fixed, the value feeds the step monitor only
WF_DI void warp_add(double *dst, double x) {
  const unsigned mask = __activemask();
  const int lane = threadIdx.x & 31;
  double t;
  if (mask == 0xffffffffu) {
    t = x;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(mask, t, o);
  } else {
    t = 0.0;
    for (unsigned m = mask; m; m &= m - 1) t += __shfl_sync(mask, x, __ffs(m) - 1);
  }
  if (lane == __ffs(mask) - 1) atomicAdd(dst, t);
}
// ekin_acc: where the kinetic energy of the corrected velocities is accumulated (step monitor), or NULL
// PHASE >= 0: the phase is known at compile time (the partitioned-mesh launches: phase 3 carries the send and no shared
// node, phase 4 the wait and only shared nodes; each instantiation drops the other's code and registers)
template <int D, bool SEPARATE_HG, int UNROLL, bool TILE_F, bool PREFETCH, bool HALO, int PHASE = -1>
WF_DI void node_update_body(const WfDev &d, const WfPar &P, int fuse_flags, int phase_in, int vbx, double *ekin_acc) {
  const int phase = PHASE >= 0 ? PHASE : phase_in;
  const bool fuse_predictor = fuse_flags & 1, udt_recompute = fuse_flags & 2, udt_skip_store = fuse_flags & 4;
  // a launch that begins by spinning on the neighbours' flags (phase 4) lets its dependents in only after the wait: their
  // resident CTAs would otherwise hold SM slots that another rank of a single-process cluster may need to get its sends out
  if (!(HALO && phase == 4)) pdl_trigger();
  int bx = vbx;
  if (HALO && phase == 3 && P.send_ctas > 0) {
    // multi-GPU: the first CTAs of the launch send this rank's partial forces of the shared nodes; the rest of the grid
    // integrates the nodes this rank does not share while they travel
    if (bx < P.send_ctas) {
      pdl_wait();
      halo_send_cta<2>(d, P, TILE_F ? 2 : (SEPARATE_HG ? 1 : 0), P.send_seq, bx / P.send_chunks, bx % P.send_chunks, P.send_chunks);
      return;
    }
    bx -= P.send_ctas;
  }
  if (HALO && phase == 4) {
    if (P.wait_seq) halo_wait_cta(d, P.wait_seq, P.wait_timeout_ns); // folded k_halo_wait (before any early exit)
    pdl_trigger();
  }
  int n = bx * blockDim.x + threadIdx.x;
  if (phase == 4) { // shared nodes only, one thread per unique shared node (after the halo wait)
    if (n >= d.n_uniq) return;
    n = d.hu_node[n];
  }
  int slice = n >> 5;
  if (slice >= d.nslices) return;
  // phase 3: shared nodes wait for phase 4.  The flag is only LOADED here and tested after the force gathers (which have
  // no side effects), so that its latency overlaps theirs instead of heading every thread's dependency chain
  int hs3 = -1;
  if (HALO && phase == 3 && d.halo_slot && n < d.nn) hs3 = __ldg(d.halo_slot + n);
  const int lane = n & 31;
  double fi[D];
#pragma unroll
  for (int c = 0; c < D; c++) fi[c] = 0.0;
  if (PREFETCH && phase != 1 && n < d.nn) {
    // the node's own state is needed only after the force sum: ask L2 for it now, without holding registers
    // (loading it early instead costs 32 registers and a third of the occupancy: measured slower)
    prefetch_l2(d.mdiag + n);
    prefetch_l2(d.bc_index + n);
#pragma unroll
    for (int c = 0; c < D; c++) {
      const long long i = (long long)c * d.np + n;
      prefetch_l2(d.prev_a + i); prefetch_l2(d.v + i); prefetch_l2(d.x + i); prefetch_l2(d.u + i);
      if (!udt_recompute) prefetch_l2(d.u_dt + i);
    }
  }
  pdl_wait();
  if (TILE_F && phase != 2) {
    // tile-reduced forces: one partial per tile that touches the node, gathered through the tile-entry table
    const long long base = d.tf_ptr[slice];
    const int width = (int)((d.tf_ptr[slice + 1] - base) >> 5);
    const unsigned *__restrict__ sl = d.tf_slots + base + lane;
    const long long cs = d.tf_stride;
    for (int j0 = 0; j0 < width; j0 += UNROLL) {
      unsigned o[UNROLL];
      double fv[UNROLL][D];
#pragma unroll
      for (int q = 0; q < UNROLL; q++) o[q] = (j0 + q < width) ? __ldg(sl + ((long long)(j0 + q) << 5)) : 0xFFFFFFFFu;
#pragma unroll
      for (int q = 0; q < UNROLL; q++)
#pragma unroll
        for (int c = 0; c < D; c++) fv[q][c] = (o[q] != 0xFFFFFFFFu) ? d.ftile[(long long)o[q] + c * cs] : 0.0;
#pragma unroll
      for (int q = 0; q < UNROLL; q++)
#pragma unroll
        for (int c = 0; c < D; c++) fi[c] += fv[q][c];
    }
  } else if (phase != 2) {
    const long long base = d.sell_ptr[slice];
    const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
    const double *__restrict__ row = d.fsell + base * D + lane;
    // padding entries hold +0.0 and are never written, so summing the full width is exact
    for (int j0 = 0; j0 < width; j0 += UNROLL) {
      double fv[UNROLL][D];
#pragma unroll
      for (int q = 0; q < UNROLL; q++)
#pragma unroll
        for (int c = 0; c < D; c++) fv[q][c] = (j0 + q < width) ? row[((long long)(j0 + q) * D + c) * 32] : 0.0;
#pragma unroll
      for (int q = 0; q < UNROLL; q++)
#pragma unroll
        for (int c = 0; c < D; c++) fi[c] += fv[q][c];
    }
    if (SEPARATE_HG) {
      const double *__restrict__ rowh = d.fsell_hg + base * D + lane;
      for (int j = 0; j < width; j++)
#pragma unroll
        for (int c = 0; c < D; c++) fi[c] -= rowh[((long long)j * D + c) * 32];
    }
  }
  if (n >= d.nn || hs3 >= 0) return;
  if (d.halo_slot && phase != 2 && phase != 3) { // shared node (phase 3 has none): add the other sharers' partials, ascending rank order
    const int u = d.halo_slot[n];
    if (u >= 0) {
#pragma unroll
      for (int c = 0; c < D; c++) fi[c] = halo_total(d, u, c, P.halo_parity, fi[c]);
    }
  }
  if (phase == 1) {
#pragma unroll
    for (int c = 0; c < D; c++) d.fi[(long long)c * d.np + n] = fi[c];
    return;
  }
  if (phase == 2) {
#pragma unroll
    for (int c = 0; c < D; c++) fi[c] = d.fi[(long long)c * d.np + n];
  }
  const double mass = d.mdiag[n];
  // non-finite scrub (Solver_explicit.C:779-784)
#pragma unroll
  for (int c = 0; c < D; c++)
    if (!isfinite(fi[c])) { fi[c] = 0.0; *d.nonfinite = 1; }

  int bi = d.bc_index[n];
  unsigned bm = (bi >= 0) ? d.bc_mask[bi] : 0u;
  const double f = 1.0 / (1.0 - P.alpha);
  double a[D], v[D], udt0[D];
#pragma unroll
  for (int c = 0; c < D; c++) {
    long long i = (long long)c * d.np + n;
    double fe = d.fe ? d.fe[i] : 0.0;
    a[c] = (fe - fi[c]) / mass;
    if (d.contforce) a[c] += d.contforce[i] / mass; // calcAccel with contact, Mechanical.C:330-335
    if (bm & (1u << c)) a[c] = 0.0;
    double pa = d.prev_a[i];
    const double vp = d.v[i];
    // the fused predictor of the previous step formed u_dt = dt (v_c + (1/2 - beta) dt a) and v_p = v_c + (1 - gamma) dt a
    // (a = 0 and v_p = v_c = the prescribed value on constrained components), hence u_dt = dt (v_p + (gamma - 1/2 - beta) dt a)
    // (a prescribed component may get a NEW value between two steps, wf_set_bc_values: its u_dt = dt * OLD value is
    // therefore always stored and read, never recomputed from the patched velocity)
    udt0[c] = (udt_recompute && !(bm & (1u << c))) ? P.dt * (vp + (P.gamma - 0.5 - P.beta) * P.dt * pa) : d.u_dt[i];
    a[c] = f * (a[c] - P.alpha * pa);
    v[c] = vp + P.gamma * P.dt * a[c];
    if (bm & (1u << c)) v[c] = d.bc_vals[3 * bi + c];
  }
  double xr = d.x[n];
  if (d.domtype == 2) {
    double xmin = key_dbl(d.xmin_key[P.xmin_cur]);
    if (xr <= xmin + 1.e-6) { a[0] = 0.0; v[0] = 0.0; }
  }
  if (ekin_acc) { // step monitor: kinetic energy of the corrected velocities (computeEnergies, Mechanical.C:2145)
    double s2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; c++) s2 += v[c] * v[c];
    warp_add(ekin_acc + (bx & 255), 0.5 * mass * s2); // 256 accumulators (wf_engine::MON_NACC), summed by the host
  }
#pragma unroll
  for (int c = 0; c < D; c++) {
    long long i = (long long)c * d.np + n;
    double udt = udt0[c] + P.beta * P.dt * P.dt * a[c];
    double xn = d.x[i] + udt;
    d.x[i] = xn;
    if (c == 0) xr = xn;
    d.prev_a[i] = a[c];
    d.u[i] = d.u[i] + udt;
    if (fuse_predictor) {
      if (!udt_skip_store || (bm & (1u << c))) d.u_dt[i] = P.dt * (v[c] + (0.5 - P.beta) * P.dt * a[c]);
      v[c] = v[c] + (1.0 - P.gamma) * P.dt * a[c];
      if (bm & (1u << c)) v[c] = d.bc_vals[3 * bi + c];
    } else {
      d.u_dt[i] = udt;
    }
    d.v[i] = v[c];
  }
  if (d.domtype == 2) { // one atomic per warp (the 2D node count would otherwise serialise on one address)
    unsigned long long key = dbl_key(xr);
    const unsigned mask = __activemask();
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_down_sync(mask, key, o);
      const int src = (threadIdx.x & 31) + o;
      if (src < 32 && ((mask >> src) & 1u) && other < key) key = other;
    }
    if ((threadIdx.x & 31) == (__ffs(mask) - 1)) atomicMin(d.xmin_key + (P.xmin_cur ^ 1), key);
  }
}

template <int D, bool SEPARATE_HG, int UNROLL, bool TILE_F = false, bool PREFETCH = false, int MINB = 1, bool HALO = false,
          int PHASE = -1>
__global__ void __launch_bounds__(TPB_N, MINB) k_node_update(WfDev d, WfPar P, int fuse_flags, int phase) {
  node_update_body<D, SEPARATE_HG, UNROLL, TILE_F, PREFETCH, HALO, PHASE>(d, P, fuse_flags, phase, blockIdx.x, d.ekin_acc);
}

// ThermalCalcs, node part (Thermal.C:103-125): dTdt = sum of the element contributions in nodel order;
// T += (dTdt + q_cont_conv / (m cp)) dt.  One warp == one SELL slice, rows of tsell are contiguous.
__global__ void __launch_bounds__(TPB_N) k_node_thermal(WfDev d, WfPar P) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = n >> 5;
  if (slice >= d.nslices) return;
  const long long base = d.sell_ptr[slice];
  const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
  const double *__restrict__ row = d.tsell + base + (n & 31);
  double dTdt = 0;
  for (int j = 0; j < width; j++) dTdt += row[(long long)j * 32]; // padding entries stay +0.0
  if (n >= d.nn) return;
  const double qc = d.q_cont_conv ? d.q_cont_conv[n] : 0.0;
  d.T[n] += (dTdt + qc * 1.0 / (d.mdiag[n] * P.cp_T)) * P.dt;
}

// nodal mass only (init, unfused CalcNodalVol + CalcNodalMassFromVol, lazy m_mdiag)
template <int K>
__global__ void __launch_bounds__(TPB_N) k_node_mass(WfDev d, WfPar P, int use_stored_voln) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = n >> 5;
  if (slice >= d.nslices) return;
  const long long base = d.sell_ptr[slice];
  const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
  const int lane = n & 31;
  const bool valid = n < d.nn;
  const int cnt = valid ? d.nodel_count[n] : 1;
  const double voln = valid ? (use_stored_voln ? d.voln[n] : d.voln_sum[n] / (double)K) : 0.0;
  double mass = 0.0;
  for (int j = 0; j < width; j++) {
    int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
    if (slot >= 0) mass += 1.0 * d.rho[slot / K] * voln / cnt;
  }
  if (valid) d.mdiag[n] = mass;
}

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_init_elem(WfDev d, WfPar P) { // InitValues (Domain_d.C:393-415)
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  d.pl_strain[e] = 0.0;
  d.sigma_y[e] = P.sy0;
}
__global__ void k_vol0_density(WfDev d) { // CalcElemInitialVol (:343) + calcElemDensity (:295)
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  double v = d.vol[e];
  d.vol_0[e] = v;
  d.rho[e] = d.rho_0[e] * v / v;
}
__global__ void k_density(WfDev d) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  d.rho[e] = d.rho_0[e] * d.vol_0[e] / d.vol[e];
}
__global__ void k_xmin(WfDev d, int slot) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
  atomicMin(d.xmin_key + slot, dbl_key(d.x[n]));
}
// sigma = -p I + tau, rebuilt on request when it is not stored per step (Mechanical.C:1775)
__global__ void k_rebuild_sigma(WfDev d, double *out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  double mp = -d.p[e];
#pragma unroll
  for (int i = 0; i < 6; i++) out[(long long)i * d.ep + e] = mp * (i < 3 ? 1. : 0.) + d.tau[(long long)i * d.ep + e];
}

// sum over the CTA (blockDim.x a multiple of 32, <= 1024); result valid in thread 0
WF_DI double block_sum(double k) {
  __shared__ double part[32];
  for (int o = 16; o > 0; o >>= 1) k += __shfl_down_sync(0xffffffffu, k, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x < 32) {
    k = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) k += __shfl_down_sync(0xffffffffu, k, o);
  }
  return k;
}

// computeEnergies (Mechanical.C:2145-2185): block-reduced, then one double atomic per CTA (diagnostic only)
template <int D>
__global__ void k_energy_kin(WfDev d) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  double k = 0.0;
  if (n < d.nn) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; c++) { double vv = d.v[(long long)c * d.np + n]; s += vv * vv; }
    k = 0.5 * d.mdiag[n] * s;
  }
  k = block_sum(k);
  if (threadIdx.x == 0 && k != 0.0) atomicAdd(d.red + 0, k);
}
__global__ void k_energy_int(WfDev d, const double *sig) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  double k = 0.0;
  if (e < d.ne) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) s += sig[(long long)i * d.ep + e] * d.str_rate[(long long)i * d.ep + e];
    k = s * d.vol[e];
  }
  k = block_sum(k);
  if (threadIdx.x == 0 && k != 0.0) atomicAdd(d.red + 1, k);
}


// ---------------------------------------------------------------------------------------------
// diagnostics and output on the device
// ---------------------------------------------------------------------------------------------
// calcNodalPressureFromElemental (Mechanical.C:1187-1212): p_node = sum p[e] vol[e] / sum vol[e]; the reference
// scatters in ascending element order, which is the order of the node's nodel list
template <int K>
__global__ void __launch_bounds__(TPB_N) k_p_node(WfDev d, double *__restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = n >> 5;
  if (slice >= d.nslices) return;
  const long long base = d.sell_ptr[slice];
  const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
  const int lane = n & 31;
  double pv = 0.0, acc = 0.0;
  for (int j = 0; j < width; j++) {
    int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
    if (slot >= 0) {
      const int e = slot / K;
      const double ve = d.vol[e];
      pv += d.p[e] * ve;
      acc += ve;
    }
  }
  if (n >= d.nn) return;
  if (acc > 0.0) pv /= acc;
  out[n] = pv;
}

WF_DI double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
WF_DI double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// calcMinEdgeLength (Domain_d.C:2224-2468): min edge length / min height; 3D elements are read as the tetrahedron of
// their first four nodes (:2243-2247), 2D elements as quadrilaterals (:2381-2413).  keys[0] = min length,
// keys[1] = min height as ordered keys (min is order-independent, so the result is the reference's bit for bit in
// the strict flavour).  Values start at 1.0e6 like the reference.
template <int D>
__global__ void __launch_bounds__(128) k_min_edge(WfDev d, double *__restrict__ elem_length, unsigned long long *keys) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double min_len = 1.0e6, eh = 1.0e6;
  bool have_h = false;
  if (e < d.ne) {
    double P[4][D];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int g = __ldg(d.elnod + (long long)i * d.ep + e);
#pragma unroll
      for (int c = 0; c < D; c++) P[i][c] = d.x[(long long)c * d.np + g];
    }
    if constexpr (D == 3) {
      const int ed[6][2] = {{1, 0}, {2, 0}, {3, 0}, {2, 1}, {3, 1}, {3, 2}};
#pragma unroll
      for (int i = 0; i < 6; i++) {
        const double a = P[ed[i][0]][0] - P[ed[i][1]][0], b = P[ed[i][0]][1] - P[ed[i][1]][1], c = P[ed[i][0]][2] - P[ed[i][1]][2];
        const double len = sqrt(a * a + b * b + c * c);
        if (len < min_len) min_len = len;
      }
      const int fc[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        double x1[3], x2[3], nr[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { x1[c] = P[fc[i][1]][c] - P[fc[i][0]][c]; x2[c] = P[fc[i][2]][c] - P[fc[i][0]][c]; }
        nr[0] = x1[1] * x2[2] - x1[2] * x2[1];
        nr[1] = x1[2] * x2[0] - x1[0] * x2[2];
        nr[2] = x1[0] * x2[1] - x1[1] * x2[0];
        const double area = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
        if (area < 1e-12) continue;
        const double inv = 1.0 / area;
        const double h = fabs((P[i][0] - P[fc[i][0]][0]) * (nr[0] * inv) + (P[i][1] - P[fc[i][0]][1]) * (nr[1] * inv) +
                              (P[i][2] - P[fc[i][0]][2]) * (nr[2] * inv));
        if (h < eh) eh = h;
      }
      have_h = true; // the reference updates min_height for every 3D element, even with the 1.0e6 sentinel
    } else {
      const double *A = P[0], *B = P[1], *C = P[2], *Dd = P[3];
      const double lenAB = sqrt((B[0] - A[0]) * (B[0] - A[0]) + (B[1] - A[1]) * (B[1] - A[1]));
      const double lenBC = sqrt((C[0] - B[0]) * (C[0] - B[0]) + (C[1] - B[1]) * (C[1] - B[1]));
      const double lenCD = sqrt((Dd[0] - C[0]) * (Dd[0] - C[0]) + (Dd[1] - C[1]) * (Dd[1] - C[1]));
      const double lenDA = sqrt((A[0] - Dd[0]) * (A[0] - Dd[0]) + (A[1] - Dd[1]) * (A[1] - Dd[1]));
      min_len = fmin(min_len, fmin(fmin(lenAB, lenBC), fmin(lenCD, lenDA)));
      const double area1 = 0.5 * fabs((B[0] - A[0]) * (C[1] - A[1]) - (C[0] - A[0]) * (B[1] - A[1]));
      const double area2 = 0.5 * fabs((C[0] - A[0]) * (Dd[1] - A[1]) - (Dd[0] - A[0]) * (C[1] - A[1]));
      const double area = area1 + area2;
      if (area > 1e-14) {
        eh = fmin(fmin(2.0 * area / lenAB, 2.0 * area / lenBC), fmin(2.0 * area / lenCD, 2.0 * area / lenDA));
        have_h = true;
      }
    }
    elem_length[e] = eh;
  }
  const double wl = warp_min(min_len), wh = warp_min(have_h ? eh : 1.0e6);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(keys + 0, dbl_key(wl));
    atomicMin(keys + 1, dbl_key(wh));
  }
}

// max |v| over the nodes (Solver_explicit.C:583-587), keys[2]
template <int D>
__global__ void __launch_bounds__(256) k_max_vel(WfDev d, unsigned long long *keys) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (n < d.nn) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < D; c++) { const double vv = d.v[(long long)c * d.np + n]; s += vv * vv; }
    m = sqrt(s);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(keys + 2, dbl_key(m));
}

// layout conversion between the private component-major arrays and the reference's interleaved records:
//   aos[i * nc + c] <-> soa[c * pitch + map(i)]   (scale: m_voln = sum / k)
// map = NULL for nodal arrays; for element arrays map[user element] = internal element (wf_host_elem_order)
__global__ void k_soa_to_aos(const double *__restrict__ soa, long long pitch, int nc, long long n, double scale, double *__restrict__ aos,
                             const int *__restrict__ map) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n * nc) return;
  const long long i = j / nc;
  const int c = (int)(j - i * nc);
  const long long is = map ? (long long)map[i] : i;
  aos[j] = soa[(long long)c * pitch + is] * scale;
}
__global__ void k_aos_to_soa(const double *__restrict__ aos, long long pitch, int nc, long long n, double *__restrict__ soa,
                             const int *__restrict__ map) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n * nc) return;
  const long long i = j / nc;
  const int c = (int)(j - i * nc);
  const long long is = map ? (long long)map[i] : i;
  soa[(long long)c * pitch + is] = aos[j];
}

// ---------------------------------------------------------------------------------------------
// unfused element kernels on stored intermediates (parity bisecting)
// ---------------------------------------------------------------------------------------------
template <int ET>
WF_DI void load_dH(const WfDev &d, int e, double (&dH)[Elem<ET>::D][Elem<ET>::K]) {
#pragma unroll
  for (int c = 0; c < Elem<ET>::D; c++)
#pragma unroll
    for (int n = 0; n < Elem<ET>::K; n++) dH[c][n] = d.dH[((long long)c * Elem<ET>::K + n) * d.ep + e];
}

template <int ET>
__global__ void k_u_strain_rates(WfDev d, WfPar P) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  int nid[K];
  load_conn<ET>(d, e, nid);
  double vl[K][D], dH[D][K], Dr[6], Wr[3];
  gather_nodal<ET>(d.v, d.np, nid, vl);
  load_dH<ET>(d, e, dH);
  strain_rates<ET>(dH, d.detJ[e], vl, d.radius ? d.radius[e] : 0.0, d.domtype, Dr, Wr);
#pragma unroll
  for (int i = 0; i < 6; i++) d.str_rate[(long long)i * d.ep + e] = Dr[i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    d.rot_rate[(long long)i * d.ep + e] = 0.0;
    d.rot_rate[(long long)(3 + i) * d.ep + e] = Wr[i];
  }
}

template <int ET>
__global__ void k_u_pressure(WfDev d, WfPar P) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  int nid[K];
  load_conn<ET>(d, e, nid);
  double vl[K][D], dH[D][K], npn[K];
  gather_nodal<ET>(d.v, d.np, nid, vl);
  gather_nodal_p<ET>(d, nid, npn);
  load_dH<ET>(d, e, dH);
  d.p[e] = elem_pressure<ET>(d, P, e, npn, d.vol[e], d.vol_0[e], d.rho[e], dH, vl, elem_in_contact<ET>(d, e));
}

__global__ void k_u_stress(WfDev d, WfPar P, double dt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  double tau[6], Dr[6], Wr[3];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    tau[i] = d.tau[(long long)i * d.ep + e];
    Dr[i] = d.str_rate[(long long)i * d.ep + e];
  }
#pragma unroll
  for (int i = 0; i < 3; i++) Wr[i] = d.rot_rate[(long long)(3 + i) * d.ep + e];
  double pl = d.pl_strain[e];
  StressOut so;
  stress_update(P, dt, d.p[e], Dr, Wr, tau, pl, d.sigma_y[e], so, P.temp);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    d.tau[(long long)i * d.ep + e] = tau[i];
    d.sigma[(long long)i * d.ep + e] = so.sig[i];
  }
  if (P.track_eps) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      long long o = (long long)i * d.ep + e;
      d.eps[o] = d.eps[o] + dt * Dr[i];
    }
  }
  d.pl_strain[e] = pl;
  d.sigma_y[e] = so.sy;
}

__global__ void k_u_artvisc(WfDev d, WfPar P) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  double Dr[6], sig[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    Dr[i] = d.str_rate[(long long)i * d.ep + e];
    sig[i] = d.sigma[(long long)i * d.ep + e];
  }
  artificial_viscosity(P, Dr, d.rho[e], d.vol[e], sig);
#pragma unroll
  for (int i = 0; i < 3; i++) d.sigma[(long long)i * d.ep + e] = sig[i];
}

template <int ET>
__global__ void k_u_forces(WfDev d, WfPar P) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  double dH[D][K], sig[6], f[K][D];
  load_dH<ET>(d, e, dH);
#pragma unroll
  for (int i = 0; i < 6; i++) sig[i] = d.sigma[(long long)i * d.ep + e];
  elem_forces<ET>(dH, sig, d.detJ[e], d.radius ? d.radius[e] : 0.0, d.domtype, d.vol_weight, f);
#pragma unroll
  for (int n = 0; n < K; n++)
#pragma unroll
    for (int c = 0; c < D; c++) d.f_elem[((long long)n * D + c) * d.ep + e] = f[n][c];
}

template <int ET>
__global__ void k_u_hourglass(WfDev d, WfPar P) {
  constexpr int K = Elem<ET>::K, D = Elem<ET>::D;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.ne) return;
  if constexpr (ET == ET_HEX8 || ET == ET_QUAD4) {
    int nid[K];
    load_conn<ET>(d, e, nid);
    double vl[K][D], fh[K][D];
    gather_nodal<ET>(d.v, d.np, nid, vl);
    if constexpr (ET == ET_HEX8) {
      if (P.hexa_hg == 0.0) return;
      hexa_hourglass(P, vl, d.vol[e], d.rho[e], fh);
    } else {
      double q[2] = {d.hg_q[e], d.hg_q[d.ep + e]};
      quad_hourglass(P, vl, d.vol[e], d.rho[e], q, fh);
      d.hg_q[e] = q[0];
      d.hg_q[d.ep + e] = q[1];
    }
#pragma unroll
    for (int n = 0; n < K; n++)
#pragma unroll
      for (int c = 0; c < D; c++) d.f_elem_hg[((long long)n * D + c) * d.ep + e] = fh[n][c];
  }
}

// unfused node kernels ------------------------------------------------------------------------
template <int K>
__global__ void k_u_nodal_vol(WfDev d) { // CalcNodalVol
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = n >> 5;
  if (slice >= d.nslices) return;
  const long long base = d.sell_ptr[slice];
  const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
  const int lane = n & 31;
  double s = 0.0;
  for (int j = 0; j < width; j++) {
    int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
    if (slot >= 0) s += d.vol[slot / K];
  }
  if (n < d.nn) { d.voln_sum[n] = s; d.voln[n] = s / (double)K; }
}

template <int K, int D>
__global__ void k_u_assembly(WfDev d) { // assemblyForces + scrub
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = n >> 5;
  if (slice >= d.nslices) return;
  const long long base = d.sell_ptr[slice];
  const int width = (int)((d.sell_ptr[slice + 1] - base) >> 5);
  const int lane = n & 31;
  double fi[D];
#pragma unroll
  for (int c = 0; c < D; c++) fi[c] = 0.0;
  for (int j = 0; j < width; j++) {
    int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
    if (slot >= 0) {
      int e = slot / K, ln = slot - e * K;
#pragma unroll
      for (int c = 0; c < D; c++) fi[c] += d.f_elem[((long long)ln * D + c) * d.ep + e];
    }
  }
  for (int j = 0; j < width; j++) {
    int slot = __ldg(d.sell_slots + base + ((long long)j << 5) + lane);
    if (slot >= 0) {
      int e = slot / K, ln = slot - e * K;
#pragma unroll
      for (int c = 0; c < D; c++) fi[c] -= d.f_elem_hg[((long long)ln * D + c) * d.ep + e];
    }
  }
  if (n >= d.nn) return;
#pragma unroll
  for (int c = 0; c < D; c++) {
    if (!isfinite(fi[c])) { fi[c] = 0.0; *d.nonfinite = 1; }
    d.fi[(long long)c * d.np + n] = fi[c];
  }
}

template <int D>
__global__ void k_u_accel(WfDev d) { // calcAccel
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
#pragma unroll
  for (int c = 0; c < D; c++) {
    long long i = (long long)c * d.np + n;
    double fe = d.fe ? d.fe[i] : 0.0;
    d.a[i] = (fe - d.fi[i]) / d.mdiag[n];
    if (d.contforce) d.a[i] += d.contforce[i] / d.mdiag[n]; // Mechanical.C:330-335
  }
}
template <int D>
__global__ void k_u_corr_accvel(WfDev d, WfPar P) { // UpdateCorrectionAccVel
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
  const double f = 1.0 / (1.0 - P.alpha);
#pragma unroll
  for (int c = 0; c < D; c++) {
    long long i = (long long)c * d.np + n;
    double a = f * (d.a[i] - P.alpha * d.prev_a[i]);
    d.a[i] = a;
    d.v[i] = d.v[i] + P.gamma * P.dt * a;
  }
}
__global__ void k_u_axis(WfDev d, WfPar P) { // axis constraint with xmin of current x
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
  double xmin = key_dbl(d.xmin_key[P.xmin_cur]);
  if (d.x[n] <= xmin + 1.e-6) { d.a[n] = 0.0; d.v[n] = 0.0; }
}
template <int D>
__global__ void k_u_corr_pos(WfDev d, WfPar P) { // UpdateCorrectionPos
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.nn) return;
#pragma unroll
  for (int c = 0; c < D; c++) {
    long long i = (long long)c * d.np + n;
    double a = d.a[i];
    double udt = d.u_dt[i] + P.beta * P.dt * P.dt * a;
    d.u_dt[i] = udt;
    d.x[i] = d.x[i] + udt;
    d.prev_a[i] = a;
    d.u[i] = d.u[i] + udt;
  }
}

// ---------------------------------------------------------------------------------------------
// multi-GPU halo exchange of partial nodal sums over peer memory (no reference exists)
//   k_halo_send<MODE>   : recompute this rank's partial sums for its shared nodes (same list order as the
//                         node kernels, so the values are the ones those kernels form) and store them
//                         straight into the neighbour's receive region over NVLink; the last block to
//                         finish publishes the exchange's sequence number in the neighbour's flag slot
//                         (system-scope release).  MODE 0 = init triple (sum vol_0, sum rho, count),
//                         1 = sum vol (+ sum vol/4 for ANP_Nodal), 2 = internal-force partial.
//   k_halo_wait         : one block; spins (acquire, system scope) until every neighbour has published
//                         the sequence number.  Kept separate from the consumers so that a waiting rank
//                         never occupies more than one CTA.
//   k_halo_finish<MODE> : per unique shared node, total = sum of the sharers' partials in ascending rank
//                         order (identical on every sharer, so all copies stay bit-identical), then the
//                         same epilogue as k_node_vol.  The force totals are formed inside k_node_update.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) k_halo_send(WfDev d, WfPar P, int sep, unsigned long long seq) {
  halo_send_cta<MODE>(d, P, sep, seq, blockIdx.y, blockIdx.x, gridDim.x);
}

__global__ void k_halo_wait(WfDev d, unsigned long long seq, unsigned long long timeout_ns) {
  halo_wait_cta(d, seq, timeout_ns);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_halo_finish(WfDev d, WfPar P, int parity) {
  if (P.wait_seq) halo_wait_cta(d, P.wait_seq, P.wait_timeout_ns); // folded k_halo_wait
  pdl_trigger(); // after the wait: see node_update_body
  // as a programmatic dependent launch of N1 the CTAs are resident (and have seen the neighbours' flags) while N1's
  // last wave drains; everything below reads what N1 / E1 wrote and overwrites N1's local-only sums of the shared nodes
  pdl_wait();
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= d.n_uniq) return;
  const int n = d.hu_node[u];
  double s, sq, rs;
  int cnt;
  halo_node_sums<MODE == 0>(d, n, MODE == 0 ? d.vol_0 : d.vol, s, sq, rs, cnt);
  if (MODE == 0) {
    const double t0 = halo_total(d, u, 0, parity, (P.press == 3) ? sq : s);
    const double t1 = halo_total(d, u, 1, parity, rs);
    const double t2 = halo_total(d, u, 2, parity, (double)cnt);
    d.voln0_sum[n] = t0;
    d.rhobar[n] = t1 / t2;
    d.nodel_count[n] = (int)t2;
    return;
  }
  s = halo_total(d, u, 0, parity, s);
  d.voln_sum[n] = s;
  if (P.press == 0) d.nodal_p[n] = s / d.voln0_sum[n];
  else if (P.press == 1) d.nodal_p[n] = P.Kbulk * (1.0 - s / d.voln0_sum[n]);
  else {
    sq = halo_total(d, u, 1, parity, sq);
    double v0 = d.voln0_sum[n], pn = 0.0;
    if (v0 > 1e-12) { double Jn = sq / v0; pn = P.Kbulk * (1.0 - Jn); }
    d.nodal_p[n] = pn;
  }
  // nodal mass of a shared node: sum_e rho_e voln / count regrouped as voln * mean(rho_e) over ALL sharers
  d.mdiag[n] = d.rhobar[n] * (s / (double)d.k);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// launch with programmatic stream serialisation (see pdl_trigger / pdl_wait); WF_PDL=0 in the environment falls back
// to ordinary launches (A/B measurements)
static bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("WF_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
template <class... KArgs, class... Args>
static void launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#define ELEM_DISPATCH(et, ...)                                          \
  switch (et) {                                                         \
    case ET_HEX8: { constexpr int ET = ET_HEX8; __VA_ARGS__; } break;   \
    case ET_TET4: { constexpr int ET = ET_TET4; __VA_ARGS__; } break;   \
    case ET_QUAD4: { constexpr int ET = ET_QUAD4; __VA_ARGS__; } break; \
    default: { constexpr int ET = ET_TRI3; __VA_ARGS__; } break;        \
  }

static void l_predict(const WfDev &d, const WfPar &P, int with_bc, cudaStream_t s) {
  if (d.dim == 3) k_predict<3><<<cdiv(d.nn, TPB_N), TPB_N, 0, s>>>(d, P, with_bc);
  else k_predict<2><<<cdiv(d.nn, TPB_N), TPB_N, 0, s>>>(d, P, with_bc);
}
static void l_bc_patch_v(const WfDev &d, const int *row_node, int nrows, cudaStream_t s) {
  if (nrows > 0) k_bc_patch_v<<<cdiv(nrows, 256), 256, 0, s>>>(d, row_node, nrows);
}
static void l_unpredict(const WfDev &d, const WfPar &P, cudaStream_t s) {
  if (d.dim == 3) k_unpredict<3><<<cdiv(d.nn, 256), 256, 0, s>>>(d, P);
  else k_unpredict<2><<<cdiv(d.nn, 256), 256, 0, s>>>(d, P);
}
static void l_impose_bc(const WfDev &d, int dim, int is_acc, double *arr, cudaStream_t s) {
  k_impose_bc<<<cdiv(d.nn, 256), 256, 0, s>>>(d, dim, is_acc, arr);
}
static void l_elem_vol(const WfDev &d, const WfPar &P, int et, int store_jac, cudaStream_t s) {
  if (!store_jac && et == ET_HEX8 && d.blk_pad_b && !P.strict && P.variant[0] == 0) {
    launch_pdl(k_elem_vol_brick<WF_BRICK_STRIDE>, d.n_bcta, WF_EBLK, (size_t)3 * WF_BRICK_STRIDE * 8, s, d, P);
    return;
  }
  if (!store_jac && P.variant[0] == 1) {
    const int stride = d.blk_pitch;
    ELEM_DISPATCH(et, k_elem_vol_staged<ET><<<cdiv(d.ne, WF_EBLK), WF_EBLK, Elem<ET>::D * stride * 8, s>>>(d, P, stride));
    return;
  }
  ELEM_DISPATCH(et, launch_pdl(k_elem_vol<ET>, cdiv(d.ne, TPB_E), TPB_E, 0, s, d, P, store_jac));
}
static void l_vol_from_detj(const WfDev &d, int et, cudaStream_t s) {
  ELEM_DISPATCH(et, k_vol_from_detj<ET><<<cdiv(d.ne, 256), 256, 0, s>>>(d));
}
static void l_node_vol(const WfDev &d, const WfPar &P, int mode, cudaStream_t s) {
  int g = cdiv((long long)d.nslices * 32, TPB_N) + ((mode == 1 || mode == 3) ? P.send_ctas : 0);
  const bool halo = d.n_neigh > 0;
  switch (d.k) {
    // register cap for 5 resident CTAs (48 registers): 0.235 -> 0.192 ms on 10M hexes; 6 and 8 CTAs (spills) 0.214 ms
    case 8:
      if (halo) launch_pdl(k_node_vol<8, 5, true>, g, TPB_N, 0, s, d, P, mode);
      else if (P.variant[1] == 5) launch_pdl(k_node_vol<8>, g, TPB_N, 0, s, d, P, mode);
      else launch_pdl(k_node_vol<8, 5>, g, TPB_N, 0, s, d, P, mode);
      break;
    case 4:
      if (halo) launch_pdl(k_node_vol<4, 5, true>, g, TPB_N, 0, s, d, P, mode);
      else launch_pdl(k_node_vol<4, 5>, g, TPB_N, 0, s, d, P, mode);
      break;
    default:
      if (halo) launch_pdl(k_node_vol<3, 5, true>, g, TPB_N, 0, s, d, P, mode);
      else launch_pdl(k_node_vol<3, 5>, g, TPB_N, 0, s, d, P, mode);
      break;
  }
}
// the tile-reduced force path (WfDev::ftile): same eligibility as the regrouped hexa kernel, default variant only
static int l_tile_forces(const WfDev &d, const WfPar &P, int separate_hg) {
  // 3D only: in 2D (1M quads) neither form of the tile reduction pays for itself (wf_engine.cu does not upload the tables)
  return d.ftile && !separate_hg && d.dim == 3 && !P.strict && !P.thermal && (P.variant[2] == 0 || P.variant[2] == 7) &&
         ((d.k == 8 && P.model < 2) || (d.k == 4 && d.tf_tab));
}
constexpr int BRICK_STRIDE = WF_BRICK_STRIDE, BRICK_WS = WF_BRICK_WS;
static void l_elem_main(const WfDev &d, const WfPar &P, int et, int separate_hg, cudaStream_t s) {
  if (et == ET_HEX8 && l_tile_forces(d, P, separate_hg)) {
    const int stride = d.blk_pitch, g = cdiv(d.ne, hexfast::TPB);
    // variant 7 = the generic tile kernel; it addresses tiles by the compact numbering, so not with a brick plan
    if (d.blk_pad_b && (P.variant[2] != 7 || d.brick_plan)) {
      constexpr size_t smem = ((size_t)7 * BRICK_STRIDE + (size_t)(hexfast::TPB / 32) * 3 * BRICK_WS) * 8;
      // 4 resident CTAs at 128 registers; measured alternatives: 5 CTAs at 96 registers (80 B of spills)
      // 0.910 vs 0.847 ms, L2 look-ahead of the next CTA's node data 0.787 vs 0.791 ms, a persistent double-buffered
      // form 1.07 vs 0.79 ms
      launch_pdl(hexfast::k_elem_main_hex_brick<BRICK_STRIDE, BRICK_WS, 4>, d.n_bcta, hexfast::TPB, smem, s, d, P);
      return;
    }
    const size_t smem = ((size_t)7 * stride + (size_t)(hexfast::TPB / 32) * 3 * d.tf_stride) * 8;
    launch_pdl(hexfast::k_elem_main_hex_tile, g, hexfast::TPB, smem, s, d, P, stride);
    return;
  }
  // the regrouped hexa kernel inlines Bilinear / Hollomon; the rate-dependent laws (Johnson-Cook, GMT) take the generic kernel
  if (!separate_hg && et == ET_HEX8 && !P.strict && P.variant[2] != 1 && P.model < 2 && !P.thermal) {
    // unique nodes of the CTA staged once in shared memory, one force record per element node (variant 9, and the
    // fallback when the force tiles of the mesh are not conflict-free)
    const int stride = d.blk_pitch;
    hexfast::k_elem_main_hex_staged<<<cdiv(d.ne, hexfast::TPB), hexfast::TPB, 7 * stride * 8, s>>>(d, P, stride);
    return;
  }
  if (et == ET_TET4 && l_tile_forces(d, P, separate_hg)) {
    const size_t smem = (size_t)(TPB_E / 32) * (12 * 32 + (d.tf_tpitch + 7) / 8) * 8;
    if (P.variant[2] == 7) k_elem_main<ET_TET4, false, false, false, true><<<cdiv(d.ne, TPB_E), TPB_E, smem, s>>>(d, P, 0);
    // five resident CTAs at 96 registers (40 B of spills) beat four without spills (0.566 vs 0.577 ms) and six (0.676 ms)
    else launch_pdl(k_elem_main<ET_TET4, false, false, false, true, 5>, cdiv(d.ne, TPB_E), TPB_E, smem, s, d, P, 0);
    return;
  }
  const int stride = d.blk_pitch;
  // measured (tools/kbench.py): per-element gathers beat block staging for tets, quads and the strict hexa kernel
  // (0.71 vs 0.78 ms at 10M tets); the staged form is kept as variant 8
  const bool staged = P.variant[2] == 8 && !P.thermal;
  if (staged) {
    if (separate_hg) { ELEM_DISPATCH(et, k_elem_main<ET, true, true><<<cdiv(d.ne, TPB_E), TPB_E, (2 * Elem<ET>::D + 1) * stride * 8, s>>>(d, P, stride)); }
    else { ELEM_DISPATCH(et, k_elem_main<ET, false, true><<<cdiv(d.ne, TPB_E), TPB_E, (2 * Elem<ET>::D + 1) * stride * 8, s>>>(d, P, stride)); }
  } else if (P.thermal) {
    if (separate_hg) { ELEM_DISPATCH(et, k_elem_main<ET, true, false, true><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P, 0)); }
    else { ELEM_DISPATCH(et, k_elem_main<ET, false, false, true><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P, 0)); }
  } else {
    if (separate_hg) { ELEM_DISPATCH(et, k_elem_main<ET, true, false><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P, 0)); }
    else { ELEM_DISPATCH(et, launch_pdl(k_elem_main<ET, false, false>, cdiv(d.ne, TPB_E), TPB_E, 0, s, d, P, 0)); }
  }
}
template <bool SEP, int U, int MINB = 1>
static void node_update_t(const WfDev &d, const WfPar &P, int fuse, int phase, cudaStream_t s) {
  int g = phase == 4 ? cdiv(std::max(d.n_uniq, 1), TPB_N) : cdiv((long long)d.nslices * 32, TPB_N) + (phase == 3 ? P.send_ctas : 0);
  if (d.n_neigh > 0) { // partitioned mesh: the instantiation that carries the folded halo send / wait
    if (d.dim == 3) launch_pdl(k_node_update<3, SEP, U, false, false, MINB, true>, g, TPB_N, 0, s, d, P, fuse, phase);
    else launch_pdl(k_node_update<2, SEP, U, false, false, MINB, true>, g, TPB_N, 0, s, d, P, fuse, phase);
    return;
  }
  if (d.dim == 3) launch_pdl(k_node_update<3, SEP, U, false, false, MINB>, g, TPB_N, 0, s, d, P, fuse, phase);
  else launch_pdl(k_node_update<2, SEP, U, false, false, MINB>, g, TPB_N, 0, s, d, P, fuse, phase);
}
static void l_node_update(const WfDev &d, const WfPar &P, int separate_hg, int fuse, int phase, cudaStream_t s) {
  if (l_tile_forces(d, P, separate_hg)) {
    const int g = phase == 4 ? cdiv(std::max(d.n_uniq, 1), TPB_N) : cdiv((long long)d.nslices * 32, TPB_N) + (phase == 3 ? P.send_ctas : 0);
    // measured on 10M hexes (tools/kbench.py): 5 resident CTAs (48 registers) + L2 prefetch of the state rows 0.60 ms;
    // 3 CTAs (67 registers) 0.66-0.75 ms; 6 CTAs (40 registers, spills) 0.61 ms; no prefetch 0.71 ms
    // (launching the pass as a programmatic dependent launch of E2 was measured: no gain on one GPU, 0.8 % slower on two)
    // phase 4 (shared nodes, begins with the flag wait) as a programmatic dependent launch of phase 3
    if (d.n_neigh > 0 && phase == 4 && P.variant[3] != 8) { launch_pdl(k_node_update<3, false, 4, true, true, 5, true, 4>, g, TPB_N, 0, s, d, P, fuse, phase); return; }
    if (d.n_neigh > 0 && phase == 3) k_node_update<3, false, 4, true, true, 5, true, 3><<<g, TPB_N, 0, s>>>(d, P, fuse, phase);
    else if (d.n_neigh > 0 && phase == 4) k_node_update<3, false, 4, true, true, 5, true, 4><<<g, TPB_N, 0, s>>>(d, P, fuse, phase);
    else if (d.n_neigh > 0) k_node_update<3, false, 4, true, true, 5, true><<<g, TPB_N, 0, s>>>(d, P, fuse, phase);
    else if (P.variant[3] == 5) k_node_update<3, false, 4, true, false, 5><<<g, TPB_N, 0, s>>>(d, P, fuse, phase);
    else if (P.variant[3] == 7) k_node_update<3, false, 4, true, true, 6><<<g, TPB_N, 0, s>>>(d, P, fuse, phase);
    else k_node_update<3, false, 4, true, true, 5><<<g, TPB_N, 0, s>>>(d, P, fuse, phase);
    return;
  }
  if (separate_hg) node_update_t<true, 4>(d, P, fuse, phase, s);
  else if (P.variant[3] == 1) node_update_t<false, 2>(d, P, fuse, phase, s);
  else if (P.variant[3] == 2) node_update_t<false, 8>(d, P, fuse, phase, s);
  else if (P.variant[3] == 5) node_update_t<false, 4>(d, P, fuse, phase, s);
  // (L2 prefetch of the state rows as in the tile path: no difference on 1 M quads, 0.0537 vs 0.0540 ms)
  else node_update_t<false, 4, 5>(d, P, fuse, phase, s); // 48 registers: 5 resident CTAs
}
static void l_node_thermal(const WfDev &d, const WfPar &P, cudaStream_t s) {
  k_node_thermal<<<cdiv((long long)d.nslices * 32, TPB_N), TPB_N, 0, s>>>(d, P);
}
static void l_node_mass(const WfDev &d, const WfPar &P, int use_stored_voln, cudaStream_t s) {
  int g = cdiv((long long)d.nslices * 32, TPB_N);
  switch (d.k) {
    case 8: k_node_mass<8><<<g, TPB_N, 0, s>>>(d, P, use_stored_voln); break;
    case 4: k_node_mass<4><<<g, TPB_N, 0, s>>>(d, P, use_stored_voln); break;
    default: k_node_mass<3><<<g, TPB_N, 0, s>>>(d, P, use_stored_voln); break;
  }
}
static void l_init_elem(const WfDev &d, const WfPar &P, cudaStream_t s) { k_init_elem<<<cdiv(d.ne, 256), 256, 0, s>>>(d, P); }
static void l_vol0_density(const WfDev &d, cudaStream_t s) { k_vol0_density<<<cdiv(d.ne, 256), 256, 0, s>>>(d); }
static void l_density(const WfDev &d, cudaStream_t s) { k_density<<<cdiv(d.ne, 256), 256, 0, s>>>(d); }
static void l_xmin(const WfDev &d, int slot, cudaStream_t s) { k_xmin<<<cdiv(d.nn, 256), 256, 0, s>>>(d, slot); }
static void l_rebuild_sigma(const WfDev &d, double *out, cudaStream_t s) { k_rebuild_sigma<<<cdiv(d.ne, 256), 256, 0, s>>>(d, out); }
static void l_energy(const WfDev &d, const double *sig, cudaStream_t s) {
  if (d.dim == 3) k_energy_kin<3><<<cdiv(d.nn, 256), 256, 0, s>>>(d);
  else k_energy_kin<2><<<cdiv(d.nn, 256), 256, 0, s>>>(d);
  if (d.ne > 0) k_energy_int<<<cdiv(d.ne, 256), 256, 0, s>>>(d, sig);
}
static void l_u_strain_rates(const WfDev &d, const WfPar &P, int et, cudaStream_t s) {
  ELEM_DISPATCH(et, k_u_strain_rates<ET><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P));
}
static void l_u_pressure(const WfDev &d, const WfPar &P, int et, cudaStream_t s) {
  ELEM_DISPATCH(et, k_u_pressure<ET><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P));
}
static void l_u_stress(const WfDev &d, const WfPar &P, double dt, cudaStream_t s) { k_u_stress<<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P, dt); }
static void l_u_artvisc(const WfDev &d, const WfPar &P, cudaStream_t s) { k_u_artvisc<<<cdiv(d.ne, 256), 256, 0, s>>>(d, P); }
static void l_u_forces(const WfDev &d, const WfPar &P, int et, cudaStream_t s) {
  ELEM_DISPATCH(et, k_u_forces<ET><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P));
}
static void l_u_hourglass(const WfDev &d, const WfPar &P, int et, cudaStream_t s) {
  ELEM_DISPATCH(et, k_u_hourglass<ET><<<cdiv(d.ne, TPB_E), TPB_E, 0, s>>>(d, P));
}
static void l_u_nodal_vol(const WfDev &d, cudaStream_t s) {
  int g = cdiv((long long)d.nslices * 32, TPB_N);
  switch (d.k) {
    case 8: k_u_nodal_vol<8><<<g, TPB_N, 0, s>>>(d); break;
    case 4: k_u_nodal_vol<4><<<g, TPB_N, 0, s>>>(d); break;
    default: k_u_nodal_vol<3><<<g, TPB_N, 0, s>>>(d); break;
  }
}
static void l_u_assembly(const WfDev &d, cudaStream_t s) {
  int g = cdiv((long long)d.nslices * 32, TPB_N);
  if (d.k == 8) k_u_assembly<8, 3><<<g, TPB_N, 0, s>>>(d);
  else if (d.k == 4 && d.dim == 3) k_u_assembly<4, 3><<<g, TPB_N, 0, s>>>(d);
  else if (d.k == 4) k_u_assembly<4, 2><<<g, TPB_N, 0, s>>>(d);
  else k_u_assembly<3, 2><<<g, TPB_N, 0, s>>>(d);
}
static void l_u_accel(const WfDev &d, cudaStream_t s) {
  if (d.dim == 3) k_u_accel<3><<<cdiv(d.nn, 256), 256, 0, s>>>(d);
  else k_u_accel<2><<<cdiv(d.nn, 256), 256, 0, s>>>(d);
}
static void l_u_corr_accvel(const WfDev &d, const WfPar &P, cudaStream_t s) {
  if (d.dim == 3) k_u_corr_accvel<3><<<cdiv(d.nn, 256), 256, 0, s>>>(d, P);
  else k_u_corr_accvel<2><<<cdiv(d.nn, 256), 256, 0, s>>>(d, P);
}
static void l_u_axis(const WfDev &d, const WfPar &P, cudaStream_t s) { k_u_axis<<<cdiv(d.nn, 256), 256, 0, s>>>(d, P); }
static void l_u_corr_pos(const WfDev &d, const WfPar &P, cudaStream_t s) {
  if (d.dim == 3) k_u_corr_pos<3><<<cdiv(d.nn, 256), 256, 0, s>>>(d, P);
  else k_u_corr_pos<2><<<cdiv(d.nn, 256), 256, 0, s>>>(d, P);
}
static void l_halo_send(const WfDev &d, const WfPar &P, int mode, int sep, unsigned long long seq, int max_count, cudaStream_t s) {
  if (d.n_neigh <= 0) return;
  dim3 g(cdiv(max_count > 0 ? max_count : 1, 128), d.n_neigh);
  if (mode == 2 && l_tile_forces(d, P, sep)) sep = 2; // partial force sums come from the tile partials
  if (mode == 0) k_halo_send<0><<<g, 128, 0, s>>>(d, P, sep, seq);
  else if (mode == 1) k_halo_send<1><<<g, 128, 0, s>>>(d, P, sep, seq);
  else k_halo_send<2><<<g, 128, 0, s>>>(d, P, sep, seq);
}
static void l_halo_wait(const WfDev &d, unsigned long long seq, unsigned long long timeout_ns, cudaStream_t s) {
  if (d.n_neigh > 0) k_halo_wait<<<1, 32 * ((d.n_neigh + 31) / 32), 0, s>>>(d, seq, timeout_ns);
}
static void l_halo_finish(const WfDev &d, const WfPar &P, int mode, int parity, cudaStream_t s) {
  if (d.n_uniq <= 0) return;
  if (mode == 0) k_halo_finish<0><<<cdiv(d.n_uniq, 128), 128, 0, s>>>(d, P, parity);
  else if (P.variant[1] == 7) k_halo_finish<1><<<cdiv(d.n_uniq, 128), 128, 0, s>>>(d, P, parity);
  // programmatic dependent launch of N1: launch latency and the flag wait overlap N1's last wave (2 GPUs, with the
  // same for phase 4 of N2: 1.553 vs 1.562 ms per step)
  else launch_pdl(k_halo_finish<1>, cdiv(d.n_uniq, 128), 128, 0, s, d, P, parity);
}

static void l_p_node(const WfDev &d, double *out, cudaStream_t s) {
  int g = cdiv((long long)d.nslices * 32, TPB_N);
  switch (d.k) {
    case 8: k_p_node<8><<<g, TPB_N, 0, s>>>(d, out); break;
    case 4: k_p_node<4><<<g, TPB_N, 0, s>>>(d, out); break;
    default: k_p_node<3><<<g, TPB_N, 0, s>>>(d, out); break;
  }
}
static void l_min_edge(const WfDev &d, double *elem_length, unsigned long long *keys, cudaStream_t s) {
  if (d.dim == 3) k_min_edge<3><<<cdiv(d.ne, 128), 128, 0, s>>>(d, elem_length, keys);
  else k_min_edge<2><<<cdiv(d.ne, 128), 128, 0, s>>>(d, elem_length, keys);
}
static void l_max_vel(const WfDev &d, unsigned long long *keys, cudaStream_t s) {
  if (d.dim == 3) k_max_vel<3><<<cdiv(d.nn, 256), 256, 0, s>>>(d, keys);
  else k_max_vel<2><<<cdiv(d.nn, 256), 256, 0, s>>>(d, keys);
}
static void l_soa_to_aos(const double *soa, long long pitch, int nc, long long n, double scale, double *aos, const int *map, cudaStream_t s) {
  if (n > 0) k_soa_to_aos<<<(unsigned)((n * nc + 255) / 256), 256, 0, s>>>(soa, pitch, nc, n, scale, aos, map);
}
static void l_aos_to_soa(const double *aos, long long pitch, int nc, long long n, double *soa, const int *map, cudaStream_t s) {
  if (n > 0) k_aos_to_soa<<<(unsigned)((n * nc + 255) / 256), 256, 0, s>>>(aos, pitch, nc, n, soa, map);
}

// Force-load every kernel of the step (CUDA loads kernels lazily, and loading one synchronises the context:
// a first launch issued while a halo wait kernel is spinning for work that the same host thread has not
// enqueued yet would deadlock).
template <class F>
static void touch(F *f) {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, f);
}
static void l_preload(int et, int dim, int k) {
  (void)k;
  // staged hexa kernel: up to 7 arrays x (128 elements x 8 nodes) doubles of shared memory
  cudaFuncSetAttribute(hexfast::k_elem_main_hex_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 1024 * 8);
  touch(hexfast::k_elem_main_hex_staged);
  cudaFuncSetAttribute(hexfast::k_elem_main_hex_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (7 * 1024 + 12 * 256) * 8);
  touch(hexfast::k_elem_main_hex_tile); touch(k_elem_main<ET_TET4, false, false, false, true>); touch(k_elem_main<ET_TET4, false, false, false, true, 5>);
  touch(k_node_update<3, false, 4, true, false, 5>); touch(k_node_update<3, false, 4, true, true, 6>); touch(k_node_update<3, false, 4, true, true, 5>);
  touch(k_predict<2>); touch(k_predict<3>); touch(k_impose_bc);
  ELEM_DISPATCH(et, touch(k_elem_vol<ET>); touch(k_elem_main<ET, true, false>); touch(k_elem_main<ET, false, false>);
                cudaFuncSetAttribute(k_elem_main<ET, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 1024 * 8);
                cudaFuncSetAttribute(k_elem_main<ET, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 1024 * 8);
                touch(k_elem_main<ET, true, true>); touch(k_elem_main<ET, false, true>));
  touch(k_node_vol<8>); touch(k_node_vol<8, 5>); touch(k_node_vol<4, 5>); touch(k_node_vol<3, 5>);
  touch(k_node_vol<8, 5, true>); touch(k_node_vol<4, 5, true>); touch(k_node_vol<3, 5, true>);
  touch(k_node_update<3, false, 4, true, true, 5, true>);
  touch(k_node_update<3, false, 4, true, true, 5, true, 3>); touch(k_node_update<3, false, 4, true, true, 5, true, 4>);
  touch(k_node_update<3, true, 4, false, false, 1, true>); touch(k_node_update<3, false, 4, false, false, 5, true>); touch(k_node_update<3, false, 4, false, false, 1, true>);
  touch(k_node_update<3, false, 2, false, false, 1, true>); touch(k_node_update<3, false, 8, false, false, 1, true>);
  touch(k_node_update<2, true, 4, false, false, 1, true>); touch(k_node_update<2, false, 4, false, false, 5, true>); touch(k_node_update<2, false, 4, false, false, 1, true>);
  touch(k_node_update<2, false, 2, false, false, 1, true>); touch(k_node_update<2, false, 8, false, false, 1, true>);
  touch(k_elem_vol_brick<WF_BRICK_STRIDE>);
  touch(hexfast::k_elem_main_hex_brick<BRICK_STRIDE, BRICK_WS, 4>);
  touch(k_node_update<3, true, 4>); touch(k_node_update<3, false, 4>); touch(k_node_update<3, false, 4, false, false, 5>); touch(k_node_update<3, false, 2>); touch(k_node_update<3, false, 8>);
  touch(k_node_update<2, true, 4>); touch(k_node_update<2, false, 4>); touch(k_node_update<2, false, 4, false, false, 5>); touch(k_node_update<2, false, 2>); touch(k_node_update<2, false, 8>);
  touch(k_halo_send<0>); touch(k_halo_send<1>); touch(k_halo_send<2>); touch(k_halo_wait);
  touch(k_halo_finish<0>); touch(k_halo_finish<1>);
  touch(k_init_elem); touch(k_vol0_density); touch(k_xmin); touch(k_energy_kin<2>); touch(k_energy_kin<3>);
  (void)dim;
}

} // namespace WF_NS

#define WF_CAT2(a, b) a##b
#define WF_CAT(a, b) WF_CAT2(a, b)
extern "C" const WfLaunch *WF_CAT(WF_NS, _table)() {
  using namespace WF_NS;
  static const WfLaunch t = {l_predict, l_impose_bc, l_elem_vol, l_vol_from_detj, l_node_vol,
                             l_elem_main, l_node_update, l_node_mass, l_init_elem, l_vol0_density, l_density, l_xmin,
                             l_rebuild_sigma, l_energy, l_u_strain_rates, l_u_pressure, l_u_stress, l_u_artvisc,
                             l_u_forces, l_u_hourglass, l_u_nodal_vol, l_u_assembly, l_u_accel, l_u_corr_accvel,
                             l_u_axis, l_u_corr_pos, l_halo_send, l_halo_wait, l_halo_finish, l_preload, l_p_node, l_min_edge, l_max_vel, l_soa_to_aos,
                             l_aos_to_soa, l_node_thermal, l_tile_forces, l_bc_patch_v, l_unpredict};
  return &t;
}
