// wf_hex_fast.cuh — regrouped main element pass for the reduced-integration hexa (WF_FAST flavour).
//
// Same algorithm as k_elem_main<ET_HEX8> (see wf_math.cuh for the reference expressions), evaluated
// through the Walsh-Hadamard structure of the trilinear hexa at its centroid:
//   * the eight nodal sign patterns of dN/dxi, dN/deta, dN/dzeta and of the four hourglass base vectors
//     (f90_ver/src/Mechanical.f90:259-263) are seven of the eight Walsh functions on the 2x2x2 cube, so
//     one 3-stage butterfly per nodal component yields the Jacobian columns / velocity-gradient modes
//     G_r and the hourglass modes h_j together (23 adds instead of 8*7 multiply-adds);
//   * dH(c,n) = sum_r A'(c,r) s_r(n) with A' = 0.125 adj(J), hence
//       L(i,c)  = sum_n v(n,i) dH(c,n)           = sum_r A'(c,r) G_r(v_i)
//       f(n,i)  = w sum_c dH(c,n) sigma(c,i)     = sum_r s_r(n) B(r,i),   B(r,i) = w sum_c A'(c,r) sigma(c,i)
//       f_hg(n,i) = -c_h sum_j h_j(v_i) Sig_j(n)
//     and the 24 nodal force components come out of one inverse butterfly per component.
// Kernels in this file, all one element per thread, 128 threads per CTA:
//   k_elem_main_hex_brick  DEFAULT on meshes whose CTA / tile node lists fit its compile-time pitches (every structured
//                          mesh in the engine's element order: 128 elements = 8x4x4 brick, 32 = 4x4x2).  Same data
//                          flow as k_elem_main_hex_tile with shared-memory pitches known at compile time (no address
//                          arithmetic per access), node copies and accumulators at bank-aware slots
//                          (wf_host_run_slots: every 64-bit access of a half-warp is conflict-free on a structured mesh),
//                          and the slots of an element's eight nodes fetched with one 16 B and one 8 B load.
//   k_elem_main_hex_tile   same with run-time pitches (any mesh whose force tiles are conflict-free).  The CTA stages x,
//                          v and the nodal ratio of its UNIQUE nodes once (cp.async through the fixed-pitch node list
//                          WfDev::blk_pad); nodal forces are summed per warp tile in shared memory and one partial per
//                          (tile, unique node) goes to HBM (WfDev::ftile).
//   k_elem_main_hex_staged same staging, one force record per element node into the node-ordered buffer (variant 9;
//                          also the fallback when the tile tables are unusable).
// (included inside the flavour namespace of wf_kernels.cu, after wf_math.cuh)

namespace hexfast {

constexpr int TPB = 128;
constexpr int ITEMS = 48;
constexpr int SMEM_BYTES = ITEMS * TPB * 8;

WF_DI void cp_async8(double *smem_dst, const double *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
WF_DI void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
WF_DI void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// forward butterfly: q[0..7] nodal values -> G[0..2] (xi, eta, zeta modes) and optionally h[0..3]
// (hourglass modes in the order of the reference table: eta*zeta, xi*zeta, xi*eta, xi*eta*zeta)
template <bool WITH_HG>
WF_DI void wht_fwd8(double q0, double q1, double q2, double q3, double q4, double q5, double q6, double q7, double (&G)[3],
                    double (&h)[4]) {
  const double sa = q1 + q0, da = q1 - q0, sb = q2 + q3, db = q2 - q3;
  const double sc = q5 + q4, dc = q5 - q4, sd = q6 + q7, dd = q6 - q7;
  const double ss1 = sb + sa, sd1 = sb - sa, ds1 = db + da;
  const double ss2 = sd + sc, sd2 = sd - sc, ds2 = dd + dc;
  G[0] = ds2 + ds1;
  G[1] = sd2 + sd1;
  G[2] = ss2 - ss1;
  if (WITH_HG) {
    const double dd1 = db - da, dd2 = dd - dc;
    h[0] = sd2 - sd1;
    h[1] = ds2 - ds1;
    h[2] = dd2 + dd1;
    h[3] = dd2 - dd1;
  }
}
// node data staged once per CTA: value of node n of this element at s[comp * stride + li[n]]
struct StagedSrc {
  const double *s;
  int stride;
  unsigned li[8];
  WF_DI double operator()(int comp, int n) const { return s[comp * stride + li[n]]; }
};
template <bool WITH_HG, class Src>
WF_DI void wht_fwd(const Src &src, int comp, double (&G)[3], double (&h)[4]) {
  wht_fwd8<WITH_HG>(src(comp, 0), src(comp, 1), src(comp, 2), src(comp, 3), src(comp, 4), src(comp, 5), src(comp, 6),
                    src(comp, 7), G, h);
}

// inverse butterfly: q_n = sum_r B[r] s_r(n) + sum_j c[j] Sig_j(n) for the eight corners
WF_DI void wht_inv8(const double (&B)[3], const double (&c)[4], double (&q)[8]) {
  const double SS1 = -B[2], SS2 = B[2];
  const double SD1 = B[1] - c[0], SD2 = B[1] + c[0];
  const double DS1 = B[0] - c[1], DS2 = B[0] + c[1];
  const double DD1 = c[2] - c[3], DD2 = c[2] + c[3];
  const double Sa = SS1 - SD1, Sb = SS1 + SD1, Da = DS1 - DD1, Db = DS1 + DD1;
  const double Sc = SS2 - SD2, Sd = SS2 + SD2, Dc = DS2 - DD2, Dd = DS2 + DD2;
  q[0] = Sa - Da; q[1] = Sa + Da; q[2] = Sb + Db; q[3] = Sb - Db;
  q[4] = Sc - Dc; q[5] = Sc + Dc; q[6] = Sd + Dd; q[7] = Sd - Dd;
}
// ... stored to the node-ordered force buffer at f[off[n] + 32*i] (component i of the entry of this element in
// node n's list)
WF_DI void wht_inv_store(const double (&B)[3], const double (&c)[4], double *__restrict__ f, const unsigned (&off)[8], int i) {
  double q[8];
  wht_inv8(B, c, q);
#pragma unroll
  for (int n = 0; n < 8; n++) f[(long long)off[n] + 32 * i] = q[n];
}
// the two ways the main pass hands nodal forces on: scattered into the node-ordered buffer (one record per element
// node), or kept in registers for the tile reduction (k_elem_main_hex_tile)
template <class OffT>
struct EmitScatter {
  double *__restrict__ f;
  const OffT &off;
  unsigned o8[8];
  WF_DI void operator()(int i, const double (&B)[3], const double (&c)[4]) {
    if (i == 0) {
#pragma unroll
      for (int n = 0; n < 8; n++) o8[n] = off(n);
    }
    wht_inv_store(B, c, f, o8, i);
  }
};
struct EmitRegs {
  double (&q)[3][8];
  WF_DI void operator()(int i, const double (&B)[3], const double (&c)[4]) { wht_inv8(B, c, q[i]); }
};

// x^y for x > 0 as exp(y log x) (relative error ~1e-15; WF_FAST only)
WF_DI double fast_pow(double x, double y) { return exp(y * log(x)); }

// ---- front half: geometry, velocity gradient and hourglass modes from the staged node data ---------------
struct HexFront {
  double A[3][3]; // A'(c,r) = 0.125 * adj(J)(c,r)
  double detJ;
  double Dr[6], Wr[3];
  double trL;     // un-normalised div v = sum_a gradN_a . v_a
  double hm[3][4];
};

template <class Src>
WF_DI void hex_front(const Src &src, HexFront &g) {
  double J[3][3], dummy[4];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double G[3];
    wht_fwd<false>(src, c, G, dummy);
    J[0][c] = 0.125 * G[0]; J[1][c] = 0.125 * G[1]; J[2][c] = 0.125 * G[2];
  }
  double (&A)[3][3] = g.A;
  A[0][0] = 0.125 * (J[1][1] * J[2][2] - J[1][2] * J[2][1]);
  A[1][0] = -0.125 * (J[1][0] * J[2][2] - J[1][2] * J[2][0]);
  A[2][0] = 0.125 * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  A[0][1] = -0.125 * (J[0][1] * J[2][2] - J[0][2] * J[2][1]);
  A[1][1] = 0.125 * (J[0][0] * J[2][2] - J[0][2] * J[2][0]);
  A[2][1] = -0.125 * (J[0][0] * J[2][1] - J[0][1] * J[2][0]);
  A[0][2] = 0.125 * (J[0][1] * J[1][2] - J[0][2] * J[1][1]);
  A[1][2] = -0.125 * (J[0][0] * J[1][2] - J[0][2] * J[1][0]);
  A[2][2] = 0.125 * (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
  // det J = sum_c J(0,c) adj(c,0)
  g.detJ = 8.0 * (J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0]);
  double L[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double G[3];
    wht_fwd<true>(src, 3 + i, G, g.hm[i]);
#pragma unroll
    for (int c = 0; c < 3; c++) L[i][c] = A[c][0] * G[0] + A[c][1] * G[1] + A[c][2] * G[2];
  }
  const double f = 1.0 / g.detJ;
  g.Dr[0] = L[0][0] * f; g.Dr[1] = L[1][1] * f; g.Dr[2] = L[2][2] * f;
  const double hf = 0.5 * f;
  g.Dr[3] = hf * (L[0][1] + L[1][0]); g.Dr[4] = hf * (L[1][2] + L[2][1]); g.Dr[5] = hf * (L[0][2] + L[2][0]);
  g.Wr[0] = hf * (L[0][1] - L[1][0]); g.Wr[1] = hf * (L[1][2] - L[2][1]); g.Wr[2] = hf * (L[0][2] - L[2][0]);
  g.trL = L[0][0] + L[1][1] + L[2][2];
}

// ---- back half: pressure, Jaumann rate + J2 radial return, element + hourglass nodal forces --------------
// J_sum = sum of nodal_p over the 8 nodes; p_prev = stored pressure (press == 1 only); emit(i, B, c) receives the
// Walsh coefficients of force component i (EmitScatter / EmitRegs)
template <class Emit>
WF_DI void hex_back(const WfDev &d, const WfPar &P, int e, bool active, const HexFront &g, const double (&tau)[6],
                    double pl, double rho_e, double sy, double J_sum, double p_prev, Emit &&emit) {
  const double vol = g.detJ * 8.0;
  double p;
  if (P.press == 0) {
    const double J_avg = J_sum * 0.125;
    if (P.stab_simple) {
      double J_bar = J_avg;
      if (J_bar < P.J_min) J_bar = 0.2;
      p = -P.Kbulk * (J_bar - 1.0);
    } else {
      // J_local uses the volume stored by E1 so that vol/vol_0 is exactly 1 for an undeformed element
      p = pressure_default3d(P, J_avg, d.vol_0[e], d.vol[e], rho_e, g.trL);
    }
  } else if (P.press == 1) {
    p = (p_prev + J_sum) * (0.25 * 8);
  } else {
    p = J_sum * 0.125;
  }

  // CalcStressStrain, Mechanical.C:1664-1839
  double sig[6];
  const double (&Dr)[6] = g.Dr;
  const double (&Wr)[3] = g.Wr;
  const double txx = tau[0], tyy = tau[1], tzz = tau[2], txy = tau[3], tyz = tau[4], txz = tau[5];
  const double wxy = Wr[0], wyz = Wr[1], wxz = Wr[2];
  // SRT + RS with the reference's tensor3 operator* (Tensor3.C:290-304), zero products dropped
  const double r_xx = 2.0 * (txy * wxy + txz * wxz);
  // (SRT+RS)_yy and (SRT+RS)_zz cancel identically under that operator
  const double r_xy = (txx * wxy - txz * wyz) + (wxy * tyy + wxz * tyz);
  const double r_yz = (txy * wxz + tyy * wyz) + (wyz * tzz - wxy * txz);
  const double r_xz = (txx * wxz + txy * wyz) + (wxy * tyz + wxz * tzz);
  const double trD3 = (1.0 / 3.0) * (Dr[0] + Dr[1] + Dr[2]);
  const double g2 = 2.0 * P.G, dt = P.dt;
  double tt[6];
  tt[0] = txx + dt * ((Dr[0] - trD3) * g2 + r_xx);
  tt[1] = tyy + dt * ((Dr[1] - trD3) * g2);
  tt[2] = tzz + dt * ((Dr[2] - trD3) * g2);
  tt[3] = txy + dt * (Dr[3] * g2 + r_xy);
  tt[4] = tyz + dt * (Dr[4] * g2 + r_yz);
  tt[5] = txz + dt * (Dr[5] * g2 + r_xz);
  // s = dev(-p I + tau) = tau - tr(tau)/3 I   (the -p I part cancels in the deviator)
  const double tr3 = (1.0 / 3.0) * (tt[0] + tt[1] + tt[2]);
  const double s0 = tt[0] - tr3, s1 = tt[1] - tr3, s2 = tt[2] - tr3;
  const double J2 = 0.5 * (s0 * s0 + s1 * s1 + s2 * s2) + (tt[3] * tt[3] + tt[4] * tt[4] + tt[5] * tt[5]);
  const double sig_trial = sqrt(3.0 * J2);
  double b = 0.0;
  bool hard = false;
  if (P.model == 1) {
    b = pl + P.eps0;
    hard = b > P.eps1;
    sy = hard ? P.Kh * fast_pow(b, P.mh) : P.sy0;
  }
  if (sy < sig_trial) {
    const double H = hard ? P.mh * sy / b : 0.0; // K m b^(m-1) = m sy / b
    const double G3 = 3.0 * P.G;
    const double dgamma = (sig_trial - sy) / (G3 + H);
    const double factor = 1.0 - (G3 * dgamma) / sig_trial;
    tt[0] = s0 * factor; tt[1] = s1 * factor; tt[2] = s2 * factor;
    tt[3] *= factor; tt[4] *= factor; tt[5] *= factor;
    pl += dgamma;
  }
  sig[0] = tt[0] - p; sig[1] = tt[1] - p; sig[2] = tt[2] - p;
  sig[3] = tt[3]; sig[4] = tt[4]; sig[5] = tt[5];
  if (P.av_alpha != 0.0 || P.av_beta != 0.0) artificial_viscosity(P, Dr, rho_e, vol, sig);
  if (!active) return;
#pragma unroll
  for (int i = 0; i < 6; i++) d.tau[(long long)i * d.ep + e] = tt[i];
  if (P.store_sigma) {
#pragma unroll
    for (int i = 0; i < 6; i++) d.sigma[(long long)i * d.ep + e] = sig[i];
  }
  if (P.track_eps) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
      long long o = (long long)i * d.ep + e;
      d.eps[o] = d.eps[o] + dt * Dr[i];
    }
  }
  d.pl_strain[e] = pl;
  d.sigma_y[e] = sy;
  d.p[e] = p;

  // element + hourglass nodal forces; symmetric sigma(c,i): (0,0)=0 (1,1)=1 (2,2)=2 (0,1)=3 (1,2)=4 (0,2)=5
  double ch = 0.0;
  if (P.hexa_hg != 0.0) ch = P.hexa_hg * fast_pow(vol, 0.6666666) * rho_e * 0.25 * P.cs0;
  const double w = 8.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double sxi = (i == 0) ? sig[0] : (i == 1 ? sig[3] : sig[5]);
    const double syi = (i == 0) ? sig[3] : (i == 1 ? sig[1] : sig[4]);
    const double szi = (i == 0) ? sig[5] : (i == 1 ? sig[4] : sig[2]);
    double B[3], c[4];
#pragma unroll
    for (int r = 0; r < 3; r++) B[r] = w * (g.A[0][r] * sxi + g.A[1][r] * syi + g.A[2][r] * szi);
#pragma unroll
    for (int j = 0; j < 4; j++) c[j] = ch * g.hm[i][j];
    emit(i, B, c);
  }
}

// Block-staged variant: the CTA loads the data of its UNIQUE nodes (x, v, nodal ratio: 7 doubles per node) into
// shared memory once, coalesced along the ascending node list, instead of every element gathering its eight
// nodes separately (a structured 128-element tile has ~520 unique nodes for 1024 element-nodes).  Element-nodes
// are addressed by 16-bit block-local indices (WfDev::lidx).
__global__ void __launch_bounds__(TPB, 4) k_elem_main_hex_staged(WfDev d, WfPar P, int stride) {
  extern __shared__ double sm[];
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  const int u0 = __ldg(d.blk_off + b), U = __ldg(d.blk_off + b + 1) - u0;
  for (int i = t; i < U; i += TPB) {
    const int g = __ldg(d.blk_nodes + u0 + i);
#pragma unroll
    for (int c = 0; c < 3; c++) cp_async8(sm + c * stride + i, d.x + (long long)c * d.np + g);
#pragma unroll
    for (int c = 0; c < 3; c++) cp_async8(sm + (3 + c) * stride + i, d.v + (long long)c * d.np + g);
    cp_async8(sm + 6 * stride + i, d.nodal_p + g);
  }
  cp_async_commit();
  const int e0 = b * TPB + t;
  const bool active = e0 < d.ne;
  const int e = active ? e0 : d.ne - 1;
  StagedSrc src;
  src.s = sm; src.stride = stride;
#pragma unroll
  for (int n = 0; n < 8; n++) src.li[n] = d.lidx[(long long)n * d.ep + e];
  unsigned off[8];
#pragma unroll
  for (int n = 0; n < 8; n++) off[n] = (unsigned)__ldg(d.pos + (long long)n * d.ep + e);
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; i++) tau[i] = d.tau[(long long)i * d.ep + e];
  const double pl = d.pl_strain[e];
  const double rho_e = d.rho[e];
  const double sy = d.sigma_y[e];
  double p_prev = 0.0;
  if (P.press == 1) p_prev = d.p[e];
  cp_async_wait_all();
  __syncthreads();
  HexFront g;
  hex_front(src, g);
  double J_sum = 0.0;
#pragma unroll
  for (int a = 0; a < 8; a++) J_sum += sm[6 * stride + src.li[a]];
  auto offf = [&](int n) { return off[n]; };
  hex_back(d, P, e, active, g, tau, pl, rho_e, sy, J_sum, p_prev, EmitScatter<decltype(offf)>{d.fsell, offf});
}

// Tile-reduced variant of the block-staged kernel (WfDev::ftile): the nodal forces of the 32 elements of a warp (one
// "force tile") are summed per unique node in shared memory and ONE partial per (tile, unique node) goes to HBM,
// instead of one record per element node (a 32-element row segment of a structured mesh has 132 unique nodes for
// 256 element nodes, so the force traffic of E2 and N2 halves, and the scatter offsets `pos` are not read at all).
// The sum is formed in rounds, one per (component, local corner): the host verified that within a tile no two
// elements share a node at the same corner (always true on structured meshes), so every round is a conflict-free
// read-modify-write separated by __syncwarp and the order of additions per node is fixed (corner 0 first ... corner
// 7 last): deterministic, no atomics, no block-level barrier.
struct EmitTile {
  double *acc;       // [3][ws] accumulators of this warp
  int ws;
  const unsigned (&ri)[8];
  unsigned amask;    // lanes that own an element
  WF_DI void operator()(int i, const double (&B)[3], const double (&c)[4]) {
    double q[8];
    wht_inv8(B, c, q);
    double *a = acc + i * ws;
#pragma unroll
    for (int n = 0; n < 8; n++) {
      a[ri[n]] += q[n];
      __syncwarp(amask);
    }
  }
};

__global__ void __launch_bounds__(TPB, 4) k_elem_main_hex_tile(WfDev d, WfPar P, int stride) {
  extern __shared__ double sm[];
  pdl_trigger();
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  const int e0 = b * TPB + t;
  const bool active = e0 < d.ne;
  const int e = active ? e0 : d.ne - 1;
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  // node ids of the CTA's unique nodes (fixed-pitch list, -1 padded): the only loads the gathers depend on go first
  constexpr int NQ = 5;
  const int *__restrict__ ids = d.blk_pad + (long long)b * stride;
  int gid[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int i = q * TPB + t;
    gid[q] = (i < stride) ? __ldg(ids + i) : -1;
  }
  // streaming state and index loads, independent of the node tables
  StagedSrc src;
  src.s = sm; src.stride = stride;
#pragma unroll
  for (int n = 0; n < 8; n++) src.li[n] = d.lidx[(long long)n * d.ep + e];
  unsigned ri[8];
#pragma unroll
  for (int n = 0; n < 8; n++) ri[n] = d.tf_idx[(long long)n * d.ep + e];
  pdl_wait(); // everything above reads constant mesh tables only
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; i++) tau[i] = d.tau[(long long)i * d.ep + e];
  const double pl = d.pl_strain[e];
  const double rho_e = d.rho[e];
  const double sy = d.sigma_y[e];
  double p_prev = 0.0;
  if (P.press == 1) p_prev = d.p[e];
  // stage x, v and the nodal ratio of every unique node
  auto stage = [&](int i, int gq) {
    if (gq >= 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) cp_async8(sm + c * stride + i, d.x + (long long)c * d.np + gq);
#pragma unroll
      for (int c = 0; c < 3; c++) cp_async8(sm + (3 + c) * stride + i, d.v + (long long)c * d.np + gq);
      cp_async8(sm + 6 * stride + i, d.nodal_p + gq);
    }
  };
#pragma unroll
  for (int q = 0; q < NQ; q++) stage(q * TPB + t, gid[q]);
  for (int i = NQ * TPB + t; i < stride; i += TPB) stage(i, __ldg(ids + i)); // lists longer than NQ * TPB
  cp_async_commit();
  // this warp's accumulators
  const int ws = d.tf_stride, lane = t & 31, warp = t >> 5;
  double *acc = sm + 7 * stride + warp * 3 * ws;
  for (int i = lane; i < 3 * ws; i += 32) acc[i] = 0.0;
  cp_async_wait_all();
  __syncthreads();
  HexFront g;
  hex_front(src, g);
  double J_sum = 0.0;
#pragma unroll
  for (int a = 0; a < 8; a++) J_sum += sm[6 * stride + src.li[a]];
  hex_back(d, P, e, active, g, tau, pl, rho_e, sy, J_sum, p_prev, EmitTile{acc, ws, ri, amask});
  __syncwarp();
  const long long tile = (long long)b * (TPB / 32) + warp;
  // number of unique nodes of this tile = largest tile-local index + 1: only that many partials are written
  unsigned m = max(max(ri[0], ri[1]), max(ri[2], ri[3]));
  m = max(m, max(max(ri[4], ri[5]), max(ri[6], ri[7])));
  const int cnt = (int)__reduce_max_sync(0xffffffffu, m) + 1;
  if (tile * 32 < d.ne) {
    double *__restrict__ out = d.ftile + tile * 3 * ws;
#pragma unroll
    for (int c = 0; c < 3; c++)
      for (int i = lane; i < cnt; i += 32) out[c * ws + i] = acc[c * ws + i];
  }
}

// ---- brick form: compile-time pitches, packed local indices ------------------------------------------------------
// Per thread slot (WfDev::brick_elem): lidx_pk = the eight 16-bit CTA-local node slots of the element (one uint4),
// tile_pk = its eight 8-bit accumulator slots, the tile's rank -> slot word and the element id (one uint4).
template <int STRIDE>
struct StagedSrcC {
  const double *s;
  unsigned li[8];
  WF_DI double operator()(int comp, int n) const { return s[comp * STRIDE + li[n]]; }
};
template <int WS, bool BATCH3>
struct EmitTileC {
  double *acc;       // [3][WS] accumulators of this warp
  uint2 pk;          // eight 8-bit tile-local slots
  unsigned amask;
  double qb[2][8];   // BATCH3: corner values of components 0 and 1, kept until component 2 arrives
  WF_DI unsigned idx(int n) const { return ((n < 4 ? pk.x : pk.y) >> (8 * (n & 3))) & 0xffu; }
  WF_DI void operator()(int i, const double (&B)[3], const double (&c)[4]) {
    double q[8];
    wht_inv8(B, c, q);
    if (!BATCH3) {
      double *a = acc + i * WS;
#pragma unroll
      for (int n = 0; n < 8; n++) {
        a[idx(n)] += q[n];
        __syncwarp(amask);
      }
    } else if (i < 2) {
#pragma unroll
      for (int n = 0; n < 8; n++) qb[i][n] = q[n];
    } else {
      // one round per corner for all three components: three independent read-modify-writes in flight per round
      // (the rounds are a serial chain of shared-memory latencies: 8 links instead of 24)
#pragma unroll
      for (int n = 0; n < 8; n++) {
        double *a = acc + idx(n);
        const double a0 = a[0], a1 = a[WS], a2 = a[2 * WS];
        a[0] = a0 + qb[0][n]; a[WS] = a1 + qb[1][n]; a[2 * WS] = a2 + q[n];
        __syncwarp(amask);
      }
    }
  }
};

template <int STRIDE, int WS, int MINB, bool BATCH3 = true>
__global__ void __launch_bounds__(TPB, MINB) k_elem_main_hex_brick(WfDev d, WfPar P) {
  extern __shared__ double sm[];
  pdl_trigger();
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  constexpr int NQ = (STRIDE + TPB - 1) / TPB;
  const int *__restrict__ ids = d.blk_pad_b + (long long)b * STRIDE;
  int gid[NQ];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int i = q * TPB + t;
    gid[q] = (i < STRIDE) ? __ldg(ids + i) : -1;
  }
  // this thread slot: the element (or the element an idle thread shadows), its node slots in the CTA copy and in the
  // tile's accumulators, and (z) the rank -> slot word of this warp's tile, used at the very end
  const uint4 lpk = __ldg(d.lidx_pk + (long long)b * TPB + t);
  const uint4 tpk = __ldg(d.tile_pk + (long long)b * TPB + t);
  const int e0 = (int)tpk.w;
  const bool active = e0 >= 0;
  const int e = active ? e0 : ~e0;
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  const uint2 rpk = make_uint2(tpk.x, tpk.y);
  const unsigned r2s = tpk.z;
  // the node list of the CTA that will follow this one on the SM: ask L2 for it now (the first thing that CTA waits for)
  if (t < (STRIDE * 4 + 127) / 128) {
    const long long nb = (long long)b + d.cta_lookahead;
    if (nb < d.n_bcta) prefetch_l2(reinterpret_cast<const char *>(d.blk_pad_b + nb * STRIDE) + t * 128);
  }
  pdl_wait(); // everything above reads constant mesh tables only
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; i++) tau[i] = d.tau[(long long)i * d.ep + e];
  const double pl = d.pl_strain[e];
  const double rho_e = d.rho[e];
  const double sy = d.sigma_y[e];
  double p_prev = 0.0;
  if (P.press == 1) p_prev = d.p[e];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int i = q * TPB + t, gq = gid[q];
    if (gq >= 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) cp_async8(sm + c * STRIDE + i, d.x + (long long)c * d.np + gq);
#pragma unroll
      for (int c = 0; c < 3; c++) cp_async8(sm + (3 + c) * STRIDE + i, d.v + (long long)c * d.np + gq);
      cp_async8(sm + 6 * STRIDE + i, d.nodal_p + gq);
    }
  }
  cp_async_commit();
  const int lane = t & 31, warp = t >> 5;
  double *acc = sm + 7 * STRIDE + warp * 3 * WS;
#pragma unroll
  for (int i = lane; i < 3 * WS; i += 32) acc[i] = 0.0;
  StagedSrcC<STRIDE> src;
  src.s = sm;
  src.li[0] = lpk.x & 0xffffu; src.li[1] = lpk.x >> 16; src.li[2] = lpk.y & 0xffffu; src.li[3] = lpk.y >> 16;
  src.li[4] = lpk.z & 0xffffu; src.li[5] = lpk.z >> 16; src.li[6] = lpk.w & 0xffffu; src.li[7] = lpk.w >> 16;
  cp_async_wait_all();
  __syncthreads();
  HexFront g;
  hex_front(src, g);
  double J_sum = 0.0;
#pragma unroll
  for (int a = 0; a < 8; a++) J_sum += sm[6 * STRIDE + src.li[a]];
  EmitTileC<WS, BATCH3> emit;
  emit.acc = acc; emit.pk = rpk; emit.amask = amask;
  hex_back(d, P, e, active, g, tau, pl, rho_e, sy, J_sum, p_prev, emit);
  __syncwarp();
  // one partial per unique node of the tile, in rank order (the accumulators sit at bank-aware slots; r2s was loaded in
  // the prologue: byte j = slot of rank lane + 32 j, 0xff = none)
  const long long tile = (long long)b * (TPB / 32) + warp;
  {
    double *__restrict__ out = d.ftile + tile * 3 * d.tf_stride;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned sl = (r2s >> (8 * j)) & 0xffu;
      if (sl != 0xffu) {
#pragma unroll
        for (int c = 0; c < 3; c++) out[c * d.tf_stride + lane + 32 * j] = acc[c * WS + sl];
      }
    }
  }
}

} // namespace hexfast
