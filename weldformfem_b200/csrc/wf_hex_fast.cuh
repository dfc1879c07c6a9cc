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
// Element-local node data (x, v of the 8 nodes) is staged in shared memory with cp.async, one private
// column per thread ([item][thread], conflict-free), so the 48 gathers are in flight together without
// holding 96 registers; no block-level synchronisation is needed because a thread only reads back
// what it requested itself.
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
WF_DI void wht_fwd(const double *col /*stride TPB*/, double (&G)[3], double (&h)[4]) {
  const double q0 = col[0 * TPB], q1 = col[1 * TPB], q2 = col[2 * TPB], q3 = col[3 * TPB];
  const double q4 = col[4 * TPB], q5 = col[5 * TPB], q6 = col[6 * TPB], q7 = col[7 * TPB];
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

// inverse butterfly: q_n = sum_r B[r] s_r(n) + sum_j c[j] Sig_j(n), stored to the node-ordered
// force buffer at f[off[n] + 32*i] (component i of the entry of this element in node n's list)
WF_DI void wht_inv_store(const double (&B)[3], const double (&c)[4], double *__restrict__ f, const unsigned (&off)[8], int i) {
  const double SS1 = -B[2], SS2 = B[2];
  const double SD1 = B[1] - c[0], SD2 = B[1] + c[0];
  const double DS1 = B[0] - c[1], DS2 = B[0] + c[1];
  const double DD1 = c[2] - c[3], DD2 = c[2] + c[3];
  const double Sa = SS1 - SD1, Sb = SS1 + SD1, Da = DS1 - DD1, Db = DS1 + DD1;
  const double Sc = SS2 - SD2, Sd = SS2 + SD2, Dc = DS2 - DD2, Dd = DS2 + DD2;
  f[(long long)off[0] + 32 * i] = Sa - Da;
  f[(long long)off[1] + 32 * i] = Sa + Da;
  f[(long long)off[2] + 32 * i] = Sb + Db;
  f[(long long)off[3] + 32 * i] = Sb - Db;
  f[(long long)off[4] + 32 * i] = Sc - Dc;
  f[(long long)off[5] + 32 * i] = Sc + Dc;
  f[(long long)off[6] + 32 * i] = Sd + Dd;
  f[(long long)off[7] + 32 * i] = Sd - Dd;
}

// x^y for x > 0 as exp(y log x) (relative error ~1e-15; WF_FAST only)
WF_DI double fast_pow(double x, double y) { return exp(y * log(x)); }

__global__ void __launch_bounds__(TPB, 4) k_elem_main_hex_fast(WfDev d, WfPar P) {
  extern __shared__ double sm[];
  const int t = threadIdx.x;
  const int e0 = blockIdx.x * TPB + t;
  const bool active = e0 < d.ne;
  const int e = active ? e0 : d.ne - 1; // tail threads shadow the last element and do not store
  double *col = sm + t;

  int nid[8];
#pragma unroll
  for (int n = 0; n < 8; n++) nid[n] = __ldg(d.elnod + (long long)n * d.ep + e);
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int n = 0; n < 8; n++) cp_async8(col + (c * 8 + n) * TPB, d.x + (long long)c * d.np + nid[n]);
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int n = 0; n < 8; n++) cp_async8(col + (24 + c * 8 + n) * TPB, d.v + (long long)c * d.np + nid[n]);
  cp_async_commit();

  // independent streaming loads
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; i++) tau[i] = d.tau[(long long)i * d.ep + e];
  double pl = d.pl_strain[e];
  const double rho_e = d.rho[e];
  double sy = d.sigma_y[e];
  double J_avg = 0.0, p;
  if (P.press == 1) p = d.p[e];
#pragma unroll
  for (int a = 0; a < 8; a++) J_avg += d.nodal_p[nid[a]];

  cp_async_wait_all();

  // ---- geometry ---------------------------------------------------------------------------------
  double J[3][3], dummy[4];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double G[3];
    wht_fwd<false>(col + (c * 8) * TPB, G, dummy);
    J[0][c] = 0.125 * G[0]; J[1][c] = 0.125 * G[1]; J[2][c] = 0.125 * G[2];
  }
  double A[3][3]; // A'(c,r) = 0.125 * adj(J)(c,r)
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
  const double detJ = 8.0 * (J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0]);
  const double vol = detJ * 8.0;

  // ---- velocity gradient + hourglass modes ----------------------------------------------------------
  double L[3][3], hm[3][4];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double G[3];
    wht_fwd<true>(col + (24 + i * 8) * TPB, G, hm[i]);
#pragma unroll
    for (int c = 0; c < 3; c++) L[i][c] = A[c][0] * G[0] + A[c][1] * G[1] + A[c][2] * G[2];
  }
  const double f = 1.0 / detJ;
  double Dr[6], Wr[3];
  Dr[0] = L[0][0] * f; Dr[1] = L[1][1] * f; Dr[2] = L[2][2] * f;
  const double hf = 0.5 * f;
  Dr[3] = hf * (L[0][1] + L[1][0]); Dr[4] = hf * (L[1][2] + L[2][1]); Dr[5] = hf * (L[0][2] + L[2][0]);
  Wr[0] = hf * (L[0][1] - L[1][0]); Wr[1] = hf * (L[1][2] - L[2][1]); Wr[2] = hf * (L[0][2] - L[2][0]);

  // ---- pressure ---------------------------------------------------------------------------------------
  if (P.press == 0) {
    J_avg *= 0.125;
    if (P.stab_simple) {
      double J_bar = J_avg;
      if (J_bar < P.J_min) J_bar = 0.2;
      p = -P.Kbulk * (J_bar - 1.0);
    } else {
      // div_v = sum_a gradN_a . v_a = trace of the un-normalised velocity gradient
      // J_local uses the volume stored by E1 so that vol/vol_0 is exactly 1 for an undeformed element
      p = pressure_default3d(P, J_avg, d.vol_0[e], d.vol[e], rho_e, L[0][0] + L[1][1] + L[2][2]);
    }
  } else if (P.press == 1) {
    p = (p + J_avg) * (0.25 * 8);
  } else {
    p = J_avg * 0.125;
  }

  // ---- Jaumann rate + J2 radial return (CalcStressStrain, Mechanical.C:1664-1839) ------------------
  double sig[6];
  {
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
    if (active) {
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
    }
  }

  // ---- element + hourglass nodal forces ----------------------------------------------------------------
  // symmetric sigma(c,i): (0,0)=0 (1,1)=1 (2,2)=2 (0,1)=3 (1,2)=4 (0,2)=5
  double ch = 0.0;
  if (P.hexa_hg != 0.0) ch = P.hexa_hg * fast_pow(vol, 0.6666666) * rho_e * 0.25 * P.cs0;
  if (!active) return;
  unsigned off[8];
#pragma unroll
  for (int n = 0; n < 8; n++) off[n] = (unsigned)__ldg(d.pos + (long long)n * d.ep + e);
  const double w = 8.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double sxi = (i == 0) ? sig[0] : (i == 1 ? sig[3] : sig[5]);
    const double syi = (i == 0) ? sig[3] : (i == 1 ? sig[1] : sig[4]);
    const double szi = (i == 0) ? sig[5] : (i == 1 ? sig[4] : sig[2]);
    double B[3], c[4];
#pragma unroll
    for (int r = 0; r < 3; r++) B[r] = w * (A[0][r] * sxi + A[1][r] * syi + A[2][r] * szi);
#pragma unroll
    for (int j = 0; j < 4; j++) c[j] = ch * hm[i][j];
    wht_inv_store(B, c, d.fsell, off, i);
  }
}

} // namespace hexfast
