// wf_math.cuh — per-element arithmetic of the explicit step, register-resident.
//
// These are fresh device functions (no heap `Matrix`, no `tensor3` objects); each
// states the reference expression it evaluates and keeps the reference's evaluation
// order so that the WF_STRICT build (compiled with -fmad=false) reproduces the
// reference CPU path bit for bit apart from libm (`pow`, `log`) rounding.
#pragma once
#include "wf_dev.h"
#include "wf_math_ids.h"

template <int ET> struct Elem;
template <> struct Elem<ET_HEX8> { static constexpr int K = 8, D = 3; };
template <> struct Elem<ET_TET4> { static constexpr int K = 4, D = 3; };
template <> struct Elem<ET_QUAD4> { static constexpr int K = 4, D = 2; };
template <> struct Elem<ET_TRI3> { static constexpr int K = 3, D = 2; };

#define WF_DI __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// Jacobian J, A = adj(J) and det J at the single Gauss point.
// calcElemJAndDerivatives (Domain_d.C:1779-1888), AdjMat (Matrix.h:693-726), calcDet (Matrix.h:592-617)
// ------------------------------------------------------------------------------------------------
template <int ET>
WF_DI void jac_adj_det(const double (&xl)[Elem<ET>::K][Elem<ET>::D], double (&A)[Elem<ET>::D][Elem<ET>::D], double &detJ) {
  constexpr int D = Elem<ET>::D;
  double J[D][D];
#pragma unroll
  for (int c = 0; c < D; c++) {
    if constexpr (ET == ET_HEX8) {
      J[0][c] = 0.125 * (-xl[0][c] + xl[1][c] + xl[2][c] - xl[3][c] - xl[4][c] + xl[5][c] + xl[6][c] - xl[7][c]);
      J[1][c] = 0.125 * (-xl[0][c] - xl[1][c] + xl[2][c] + xl[3][c] - xl[4][c] - xl[5][c] + xl[6][c] + xl[7][c]);
      J[2][c] = 0.125 * (-xl[0][c] - xl[1][c] - xl[2][c] - xl[3][c] + xl[4][c] + xl[5][c] + xl[6][c] + xl[7][c]);
    } else if constexpr (ET == ET_TET4) {
      J[0][c] = xl[1][c] - xl[0][c];
      J[1][c] = xl[2][c] - xl[0][c];
      J[2][c] = xl[3][c] - xl[0][c];
    } else if constexpr (ET == ET_QUAD4) {
      J[0][c] = 0.25 * (-xl[0][c] + xl[1][c] + xl[2][c] - xl[3][c]);
      J[1][c] = 0.25 * (-xl[0][c] - xl[1][c] + xl[2][c] + xl[3][c]);
    } else {
      J[0][c] = (xl[0][c] - xl[2][c]);
      J[1][c] = (xl[1][c] - xl[2][c]);
    }
  }
  if constexpr (D == 2) {
    // 2x2 "adjugate" exactly as shipped: no minus signs, J11 used twice (Matrix.h:696-698)
    A[0][0] = J[1][1]; A[0][1] = J[1][0];
    A[1][0] = J[0][1]; A[1][1] = J[1][1];
    detJ = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  } else {
    constexpr int X = 2;
    // cofactor matrix, then transposed: A(i,j) = cof(j,i)
    A[0][0] = (J[1][1] * J[X][X] - J[1][X] * J[X][1]);
    A[1][0] = -(J[1][0] * J[X][X] - J[1][X] * J[X][0]);
    A[X][0] = (J[1][0] * J[X][1] - J[1][1] * J[X][0]);
    A[0][1] = -(J[0][1] * J[X][X] - J[0][X] * J[X][1]);
    A[1][1] = (J[0][0] * J[X][X] - J[0][X] * J[X][0]);
    A[X][1] = -(J[0][0] * J[X][1] - J[0][1] * J[X][0]);
    A[0][X] = (J[0][1] * J[1][X] - J[0][X] * J[1][1]);
    A[1][X] = -(J[0][0] * J[1][X] - J[0][X] * J[1][0]);
    A[X][X] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
    detJ = J[0][0] * J[1][1] * J[X][X] - J[0][0] * J[1][X] * J[X][1] - J[0][1] * J[1][0] * J[X][X] +
           J[0][1] * J[1][X] * J[X][0] + J[0][X] * J[1][0] * J[X][1] - J[0][X] * J[1][1] * J[X][0];
  }
}

// dH(c,n) = dN_n/dX_c * detJ  (Domain_d.C:1798-1801, 1813-1817, 1843-1850, 1882-1887)
template <int ET>
WF_DI void shape_derivs(const double (&A)[Elem<ET>::D][Elem<ET>::D], double (&dH)[Elem<ET>::D][Elem<ET>::K]) {
  constexpr int D = Elem<ET>::D;
  constexpr int X = (D == 3) ? 2 : 0; // only indexed inside 3D branches
#pragma unroll
  for (int c = 0; c < D; c++) {
    if constexpr (ET == ET_HEX8) {
      dH[c][0] = 0.125 * (-A[c][0] - A[c][1] - A[c][X]);
      dH[c][1] = 0.125 * (A[c][0] - A[c][1] - A[c][X]);
      dH[c][2] = 0.125 * (A[c][0] + A[c][1] - A[c][X]);
      dH[c][3] = 0.125 * (-A[c][0] + A[c][1] - A[c][X]);
      dH[c][4] = 0.125 * (-A[c][0] - A[c][1] + A[c][X]);
      dH[c][5] = 0.125 * (A[c][0] - A[c][1] + A[c][X]);
      dH[c][6] = 0.125 * (A[c][0] + A[c][1] + A[c][X]);
      dH[c][7] = 0.125 * (-A[c][0] + A[c][1] + A[c][X]);
    } else if constexpr (ET == ET_TET4) {
      dH[c][0] = -A[c][0] - A[c][1] - A[c][X];
      dH[c][1] = A[c][0];
      dH[c][2] = A[c][1];
      dH[c][3] = A[c][X];
    } else if constexpr (ET == ET_QUAD4) {
      dH[c][0] = 0.25 * (-A[c][0] - A[c][1]);
      dH[c][1] = 0.25 * (A[c][0] - A[c][1]);
      dH[c][2] = 0.25 * (A[c][0] + A[c][1]);
      dH[c][3] = 0.25 * (-A[c][0] + A[c][1]);
    } else {
      dH[c][0] = (A[c][0]);
      dH[c][1] = (A[c][1]);
      dH[c][2] = (-A[c][0] - A[c][1]);
    }
  }
}

// radius = (sum_n x_n.r) / k   (Calc_Element_Radius, Domain_d.C:2140-2183)
template <int ET>
WF_DI double elem_radius(const double (&xl)[Elem<ET>::K][Elem<ET>::D]) {
  double r = 0.0;
#pragma unroll
  for (int n = 0; n < Elem<ET>::K; n++) r += xl[n][0];
  return r / (double)Elem<ET>::K;
}

template <int ET> WF_DI double gauss_w() { // Mechanical.C:269-282
  return ET == ET_HEX8 ? 8.0 : (ET == ET_TET4 ? 1.0 / 6.0 : (ET == ET_QUAD4 ? 4.0 : 1.0 / 2.0));
}

// vol = detJ * w * f   (CalcElemVol, Mechanical.C:264-293)
template <int ET>
WF_DI double elem_volume(double detJ, double radius, int domtype, int vol_weight) {
  double f = 1.0;
  if (Elem<ET>::D == 2 && domtype == 2 && vol_weight) f = radius;
  double vol = 0.0;
  vol += detJ * gauss_w<ET>() * f;
  return vol;
}

// D (sym, flat [xx,yy,zz,xy,yz,xz]) and W (upper part [xy,yz,xz])  (calcElemStrainRates, Mechanical.C:41-126)
template <int ET>
WF_DI void strain_rates(const double (&dH)[Elem<ET>::D][Elem<ET>::K], double detJ, const double (&vl)[Elem<ET>::K][Elem<ET>::D],
                        double radius, int domtype, double (&Dr)[6], double (&Wr)[3]) {
  constexpr int D = Elem<ET>::D, K = Elem<ET>::K;
  constexpr int X = (D == 3) ? 2 : 0; // only indexed inside 3D branches
  double dxx = 0.0, dyy = 0.0, dzz = 0.0, dxy = 0.0, dyz = 0.0, dxz = 0.0, wxy = 0.0, wyz = 0.0, wxz = 0.0;
  const double f = 1.0 / detJ;
#pragma unroll
  for (int n = 0; n < K; n++) {
    dxx = dxx + dH[0][n] * f * vl[n][0];
    dyy = dyy + dH[1][n] * f * vl[n][1];
    if constexpr (D == 3) dzz = dzz + dH[X][n] * f * vl[n][X];
    dxy = dxy + f * (dH[1][n] * vl[n][0] + dH[0][n] * vl[n][1]);
    wxy = wxy + f * (dH[1][n] * vl[n][0] - dH[0][n] * vl[n][1]);
    if (D == 2 && domtype == 2) {
      double fa = 0.25;
      if (K == 3) fa = 0.333;
      dzz = dzz + fa * vl[n][0] / radius;
    }
    if constexpr (D == 3) {
      dyz = dyz + f * (dH[X][n] * vl[n][1] + dH[1][n] * vl[n][X]);
      dxz = dxz + f * (dH[X][n] * vl[n][0] + dH[0][n] * vl[n][X]);
      wyz = wyz + f * (dH[X][n] * vl[n][1] - dH[1][n] * vl[n][X]);
      wxz = wxz + f * (dH[X][n] * vl[n][0] - dH[0][n] * vl[n][X]);
    }
  }
  Dr[0] = dxx; Dr[1] = dyy; Dr[2] = dzz;
  Dr[3] = dxy * 0.5; Dr[4] = dyz * 0.5; Dr[5] = dxz * 0.5;
  Wr[0] = wxy * 0.5; Wr[1] = wyz * 0.5; Wr[2] = wxz * 0.5;
}

// calcElemPressure (Mechanical.C:691-819).  J_avg is the mean of the nodal volume ratios; div_v the
// un-normalised sum_a gradN_a . v_a (3D only); is_contact = an element node carries a contact force (:729-747).
WF_DI double pressure_default3d(const WfPar &P, double J_avg, double vol0, double vol1, double rho_e, double div_v,
                                bool is_contact = false) {
  const double K = P.Kbulk;
  if (P.stab_simple) {
    double J_bar = (1 - 0.0) * J_avg; // alpha = 0
    if (J_bar < P.J_min) J_bar = 0.2;
    return -K * ((1.0 - 0.0) * (J_bar - 1.0));
  }
  double J_local = vol1 / vol0;
  double h = pow(vol1, 1.0 / 3.0);
  double alpha = is_contact ? P.alpha_contact : P.alpha_free;
  double J_bar = alpha * J_local + (1 - alpha) * J_avg;
  if (J_bar < P.J_min) J_bar = 0.2;
  double p_physical = -K * (P.log_factor * log(J_bar) + (1.0 - P.log_factor) * (J_bar - 1.0));
  double c = sqrt(K / rho_e);
  double tau = h / (2.0 * c);
  double p_pspg = 0.0;
  double p_hg = (is_contact ? P.hg_coeff_contact : P.hg_coeff_free) * K * fabs(J_local - J_avg);
  double p_q = 0.0;
  if (div_v < 0.0) {
    double a1 = P.pspg_scale * tau * div_v * K, a2 = P.p_pspg_bulkfac * K;
    p_pspg = (a2 < a1) ? a2 : a1;
    double q1 = P.av_coeff_div * rho_e * h * c * (-div_v);
    double delta_J = 1.0 - J_local;
    double q2 = P.av_coeff_bulk * K * delta_J;
    if (is_contact) p_q = 0.5 * (q1 + q2);
    else p_q = (q1 < q2) ? q2 : q1;
  }
  return p_physical + p_pspg + p_hg + p_q;
}

// CalcHollomonYieldStress / TangentModulus (Material.cuh:353-364, 389-395)
WF_DI double hollomon_sy(const WfPar &P, double strain) {
  if (strain + P.eps0 > P.eps1) return P.Kh * pow(strain + P.eps0, P.mh);
  return P.sy0;
}
WF_DI double hollomon_et(const WfPar &P, double strain) {
  if (strain + P.eps0 > P.eps1) return P.Kh * P.mh * pow(strain + P.eps0, (P.mh - 1.0));
  return 0.;
}

// CalcJohnsonCookYieldStress / TangentModulus (Material.cuh:377-387, 397-412); mq = A B n C eps_0 m T_m T_t
WF_DI double jc_sy(const WfPar &P, double strain, double strain_rate, double temp) {
  const double T_h = (temp - P.mq[7]) / (P.mq[6] - P.mq[7]);
  double sr = strain_rate;
  if (strain_rate == 0.0) sr = 1.e-5;
  return (P.mq[0] + P.mq[1] * pow(strain, P.mq[2])) * (1.0 + P.mq[3] * log(sr / P.mq[4])) * (1.0 - pow(T_h, P.mq[5]));
}
WF_DI double jc_et(const WfPar &P, double plstrain, double strain_rate, double temp) {
  const double T_h = (temp - P.mq[7]) / (P.mq[6] - P.mq[7]);
  if (plstrain > 0.)
    return P.mq[2] * P.mq[1] * pow(plstrain, P.mq[2] - 1.) * (1.0 + P.mq[3] * log(strain_rate / P.mq[4])) * (1.0 - pow(T_h, P.mq[5]));
  return P.young * 0.1;
}
// CalcGMTYieldStress / TangentModulus (Material.cuh:418-483); mq = n1 n2 C1 C2 m1 m2 I1 I2 + ranges of e, er, T
WF_DI void gmt_clamp(const WfPar &P, double &e, double &er, double &T) {
  if (e < P.mq[8]) e = P.mq[8]; else if (e > P.mq[9]) e = P.mq[9];
  if (er < P.mq[10]) er = P.mq[10]; else if (er > P.mq[11]) er = P.mq[11];
  if (T < P.mq[12]) T = P.mq[12]; else if (T > P.mq[13]) T = P.mq[13];
}
WF_DI double gmt_sy(const WfPar &P, double strain, double strain_rate, double temp) {
  double e = strain, er = strain_rate, T = temp;
  gmt_clamp(P, e, er, T);
  return P.mq[2] * exp(P.mq[3] * T) * pow(e, P.mq[0] * T + P.mq[1]) * exp((P.mq[6] * T + P.mq[7]) / e) * pow(er, P.mq[4] * T + P.mq[5]);
}
WF_DI double gmt_et(const WfPar &P, double plstrain, double strain_rate, double temp) {
  double e = plstrain, er = strain_rate, T = temp;
  gmt_clamp(P, e, er, T);
  return P.mq[2] * exp(P.mq[3] * T) * pow(er, P.mq[4] * T + P.mq[5]) *
         pow(e, T * P.mq[0] + P.mq[1] - 2.0) * (-P.mq[6] * T - P.mq[7] + e * (P.mq[0] * T + P.mq[1])) * exp((P.mq[6] * T + P.mq[7]) / e);
}

// CalcStressStrain (Mechanical.C:1664-1839): Jaumann rate + J2 radial return.
// tau, eps: flat symmetric; Dr flat symmetric; Wr = (Wxy, Wyz, Wxz).
// The SRT / RS terms follow the nine expressions of tensor3 operator* (Tensor3.C:290-304)
// applied to (tau, Trans(W)) and (W, tau); only the six stored components are formed.
struct StressOut { double sig[6]; double sy; double dep; };
WF_DI void stress_update(const WfPar &P, double dt, double p, const double (&Dr)[6], const double (&Wr)[3],
                         double (&tau)[6], double &pl, double sy_prev, StressOut &o, double temp) {
  // tau: xx=0 yy=1 zz=2 xy=3 yz=4 xz=5
  const double txx = tau[0], tyy = tau[1], tzz = tau[2], txy = tau[3], tyz = tau[4], txz = tau[5];
  const double wxy = Wr[0], wyz = Wr[1], wxz = Wr[2];
  // b = Trans(W): b.xx=b.yy=b.zz=0, b.xy=-wxy, b.xz=-wxz, b.yx=wxy, b.yz=-wyz, b.zx=wxz, b.zy=wyz
  // SRT = a*b with a = tau (symmetric)
  const double srt_xx = txx * 0.0 + txy * wxy + txz * wxz;
  const double srt_xy = txx * wxy + txy * 0.0 + txz * (-wyz);
  const double srt_xz = txx * wxz + txy * wyz + txz * 0.0;
  const double srt_yy = txy * wxy + tyy * 0.0 + tyz * (-wyz);
  const double srt_yz = txy * wxz + tyy * wyz + tyz * 0.0;
  const double srt_zz = txz * wxz + tyz * wyz + tzz * 0.0;
  // RS = a*b with a = W (a.xy=wxy, a.xz=wxz, a.yx=-wxy, a.yz=wyz, a.zx=-wxz, a.zy=-wyz), b = tau
  const double rs_xx = 0.0 * txx + wxy * txy + wxz * txz;
  const double rs_xy = 0.0 * txy + wxy * tyy + wxz * tyz;
  const double rs_xz = 0.0 * txz + wxy * tyz + wxz * tzz;
  const double rs_yy = (-wxy) * txy + 0.0 * tyy + wyz * tyz;
  const double rs_yz = (-wxy) * txz + 0.0 * tyz + wyz * tzz;
  const double rs_zz = (-wxz) * txz + (-wyz) * tyz + 0.0 * tzz;

  const double trD3 = 1.0 / 3.0 * (Dr[0] + Dr[1] + Dr[2]);
  const double g2 = 2.0 * P.G;
  double t[6];
  t[0] = txx + dt * ((Dr[0] - trD3 * 1.) * g2 + srt_xx + rs_xx);
  t[1] = tyy + dt * ((Dr[1] - trD3 * 1.) * g2 + srt_yy + rs_yy);
  t[2] = tzz + dt * ((Dr[2] - trD3 * 1.) * g2 + srt_zz + rs_zz);
  t[3] = txy + dt * ((Dr[3] - trD3 * 0.) * g2 + srt_xy + rs_xy);
  t[4] = tyz + dt * ((Dr[4] - trD3 * 0.) * g2 + srt_yz + rs_yz);
  t[5] = txz + dt * ((Dr[5] - trD3 * 0.) * g2 + srt_xz + rs_xz);

  // Sigma_trial = -p*I + tau ; s = Sigma_trial - (1/3 tr) I
  const double mp = -p;
  double st[6];
  st[0] = mp * 1. + t[0]; st[1] = mp * 1. + t[1]; st[2] = mp * 1. + t[2];
  st[3] = mp * 0. + t[3]; st[4] = mp * 0. + t[4]; st[5] = mp * 0. + t[5];
  const double tr3 = (1.0 / 3.0) * (st[0] + st[1] + st[2]);
  double s[6];
  s[0] = st[0] - tr3 * 1.; s[1] = st[1] - tr3 * 1.; s[2] = st[2] - tr3 * 1.;
  s[3] = st[3] - tr3 * 0.; s[4] = st[4] - tr3 * 0.; s[5] = st[5] - tr3 * 0.;
  const double J2 = 0.5 * (s[0] * s[0] + 2.0 * s[3] * s[3] + 2.0 * s[5] * s[5] + s[1] * s[1] + 2.0 * s[4] * s[4] + s[2] * s[2]);
  const double sig_trial = sqrt(3.0 * J2);

  double sy = sy_prev;
  double esr = 0.0; // effective strain rate (Mechanical.C:1701-1705), only the rate-dependent laws use it
  if (P.model >= 2) {
    esr = sqrt(0.5 * ((Dr[0] - Dr[1]) * (Dr[0] - Dr[1]) + (Dr[1] - Dr[2]) * (Dr[1] - Dr[2]) + (Dr[2] - Dr[0]) * (Dr[2] - Dr[0])) +
               3.0 * (Dr[3] * Dr[3] + Dr[4] * Dr[4] + Dr[5] * Dr[5]));
  }
  if (P.model == 1) sy = hollomon_sy(P, pl);
  else if (P.model == 2) sy = jc_sy(P, pl, esr, temp);
  else if (P.model == 3) sy = gmt_sy(P, pl, esr, temp);
  double dep = 0.0;
  if (P.model >= 2) esr = esr < P.max_edot ? esr : P.max_edot; // min(eff_strain_rate, m_max_edot), :1739
  if (sy < sig_trial) {
    double Et = 0.0; // BILINEAR: uninitialised in the reference (UB); treated as perfectly plastic
    if (P.model == 1) Et = hollomon_et(P, pl);
    else if (P.model == 2) Et = jc_et(P, pl, esr, temp);
    else if (P.model == 3) Et = gmt_et(P, pl, esr, temp);
    const double H = Et, G = P.G;
    const double dgamma = (sig_trial - sy) / (3.0 * G + H);
    const double factor = 1.0 - (3.0 * G * dgamma) / sig_trial;
#pragma unroll
    for (int i = 0; i < 6; i++) t[i] = s[i] * factor;
    pl += dgamma;
    dep = dgamma;
  }
  o.sig[0] = mp * 1. + t[0]; o.sig[1] = mp * 1. + t[1]; o.sig[2] = mp * 1. + t[2];
  o.sig[3] = mp * 0. + t[3]; o.sig[4] = mp * 0. + t[4]; o.sig[5] = mp * 0. + t[5];
#pragma unroll
  for (int i = 0; i < 6; i++) tau[i] = t[i];
  o.sy = sy;
  o.dep = dep;
}

// plastic work rate m_q_plheat (Mechanical.C:1787-1818): plheatfrac * sigma : (strain_pl_incr / dt)
WF_DI double plastic_heat(const WfPar &P, double dt, const StressOut &o) {
  if (!(o.dep > 0.0)) return 0.0;
  const double f = o.dep / o.sy;
  const double (&S)[6] = o.sig; // xx yy zz xy yz xz
  const double idt = 1. / dt;
  const double exx = (f * (S[0] - 0.5 * (S[1] + S[2]))) * idt, eyy = (f * (S[1] - 0.5 * (S[0] + S[2]))) * idt;
  const double ezz = (f * (S[2] - 0.5 * (S[0] + S[1]))) * idt;
  const double exy = (1.5 * f * (S[3])) * idt, exz = (1.5 * f * (S[5])) * idt, eyz = (1.5 * f * (S[4])) * idt;
  return P.plheatfrac * (S[0] * exx + 2.0 * S[3] * exy + 2.0 * S[5] * exz + S[1] * eyy + 2.0 * S[4] * eyz + S[2] * ezz);
}

// calcArtificialViscosity (Mechanical.C:1948-1977): Wilkins q added to the stress diagonal
WF_DI void artificial_viscosity(const WfPar &P, const double (&Dr)[6], double rho_e, double vol, double (&sig)[6]) {
  const double q_max = 1e9;
  double c = sqrt(P.Kbulk / rho_e);
  double eps_v = (Dr[0] + Dr[1] + Dr[2]);
  if (fabs(eps_v) > 1e-12) {
    double l = pow(vol, 1.0 / 3.0);
    double el = eps_v * l;
    double q = P.av_alpha * c * fabs(eps_v) * l + P.av_beta * (el * el); // pow(x,2) == x*x exactly
    q = (q_max < q) ? q_max : q;
    double q_signed = (eps_v > 0) ? -q : q;
    sig[0] += q_signed; sig[1] += q_signed; sig[2] += q_signed;
  }
}

// calcElemForces (Mechanical.C:375-481): f(n,i) = w * [ dH(i,n) s_ii fc + sum_{j!=i} dH(j,n) s_ij ] (+ axisymmetric terms)
template <int ET>
WF_DI void elem_forces(const double (&dH)[Elem<ET>::D][Elem<ET>::K], const double (&sig)[6], double detJ, double radius,
                       int domtype, int vol_weight, double (&f)[Elem<ET>::K][Elem<ET>::D]) {
  constexpr int D = Elem<ET>::D, K = Elem<ET>::K;
  constexpr int X = (D == 3) ? 2 : 0; // only indexed inside 3D branches
  const double w = gauss_w<ET>();
  double fc = 1.0;
  if (D == 2 && domtype == 2 && vol_weight) fc = radius;
#pragma unroll
  for (int n = 0; n < K; n++) {
    double fx = 0.0, fy = 0.0, fz = 0.0;
    fx += dH[0][n] * sig[0] * fc;
    fy += dH[1][n] * sig[1] * fc;
    if constexpr (D == 3) fz += dH[X][n] * sig[2] * fc;
    if constexpr (D == 2) {
      if (domtype != 2) {
        fx += dH[1][n] * sig[3];
        fy += dH[0][n] * sig[3];
      } else {
        const double sigma_rr = sig[0], sigma_tt = sig[2], sigma_rz = sig[3];
        const double ff = detJ / (double)K;
        if (vol_weight) {
          fx += dH[1][n] * sigma_rz * radius + (sigma_rr - sigma_tt) * ff;
          fy += dH[0][n] * sigma_rz * radius + sigma_rz * ff;
        } else {
          const double fa = ff / radius;
          fx += dH[1][n] * sigma_rz - (sigma_rr - sigma_tt) * fa;
          fy += dH[0][n] * sigma_rz - sigma_rz * fa;
        }
      }
    } else {
      fx += dH[1][n] * sig[3] + dH[X][n] * sig[5];
      fy += dH[0][n] * sig[3] + dH[X][n] * sig[4];
      fz += dH[1][n] * sig[4] + dH[0][n] * sig[5];
    }
    f[n][0] = fx * w;
    f[n][1] = fy * w;
    if constexpr (D == 3) f[n][X] = fz * w;
  }
}

// 3D hexa viscous hourglass force (Flanagan-Belytschko / Goudreau), restated from
// f90_ver/src/Mechanical.f90:241-344 (the C++ at this commit has none):
//   hmod(d,j) = sum_n v(n,d) Sig(j,n),  f(n,d) = (0 - sum_j hmod(d,j) Sig(j,n)) * c_h,
//   c_h = coeff * vol**0.6666666 * rho * 0.25 * cs0,   Sig = 4x8 table of +-1.
WF_DI void hexa_hourglass(const WfPar &P, const double (&vl)[8][3], double vol, double rho_e, double (&fh)[8][3]) {
  const double Sig[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1}, {1, -1, -1, 1, -1, 1, 1, -1},
                            {1, -1, 1, -1, 1, -1, 1, -1}, {-1, 1, -1, 1, 1, -1, 1, -1}};
  double hmod[3][4];
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double h = 0.0;
#pragma unroll
      for (int n = 0; n < 8; n++) h = h + vl[n][c] * Sig[j][n];
      hmod[c][j] = h;
    }
  const double c_h = P.hexa_hg * pow(vol, 0.6666666) * rho_e * 0.25 * P.cs0;
#pragma unroll
  for (int n = 0; n < 8; n++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double f = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) f = f - hmod[c][j] * Sig[j][n];
      fh[n][c] = f * c_h;
    }
}

// 2D quad hourglass as shipped (calcElemHourglassForces, Mechanical.C:1842-1943):
// viscous + stiffness form with internal variable hg_q.
WF_DI void quad_hourglass(const WfPar &P, const double (&vl)[4][2], double vol, double rho_e, double (&q)[2], double (&fh)[4][2]) {
  const double Sig[4] = {0.25 * 1, 0.25 * -1, 0.25 * 1, 0.25 * -1};
  double hmod[2] = {0.0, 0.0};
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int n = 0; n < 4; n++) hmod[c] += vl[n][c] * Sig[n];
#pragma unroll
  for (int c = 0; c < 2; c++) q[c] += P.dt * hmod[c];
  const double k_h = P.hg_stiff * P.Kbulk * vol;
  const double c_h = P.hg_visc * rho_e * P.cs0 * pow(vol, 1.0 / 3.0);
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int n = 0; n < 4; n++) {
      double f_visc = -c_h * hmod[c] * Sig[n];
      double f_el = -k_h * q[c] * Sig[n];
      double acc = 0.0;
      acc += f_visc + f_el;
      fh[n][c] = acc;
    }
}

// order-preserving map double -> uint64 for atomicMin
WF_DI unsigned long long dbl_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
WF_DI double key_dbl(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
