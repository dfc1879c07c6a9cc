/* TEST INFRASTRUCTURE (oracle) — not product code.  See wf_oracle.h.
 *
 * Plain-C restatement of the reference CPU path of the explicit step of
 * MetFEM::Domain_d.  Arrays use the reference's layouts and member names
 * (include/common/Domain_d.h:837-1044).  Operation order follows the
 * reference expression by expression so that, compiled without FMA
 * contraction, results are bit-identical to the compiled reference.
 */
#define _GNU_SOURCE
#include "wf_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BILINEAR 0
#define HOLLOMON 1
#define JOHNSON_COOK 2
#define GMT 3
#define DOM_PLANE_STRAIN 0
#define DOM_AXISYMM 2
#define DOM_3D 3

struct wfo_domain {
  int dim, k, nn, ne, domtype, vol_weight;
  /* nodes */
  double *x, *v, *a, *u, *u_dt, *prev_a, *m_fi, *m_fe, *m_mdiag, *m_voln, *p_node;
  double *m_voln_0, *m_Jn; /* oracle-only views of calcElemPressure's temporaries */
  /* elements */
  double *dHx, *dHy, *dHz, *m_detJ, *vol, *vol_0, *rho, *rho_0, *p, *pl_strain, *sigma_y, *m_radius;
  double *m_str_rate, *m_rot_rate, *m_sigma, *m_tau, *m_eps, *m_strain_pl_incr;
  double *m_f_elem, *m_f_elem_hg, *m_hg_q;
  unsigned *m_elnod;
  int *m_nodel, *m_nodel_loc, *m_nodel_offset, *m_nodel_count, nodel_tot;
  /* bcs: insertion-ordered lists per dim (Domain_d.C:1057-1061) */
  int *bc_nod[3], bc_count[3], bc_cap[3];
  double *bc_val[3];
  /* material (Material.cuh:15-34, 90-104) */
  int model;
  double E, nu, Kbulk, G, rho0, sy0, Kh, mh, eps0, eps1, cs0;
  int thermal; /* m_thermal */
  double k_T, cp_T, exp_T, plheatfraction, *T, *m_dTedt, *m_q_plheat, *q_cont_conv, tm_heat_cond, tm_T_const;
  double mq[14], temp, max_edot; /* JC: A B n C eps_0 m T_m T_t ; GMT: n1 n2 C1 C2 m1 m2 I1 I2 e_min e_max er_min er_max T_min T_max */
  /* options */
  double stab[12]; /* alpha_free alpha_contact hg_coeff_free hg_coeff_contact av_coeff_div av_coeff_bulk
                      log_factor pspg_scale p_pspg_bulkfac J_min hg_visc hg_stiff */
  int press_variant;
  double av[2], hexa_hg;
  double dt, alpha, beta, gamma, time;
  double *m_elem_length, m_min_length, m_min_height; /* calcMinEdgeLength products */
  /* contact with rigid surfaces (Contact.C, Mesh.h/Mesh.C) */
  int contact, faceCount, *face_nodes /*[faceCount][4]*/, *face_count, *face_elem, *m_mesh_in_contact;
  unsigned char *ext_nodes;
  double *contforce, *ut_prev, *node_area, *m_elem_area;
  double m_contPF, end_t;
  long step_count;
  struct { /* TriMesh_d */
    int dimension, nodecount, elemcount, *elnode, *ele_mesh_id, *nfar;
    double *node, *node_v, *v_orig, *centroid, *normal, *pplane; /* double3 arrays as xyzxyz */
    double mu_sta, mu_dyn;
  } tm;
};

static void *zalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

wfo_domain *wfo_new(void) {
  wfo_domain *d = (wfo_domain *)calloc(1, sizeof(wfo_domain));
  d->dim = 3;
  d->domtype = DOM_3D;
  d->stab[11] = 0.1; /* hg_stiff default, Domain_d.h:294 */
  d->m_contPF = 0.1;  /* Domain_d.h:256 */
  d->max_edot = 1.0e6; /* Domain_d.h:824 */
  d->plheatfraction = 0.9; /* Domain_d.h:271 */
  return d;
}

static void tm_free(wfo_domain *d);
static void free_mesh(wfo_domain *d) {
  double **dp[] = {&d->x, &d->v, &d->a, &d->u, &d->u_dt, &d->prev_a, &d->m_fi, &d->m_fe, &d->m_mdiag, &d->m_voln,
                   &d->p_node, &d->m_voln_0, &d->m_Jn, &d->dHx, &d->dHy, &d->dHz, &d->m_detJ, &d->vol, &d->vol_0,
                   &d->rho, &d->rho_0, &d->p, &d->pl_strain, &d->sigma_y, &d->m_radius, &d->m_str_rate,
                   &d->m_rot_rate, &d->m_sigma, &d->m_tau, &d->m_eps, &d->m_strain_pl_incr, &d->m_f_elem,
                   &d->m_f_elem_hg, &d->m_hg_q};
  for (size_t i = 0; i < sizeof(dp) / sizeof(dp[0]); i++) { free(*dp[i]); *dp[i] = NULL; }
  free(d->m_elnod); d->m_elnod = NULL;
  free(d->T); free(d->m_dTedt); free(d->m_q_plheat); free(d->q_cont_conv);
  d->T = d->m_dTedt = d->m_q_plheat = d->q_cont_conv = NULL;
  free(d->contforce); free(d->ut_prev); free(d->node_area); free(d->m_elem_area); free(d->ext_nodes);
  free(d->m_mesh_in_contact); free(d->face_nodes); free(d->face_count); free(d->face_elem);
  d->contforce = d->ut_prev = d->node_area = d->m_elem_area = NULL; d->ext_nodes = NULL;
  d->m_mesh_in_contact = d->face_nodes = d->face_count = d->face_elem = NULL; d->faceCount = 0; d->contact = 0;
  free(d->m_nodel); free(d->m_nodel_loc); free(d->m_nodel_offset); free(d->m_nodel_count);
  d->m_nodel = d->m_nodel_loc = d->m_nodel_offset = d->m_nodel_count = NULL;
}

void wfo_free(wfo_domain *d) {
  if (!d) return;
  free_mesh(d);
  tm_free(d);
  for (int i = 0; i < 3; i++) { free(d->bc_nod[i]); free(d->bc_val[i]); }
  free(d);
}

void wfo_set_threads(int n) { omp_set_num_threads(n); }
int wfo_max_threads(void) { return omp_get_max_threads(); }

/* Domain_d::setAxiSymm (Domain_d.h:666-670) / m_domtype */
void wfo_set_domtype(wfo_domain *d, int domtype, int vol_weight) {
  d->domtype = domtype;
  if (domtype == DOM_AXISYMM) { d->dim = 2; d->vol_weight = vol_weight; }
}

/* Domain_d::SetDimension (Domain_d.C:457-621), zero-filled */
static void set_dimension(wfo_domain *d, int nn, int ne) {
  free_mesh(d);
  d->nn = nn; d->ne = ne;
  size_t nd = (size_t)nn * d->dim, nk = (size_t)ne * d->k;
  d->x = zalloc(nd, 8); d->v = zalloc(nd, 8); d->a = zalloc(nd, 8); d->u = zalloc(nd, 8);
  d->u_dt = zalloc(nd, 8); d->prev_a = zalloc(nd, 8); d->m_fi = zalloc(nd, 8); d->m_fe = zalloc(nd, 8);
  d->m_mdiag = zalloc(nn, 8); d->m_voln = zalloc(nn, 8); d->p_node = zalloc(nn, 8);
  d->m_voln_0 = zalloc(nn, 8); d->m_Jn = zalloc(nn, 8);
  d->dHx = zalloc(nk, 8); d->dHy = zalloc(nk, 8); d->dHz = zalloc(nk, 8);
  d->m_detJ = zalloc(ne, 8); d->vol = zalloc(ne, 8); d->vol_0 = zalloc(ne, 8); d->rho = zalloc(ne, 8);
  d->rho_0 = zalloc(ne, 8); d->p = zalloc(ne, 8); d->pl_strain = zalloc(ne, 8); d->sigma_y = zalloc(ne, 8);
  d->m_radius = zalloc(ne, 8);
  d->m_str_rate = zalloc(6 * (size_t)ne, 8); d->m_rot_rate = zalloc(6 * (size_t)ne, 8);
  d->m_sigma = zalloc(6 * (size_t)ne, 8); d->m_tau = zalloc(6 * (size_t)ne, 8); d->m_eps = zalloc(6 * (size_t)ne, 8);
  d->m_strain_pl_incr = zalloc(6 * (size_t)ne, 8);
  d->m_f_elem = zalloc(nk * d->dim, 8); d->m_f_elem_hg = zalloc(nk * d->dim, 8);
  d->m_hg_q = zalloc(nk * d->dim, 8);
  d->m_elnod = zalloc(nk, sizeof(unsigned));
  d->contforce = zalloc(nd + 3, 8); d->ut_prev = zalloc(nd + 3, 8); /* +3: the reference writes ut_prev[dim*i+2] in 2D */
  d->node_area = zalloc(nn, 8); d->m_elem_area = zalloc(ne, 8);
  d->T = zalloc(nn > ne ? nn : ne, 8); d->m_dTedt = zalloc(nk > (size_t)nn ? nk : (size_t)nn, 8); d->m_q_plheat = zalloc(ne, 8);
  d->q_cont_conv = zalloc(nn, 8);
  d->ext_nodes = zalloc(nn, 1);
  d->m_mesh_in_contact = zalloc(nn, sizeof(int));
  for (int n = 0; n < nn; n++) d->m_mesh_in_contact[n] = -1;
}

/* Domain_d::setNodElem (Domain_d.C:1508-1611) == tail of AddBoxLength (:1435-1481):
 * count pass, exclusive prefix sum, fill in ascending element id then local node. */
static void set_nod_elem(wfo_domain *d) {
  int nn = d->nn, ne = d->ne, k = d->k;
  d->m_nodel_count = zalloc(nn, sizeof(int));
  d->m_nodel_offset = zalloc(nn, sizeof(int));
  for (int e = 0; e < ne; e++)
    for (int ln = 0; ln < k; ln++) d->m_nodel_count[d->m_elnod[(size_t)e * k + ln]]++;
  int tot = 0;
  for (int n = 0; n < nn; n++) { d->m_nodel_offset[n] = tot; tot += d->m_nodel_count[n]; }
  d->nodel_tot = tot;
  d->m_nodel = zalloc(tot, sizeof(int));
  d->m_nodel_loc = zalloc(tot, sizeof(int));
  for (int n = 0; n < nn; n++) d->m_nodel_count[n] = 0;
  for (int e = 0; e < ne; e++)
    for (int ln = 0; ln < k; ln++) {
      int n = d->m_elnod[(size_t)e * k + ln];
      d->m_nodel[d->m_nodel_offset[n] + d->m_nodel_count[n]] = e;
      d->m_nodel_loc[d->m_nodel_offset[n] + d->m_nodel_count[n]] = ln;
      d->m_nodel_count[n]++;
    }
}

void wfo_set_mesh(wfo_domain *d, int dim, int k, int nn, int ne, const double *x, const int *elnod) {
  d->dim = dim; d->k = k;
  set_dimension(d, nn, ne);
  memcpy(d->x, x, sizeof(double) * (size_t)nn * dim);
  for (size_t i = 0; i < (size_t)ne * k; i++) d->m_elnod[i] = (unsigned)elnod[i];
  set_nod_elem(d);
}

/* Domain_d::AddBoxLength (Domain_d.C:1136-1504) */
void wfo_box(wfo_domain *d, const double *V, const double *L, double r, int tritet) {
  int nel[3];
  d->dim = (L[2] > 0.0) ? 3 : 2;
  nel[0] = (int)(L[0] / (2.0 * r));
  nel[1] = (int)(L[1] / (2.0 * r));
  if (d->dim == 2) { nel[2] = 1; d->k = tritet ? 3 : 4; }
  else { nel[2] = (int)(L[2] / (2.0 * r)); d->k = tritet ? 4 : 8; }
  int nc = (d->dim == 2) ? (nel[0] + 1) * (nel[1] + 1) : (nel[0] + 1) * (nel[1] + 1) * (nel[2] + 1);
  int ne = nel[0] * nel[1] * nel[2];
  if (tritet) ne *= (d->dim == 2) ? 2 : 6;
  set_dimension(d, nc, ne);
  /* coordinates by repeated += 2r accumulation (:1205-1234) */
  int p = 0, kmax = (d->dim == 2) ? 1 : nel[2] + 1, dim = d->dim;
  double Xx, Xy, Xz = V[2];
  for (int kk = 0; kk < kmax; kk++) {
    Xy = V[1];
    for (int j = 0; j < nel[1] + 1; j++) {
      Xx = V[0];
      for (int i = 0; i < nel[0] + 1; i++) {
        d->x[dim * p] = Xx; d->x[dim * p + 1] = Xy;
        if (dim == 3) d->x[dim * p + 2] = Xz;
        p++;
        Xx = Xx + 2.0 * r;
      }
      Xy = Xy + 2.0 * r;
    }
    Xz = Xz + 2 * r;
  }
  unsigned *el = d->m_elnod;
  size_t ei = 0;
  int nx1 = nel[0] + 1;
  if (dim == 2) {
    for (int ey = 0; ey < nel[1]; ey++)
      for (int ex = 0; ex < nel[0]; ex++) {
        int nb1 = nx1 * ey + ex, nb2 = nx1 * (ey + 1) + ex;
        if (!tritet) { /* :1273-1278 */
          el[ei] = nb1; el[ei + 1] = nb1 + 1; el[ei + 2] = nb2 + 1; el[ei + 3] = nb2; ei += 4;
        } else { /* :1296-1303 */
          el[ei] = nb1; el[ei + 1] = nb1 + 1; el[ei + 2] = nb2; ei += 3;
          el[ei] = nb1 + 1; el[ei + 1] = nb2 + 1; el[ei + 2] = nb2; ei += 3;
        }
      }
  } else {
    int nnodz = nx1 * (nel[1] + 1);
    for (int ez = 0; ez < nel[2]; ez++)
      for (int ey = 0; ey < nel[1]; ey++)
        for (int ex = 0; ex < nel[0]; ex++) {
          int nb1 = nnodz * ez + nx1 * ey + ex, nb2 = nnodz * ez + nx1 * (ey + 1) + ex;
          int nh[8] = {nb1, nb1 + 1, nb2 + 1, nb2, nb1 + nnodz, nb1 + nnodz + 1, nb2 + nnodz + 1, nb2 + nnodz};
          if (!tritet) { /* :1321-1332 */
            for (int i = 0; i < 8; i++) el[ei + i] = nh[i];
            ei += 8;
          } else { /* :1383-1388 */
            static const int t[6][4] = {{0, 1, 3, 5}, {1, 2, 3, 5}, {0, 5, 3, 4}, {4, 5, 3, 7}, {5, 6, 3, 7}, {5, 2, 3, 6}};
            for (int q = 0; q < 6; q++) { for (int i = 0; i < 4; i++) el[ei + i] = nh[t[q][i]]; ei += 4; }
          }
        }
  }
  set_nod_elem(d);
}

/* src/explicit/main.C:460-581; Elastic_ (Material.cuh:24-28); InitHollomon (:90-104) */
void wfo_set_material(wfo_domain *d, double E, double nu, double rho0, int model, double sy0, double K, double m) {
  d->E = E; d->nu = nu; d->rho0 = rho0; d->model = model; d->sy0 = sy0;
  d->Kbulk = E / (3.0 * (1.0 - 2.0 * nu));
  d->G = E / (2.0 * (1.0 + nu));
  if (model == HOLLOMON) {
    d->Kh = K; d->mh = m;
    d->eps0 = sy0 / E;
    d->eps1 = pow(sy0 / K, 1. / m);
  }
  d->cs0 = sqrt(d->Kbulk / rho0);
  for (int e = 0; e < d->ne; e++) d->rho_0[e] = rho0; /* setDensity, Domain_d.C:951-958 */
}

/* Johnson-Cook (model 2, q = A B n C eps_0 m T_m T_t) / GMT (model 3, q = n1 n2 C1 C2 m1 m2 I1 I2 + 6 range limits) */
void wfo_set_material_ext(wfo_domain *d, double E, double nu, double rho0, int model, double sy0, const double *q, double temp) {
  wfo_set_material(d, E, nu, rho0, BILINEAR, sy0, 0.0, 1.0);
  d->model = model;
  memcpy(d->mq, q, sizeof(double) * (model == JOHNSON_COOK ? 8 : 14));
  d->temp = temp;
}
void wfo_set_max_edot(wfo_domain *d, double v) { d->max_edot = v; }
void wfo_set_stab(wfo_domain *d, const double *s) { memcpy(d->stab, s, sizeof(d->stab)); }
void wfo_set_options(wfo_domain *d, int press_variant, double av_alpha, double av_beta, double hexa_hg) {
  d->press_variant = press_variant; d->av[0] = av_alpha; d->av[1] = av_beta; d->hexa_hg = hexa_hg;
}

/* AddBCVelNode / AllocateBCs (Domain_d.C:1057-1107) */
void wfo_add_bc(wfo_domain *d, int node, int dim, double val) {
  if (dim < 0 || dim > 2) return;
  if (d->bc_count[dim] == d->bc_cap[dim]) {
    d->bc_cap[dim] = d->bc_cap[dim] ? 2 * d->bc_cap[dim] : 64;
    d->bc_nod[dim] = realloc(d->bc_nod[dim], sizeof(int) * d->bc_cap[dim]);
    d->bc_val[dim] = realloc(d->bc_val[dim], sizeof(double) * d->bc_cap[dim]);
  }
  d->bc_nod[dim][d->bc_count[dim]] = node;
  d->bc_val[dim][d->bc_count[dim]] = val;
  d->bc_count[dim]++;
}
void wfo_allocate_bcs(wfo_domain *d) { (void)d; }

/* ImposeBCV / ImposeBCA (Domain_d.C:1109-1134) */
static void ImposeBCV(wfo_domain *d, int dim) {
  for (int n = 0; n < d->bc_count[dim]; n++) d->v[d->dim * d->bc_nod[dim][n] + dim] = d->bc_val[dim][n];
}
static void ImposeBCA(wfo_domain *d, int dim) {
  for (int n = 0; n < d->bc_count[dim]; n++) d->a[d->dim * d->bc_nod[dim][n] + dim] = 0.0;
}

/* UpdatePrediction (Domain_d.C:961-974) */
static void UpdatePrediction(wfo_domain *d) {
  double dt = d->dt;
#pragma omp parallel for
  for (int i = 0; i < d->nn; i++)
    for (int j = 0; j < d->dim; j++) {
      int ig = i * d->dim + j;
      d->u_dt[ig] = dt * (d->v[ig] + (0.5 - d->beta) * dt * d->prev_a[ig]);
      d->v[ig] += (1.0 - d->gamma) * dt * d->prev_a[ig];
    }
}

/* UpdateCorrectionAccVel (Domain_d.C:981-997) */
static void UpdateCorrectionAccVel(wfo_domain *d) {
  double f = 1.0 / (1.0 - d->alpha), dt = d->dt;
#pragma omp parallel for
  for (int i = 0; i < d->nn; i++)
    for (int j = 0; j < d->dim; j++) {
      int ig = i * d->dim + j;
      d->a[ig] = f * (d->a[ig] - d->alpha * d->prev_a[ig]);
      d->v[ig] += d->gamma * dt * d->a[ig];
    }
}

/* UpdateCorrectionPos (Domain_d.C:1005-1025) */
static void UpdateCorrectionPos(wfo_domain *d) {
  double dt = d->dt;
#pragma omp parallel for
  for (int i = 0; i < d->nn; i++)
    for (int j = 0; j < d->dim; j++) {
      int ig = i * d->dim + j;
      d->u_dt[ig] += d->beta * dt * dt * d->a[ig];
      d->x[ig] += d->u_dt[ig];
      d->prev_a[ig] = d->a[ig];
      d->u[ig] += d->u_dt[ig];
    }
}

/* calcDet (Matrix.h:592-617) */
static double det2(const double A[3][3]) { return A[0][0] * A[1][1] - A[0][1] * A[1][0]; }
static double det3(const double A[3][3]) {
  return A[0][0] * A[1][1] * A[2][2] - A[0][0] * A[1][2] * A[2][1] - A[0][1] * A[1][0] * A[2][2] +
         A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] - A[0][2] * A[1][1] * A[2][0];
}
/* AdjMat (Matrix.h:693-726): 2x2 as shipped (no minus signs, A11 twice); 3x3 = cofactor^T */
static void adj2(const double A[3][3], double R[3][3]) {
  R[0][0] = A[1][1]; R[0][1] = A[1][0];
  R[1][0] = A[0][1]; R[1][1] = A[1][1];
}
static void adj3(const double A[3][3], double R[3][3]) {
  double c[3][3];
  c[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]);
  c[0][1] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]);
  c[0][2] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
  c[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]);
  c[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]);
  c[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]);
  c[2][0] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]);
  c[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]);
  c[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = c[j][i];
}

/* calcElemJAndDerivatives (Domain_d.C:1701-2044), 1 Gauss point */
static void calcElemJAndDerivatives(wfo_domain *d) {
  const int dim = d->dim, k = d->k;
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double x2[8][3] = {{0}}, J[3][3] = {{0}}, A[3][3] = {{0}}, dH[3][8] = {{0}};
    for (int i = 0; i < k; i++)
      for (int c = 0; c < dim; c++) x2[i][c] = d->x[(size_t)dim * d->m_elnod[(size_t)e * k + i] + c];
    if (dim == 2) {
      if (k == 4) { /* :1779-1802 */
        for (int c = 0; c < 2; c++) {
          J[0][c] = 0.25 * (-x2[0][c] + x2[1][c] + x2[2][c] - x2[3][c]);
          J[1][c] = 0.25 * (-x2[0][c] - x2[1][c] + x2[2][c] + x2[3][c]);
        }
        adj2(J, A);
        for (int c = 0; c < 2; c++) {
          dH[c][0] = 0.25 * (-A[c][0] - A[c][1]);
          dH[c][1] = 0.25 * (A[c][0] - A[c][1]);
          dH[c][2] = 0.25 * (A[c][0] + A[c][1]);
          dH[c][3] = 0.25 * (-A[c][0] + A[c][1]);
        }
      } else { /* triangle :1803-1819 */
        for (int c = 0; c < 2; c++) {
          J[0][c] = (x2[0][c] - x2[2][c]);
          J[1][c] = (x2[1][c] - x2[2][c]);
        }
        adj2(J, A);
        for (int c = 0; c < 2; c++) {
          dH[c][0] = (A[c][0]);
          dH[c][1] = (A[c][1]);
          dH[c][2] = (-A[c][0] - A[c][1]);
        }
      }
      d->m_detJ[e] = det2(J);
    } else {
      if (k == 8) { /* :1821-1853 */
        for (int c = 0; c < 3; c++) {
          J[0][c] = 0.125 * (-x2[0][c] + x2[1][c] + x2[2][c] - x2[3][c] - x2[4][c] + x2[5][c] + x2[6][c] - x2[7][c]);
          J[1][c] = 0.125 * (-x2[0][c] - x2[1][c] + x2[2][c] + x2[3][c] - x2[4][c] - x2[5][c] + x2[6][c] + x2[7][c]);
          J[2][c] = 0.125 * (-x2[0][c] - x2[1][c] - x2[2][c] - x2[3][c] + x2[4][c] + x2[5][c] + x2[6][c] + x2[7][c]);
        }
        adj3(J, A);
        for (int c = 0; c < 3; c++) {
          dH[c][0] = 0.125 * (-A[c][0] - A[c][1] - A[c][2]);
          dH[c][1] = 0.125 * (A[c][0] - A[c][1] - A[c][2]);
          dH[c][2] = 0.125 * (A[c][0] + A[c][1] - A[c][2]);
          dH[c][3] = 0.125 * (-A[c][0] + A[c][1] - A[c][2]);
          dH[c][4] = 0.125 * (-A[c][0] - A[c][1] + A[c][2]);
          dH[c][5] = 0.125 * (A[c][0] - A[c][1] + A[c][2]);
          dH[c][6] = 0.125 * (A[c][0] + A[c][1] + A[c][2]);
          dH[c][7] = 0.125 * (-A[c][0] + A[c][1] + A[c][2]);
        }
      } else { /* tetra :1855-1888 */
        for (int c = 0; c < 3; c++) {
          J[0][c] = x2[1][c] - x2[0][c];
          J[1][c] = x2[2][c] - x2[0][c];
          J[2][c] = x2[3][c] - x2[0][c];
        }
        adj3(J, A);
        for (int c = 0; c < 3; c++) {
          dH[c][0] = -A[c][0] - A[c][1] - A[c][2];
          dH[c][1] = A[c][0];
          dH[c][2] = A[c][1];
          dH[c][3] = A[c][2];
        }
      }
      d->m_detJ[e] = det3(J);
    }
    for (int j = 0; j < k; j++) {
      d->dHx[(size_t)e * k + j] = dH[0][j];
      d->dHy[(size_t)e * k + j] = dH[1][j];
      if (dim == 3) d->dHz[(size_t)e * k + j] = dH[2][j];
    }
  }
}

/* Calc_Element_Radius (Domain_d.C:2140-2183) */
static void Calc_Element_Radius(wfo_domain *d) {
  for (int e = 0; e < d->ne; e++) {
    d->m_radius[e] = 0.0;
    for (int ln = 0; ln < d->k; ln++) d->m_radius[e] += d->x[(size_t)d->dim * d->m_elnod[(size_t)e * d->k + ln]];
    d->m_radius[e] /= d->k;
  }
}

static double gauss_w(const wfo_domain *d) { /* Mechanical.C:269-282, 380-389 */
  if (d->dim == 2) return d->k == 4 ? 4 : 1.0 / 2.0;
  return d->k == 4 ? 1.0 / 6.0 : 8.0;
}

/* CalcElemVol (Mechanical.C:264-293) */
static void CalcElemVol(wfo_domain *d) {
  double w = gauss_w(d);
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double f = 1.0;
    if (d->dim == 2 && d->domtype == DOM_AXISYMM && d->vol_weight) f = d->m_radius[e];
    d->vol[e] = 0.0;
    d->vol[e] += d->m_detJ[e] * w * f;
  }
}
static void CalcElemInitialVol(wfo_domain *d) { /* Mechanical.C:343-357 */
  CalcElemVol(d);
  for (int e = 0; e < d->ne; e++) d->vol_0[e] = d->vol[e];
}
static void calcElemDensity(wfo_domain *d) { /* Mechanical.C:295-319 */
  for (int e = 0; e < d->ne; e++) d->rho[e] = d->rho_0[e] * d->vol_0[e] / d->vol[e];
}

/* CalcNodalVol (Mechanical.C:1555-1572) */
static void CalcNodalVol(wfo_domain *d) {
#pragma omp parallel for
  for (int n = 0; n < d->nn; n++) {
    d->m_voln[n] = 0.0;
    for (int e = 0; e < d->m_nodel_count[n]; e++) d->m_voln[n] += d->vol[d->m_nodel[d->m_nodel_offset[n] + e]];
    d->m_voln[n] /= d->k;
  }
}
/* CalcNodalMassFromVol (Mechanical.C:1576-1601) */
static void CalcNodalMassFromVol(wfo_domain *d) {
#pragma omp parallel for
  for (int n = 0; n < d->nn; n++) {
    double mass = 0.0, f = 1.0;
    for (int e = 0; e < d->m_nodel_count[n]; e++) {
      int eg = d->m_nodel[d->m_nodel_offset[n] + e];
      mass += f * d->rho[eg] * d->m_voln[n] / d->m_nodel_count[n];
    }
    d->m_mdiag[n] = mass;
  }
}

#define VEL(e, n, c) d->v[(size_t)dim * d->m_elnod[(size_t)(e)*k + (n)] + (c)]
#define DH(c, e, n) ((c) == 0 ? d->dHx[(size_t)(e)*k + (n)] : ((c) == 1 ? d->dHy[(size_t)(e)*k + (n)] : d->dHz[(size_t)(e)*k + (n)]))

/* calcElemStrainRates (Mechanical.C:41-126) */
static void calcElemStrainRates(wfo_domain *d) {
  const int dim = d->dim, k = d->k;
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double D[3][3] = {{0}}, W[3][3] = {{0}};
    double f = 1.0 / d->m_detJ[e];
    for (int n = 0; n < k; n++) {
      for (int c = 0; c < dim; c++) D[c][c] = D[c][c] + DH(c, e, n) * f * VEL(e, n, c);
      D[0][1] = D[0][1] + f * (DH(1, e, n) * VEL(e, n, 0) + DH(0, e, n) * VEL(e, n, 1));
      W[0][1] = W[0][1] + f * (DH(1, e, n) * VEL(e, n, 0) - DH(0, e, n) * VEL(e, n, 1));
      if (d->domtype == DOM_AXISYMM) {
        double fa = 0.25;
        if (k == 3) fa = 0.333;
        D[2][2] = D[2][2] + fa * VEL(e, n, 0) / d->m_radius[e];
      }
      if (dim == 3) {
        D[1][2] = D[1][2] + f * (DH(2, e, n) * VEL(e, n, 1) + DH(1, e, n) * VEL(e, n, 2));
        D[0][2] = D[0][2] + f * (DH(2, e, n) * VEL(e, n, 0) + DH(0, e, n) * VEL(e, n, 2));
        W[1][2] = W[1][2] + f * (DH(2, e, n) * VEL(e, n, 1) - DH(1, e, n) * VEL(e, n, 2));
        W[0][2] = W[0][2] + f * (DH(2, e, n) * VEL(e, n, 0) - DH(0, e, n) * VEL(e, n, 2));
      }
    }
    D[0][1] *= 0.5; D[0][2] *= 0.5; D[1][2] *= 0.5;
    W[0][1] *= 0.5; W[0][2] *= 0.5; W[1][2] *= 0.5;
    double *sr = d->m_str_rate + 6 * (size_t)e, *rr = d->m_rot_rate + 6 * (size_t)e;
    sr[0] = D[0][0]; sr[1] = D[1][1]; sr[2] = D[2][2]; sr[3] = D[0][1]; sr[4] = D[1][2]; sr[5] = D[0][2];
    rr[0] = 0.0; rr[1] = 0.0; rr[2] = 0.0; rr[3] = W[0][1]; rr[4] = W[1][2]; rr[5] = W[0][2];
  }
}

/* calcElemPressure (Mechanical.C:691-819) */
static void calcElemPressure(wfo_domain *d) {
  const int dim = d->dim, k = d->k;
  const double *s = d->stab;
  double *voln_0 = d->m_voln_0, *voln = d->m_Jn; /* m_Jn holds sum(vol) here; exposed for bisecting */
#pragma omp parallel for
  for (int n = 0; n < d->nn; n++) {
    voln_0[n] = voln[n] = 0.0;
    for (int i = 0; i < d->m_nodel_count[n]; ++i) {
      int e = d->m_nodel[d->m_nodel_offset[n] + i];
      voln_0[n] += d->vol_0[e];
      voln[n] += d->vol[e];
    }
  }
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double K = d->Kbulk, rho_e = d->rho[e], vol0 = d->vol_0[e], vol1 = d->vol[e];
    double J_local = vol1 / vol0;
    double h = pow(vol1, 1.0 / 3.0);
    double J_avg = 0.0;
    for (int a = 0; a < k; ++a) {
      int nid = d->m_elnod[(size_t)e * k + a];
      J_avg += voln[nid] / voln_0[nid];
    }
    J_avg /= k;
    int is_contact = 0; /* Mechanical.C:729-747: any element node with a non-zero contact force */
    if (d->contact)
      for (int a = 0; a < k && !is_contact; ++a) {
        const double *cf = d->contforce + (size_t)dim * d->m_elnod[(size_t)e * k + a];
        double c2 = cf[0] * cf[0] + cf[1] * cf[1] + (dim == 3 ? cf[2] * cf[2] : 0.0 * 0.0);
        if (c2 > 0) is_contact = 1;
      }
    double alpha = is_contact ? s[1] : s[0];
    double J_bar = alpha * J_local + (1 - alpha) * J_avg;
    if (J_bar < s[9]) J_bar = 0.2;
    double p_physical = -K * (s[6] * log(J_bar) + (1.0 - s[6]) * (J_bar - 1.0));
    double c = sqrt(K / rho_e);
    double tau = h / (2.0 * c);
    double div_v = 0.0;
    if (dim > 2)
      for (int a = 0; a < k; ++a)
        div_v += DH(0, e, a) * VEL(e, a, 0) + DH(1, e, a) * VEL(e, a, 1) + DH(2, e, a) * VEL(e, a, 2);
    double p_pspg = 0.0;
    double p_hg = (is_contact ? s[3] : s[2]) * K * fabs(J_local - J_avg);
    double p_q = 0.0;
    if (div_v < 0.0) {
      p_pspg = fmin(s[7] * tau * div_v * K, s[8] * K);
      double q1 = s[4] * rho_e * h * c * (-div_v);
      double delta_J = 1.0 - J_local;
      double q2 = s[5] * K * delta_J;
      if (is_contact) p_q = 0.5 * (q1 + q2);
      else p_q = q1 > q2 ? q1 : q2; /* std::max(q1,q2) */
    }
    d->p[e] = p_physical + p_pspg + p_hg + p_q;
  }
}

/* calcElemPressureLocal (Mechanical.C:1165-1170) */
static void calcElemPressureLocal(wfo_domain *d) {
  for (int e = 0; e < d->ne; e++) d->p[e] = d->Kbulk * (1.0 - d->vol[e] / d->vol_0[e]);
}

/* calcElemPressureANP as shipped (Mechanical.C:1220-1250): accumulates into p without zeroing */
static void calcElemPressureANP(wfo_domain *d) {
  double *pn = malloc(sizeof(double) * d->nn);
  for (int n = 0; n < d->nn; n++) {
    double v0 = 0.0, v1 = 0.0;
    for (int e = 0; e < d->m_nodel_count[n]; e++) {
      int eg = d->m_nodel[d->m_nodel_offset[n] + e];
      v0 += d->vol_0[eg];
      v1 += d->vol[eg];
    }
    pn[n] = d->Kbulk * (1.0 - v1 / v0);
    d->p_node[n] = pn[n];
  }
  for (int e = 0; e < d->ne; e++) {
    for (int ln = 0; ln < d->k; ln++) d->p[e] += pn[d->m_elnod[(size_t)e * d->k + ln]];
    d->p[e] *= 0.25 * d->k;
  }
  free(pn);
}

/* calcElemPressureANP_Nodal (Mechanical.C:1253-1299) */
static void calcElemPressureANP_Nodal(wfo_domain *d) {
  double *pn = malloc(sizeof(double) * d->nn);
  for (int n = 0; n < d->nn; n++) {
    double v0 = 0.0, v1 = 0.0;
    for (int i = 0; i < d->m_nodel_count[n]; ++i) {
      int e = d->m_nodel[d->m_nodel_offset[n] + i];
      v0 += d->vol_0[e] / 4.0;
      v1 += d->vol[e] / 4.0;
    }
    if (v0 > 1e-12) {
      double Jn = v1 / v0;
      pn[n] = d->Kbulk * (1.0 - Jn);
    } else pn[n] = 0.0;
    d->p_node[n] = pn[n];
  }
  for (int e = 0; e < d->ne; e++) {
    d->p[e] = 0.0;
    for (int a = 0; a < d->k; ++a) d->p[e] += pn[d->m_elnod[(size_t)e * d->k + a]];
    d->p[e] /= d->k;
  }
  free(pn);
}

/* TEST-ONLY (press_variant 2): the historical incremental pressure law the validation/ files were printed with,
 * p = -tr(sigma_prev)/3 - K tr(D) dt  — the commented-out Domain_d::calcElemPressure_Hex (Mechanical.C:576-603) =
 * calc_elem_pressure_from_strain (f90_ver/src/Mechanical.f90:525-547), one Gauss point.  It exists so that the
 * restated hexa hourglass can be pinned against every printed digit of validation/1elem_3d_red_int_f_0.06.txt;
 * neither the engine nor the reference's current solver has it. */
static void calcElemPressure_Historical(wfo_domain *d) {
  for (int e = 0; e < d->ne; e++) {
    const double *D = d->m_str_rate + (size_t)6 * e, *sg = d->m_sigma + (size_t)6 * e;
    double press_inc = (D[0] + D[1] + D[2]) * d->dt;
    press_inc = -press_inc / 1.0;
    const double trace = sg[0] + sg[1] + sg[2];
    d->p[e] = -1.0 / 3.0 * trace + d->Kbulk * press_inc;
  }
}

static void pressure(wfo_domain *d) { /* Solver_explicit.C:735-746 */
  if (d->press_variant == 2) { calcElemPressure_Historical(d); return; }
  if (d->press_variant == 0) { if (d->dim == 3) calcElemPressure(d); else calcElemPressureLocal(d); }
  else if (d->press_variant == 1) calcElemPressureANP(d);
  else if (d->press_variant == 3) calcElemPressureANP_Nodal(d);
}

/* calcMinEdgeLength (Domain_d.C:2224-2468): minimum edge length and minimum height; every 3D element is treated
 * as the tetrahedron of its first four nodes (:2243-2247), every 2D element as the quadrilateral of four nodes
 * (:2381-2413).  The angle / Jnorm diagnostics of the same function feed only the remesher and are not restated. */
static double len3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static void calcMinEdgeLength(wfo_domain *d) {
  double min_len = 1.0e6, min_height = 1.0e6;
  if (!d->m_elem_length) d->m_elem_length = (double *)zalloc((size_t)d->ne, sizeof(double));
  for (int e = 0; e < d->ne; e++) {
    double elem_min_height = 1.0e6;
    const unsigned *en = d->m_elnod + (size_t)d->k * e;
    if (d->dim == 3) {
      double P[4][3];
      for (int i = 0; i < 4; i++)
        for (int c = 0; c < 3; c++) P[i][c] = d->x[3 * (size_t)en[i] + c];
      const int ed[6][2] = {{1, 0}, {2, 0}, {3, 0}, {2, 1}, {3, 1}, {3, 2}};
      for (int i = 0; i < 6; i++) {
        double v[3];
        for (int c = 0; c < 3; c++) v[c] = P[ed[i][0]][c] - P[ed[i][1]][c];
        double len = len3(v);
        if (len < min_len) min_len = len;
      }
      const int fc[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
      for (int i = 0; i < 4; i++) {
        double x1[3], x2[3], nrm[3], vec[3];
        for (int c = 0; c < 3; c++) {
          x1[c] = P[fc[i][1]][c] - P[fc[i][0]][c];
          x2[c] = P[fc[i][2]][c] - P[fc[i][0]][c];
        }
        nrm[0] = x1[1] * x2[2] - x1[2] * x2[1];
        nrm[1] = x1[2] * x2[0] - x1[0] * x2[2];
        nrm[2] = x1[0] * x2[1] - x1[1] * x2[0];
        double area = len3(nrm);
        if (area < 1e-12) continue;
        const double inv = 1.0 / area; /* double3 operator/ multiplies by the reciprocal (double3_c.h:95-99) */
        for (int c = 0; c < 3; c++) nrm[c] = nrm[c] * inv;
        for (int c = 0; c < 3; c++) vec[c] = P[i][c] - P[fc[i][0]][c];
        double height = fabs(vec[0] * nrm[0] + vec[1] * nrm[1] + vec[2] * nrm[2]);
        if (height < elem_min_height) elem_min_height = height;
      }
      d->m_elem_length[e] = elem_min_height;
      if (elem_min_height < min_height) min_height = elem_min_height;
    } else {
      double A[2], B[2], Cc[2], D[2];
      for (int c = 0; c < 2; c++) {
        A[c] = d->x[2 * (size_t)en[0] + c]; B[c] = d->x[2 * (size_t)en[1] + c];
        Cc[c] = d->x[2 * (size_t)en[2] + c]; D[c] = d->x[2 * (size_t)en[3] + c];
      }
      double lenAB = sqrt((B[0] - A[0]) * (B[0] - A[0]) + (B[1] - A[1]) * (B[1] - A[1]));
      double lenBC = sqrt((Cc[0] - B[0]) * (Cc[0] - B[0]) + (Cc[1] - B[1]) * (Cc[1] - B[1]));
      double lenCD = sqrt((D[0] - Cc[0]) * (D[0] - Cc[0]) + (D[1] - Cc[1]) * (D[1] - Cc[1]));
      double lenDA = sqrt((A[0] - D[0]) * (A[0] - D[0]) + (A[1] - D[1]) * (A[1] - D[1]));
      min_len = fmin(min_len, fmin(fmin(lenAB, lenBC), fmin(lenCD, lenDA)));
      double area1 = 0.5 * fabs((B[0] - A[0]) * (Cc[1] - A[1]) - (Cc[0] - A[0]) * (B[1] - A[1]));
      double area2 = 0.5 * fabs((Cc[0] - A[0]) * (D[1] - A[1]) - (D[0] - A[0]) * (Cc[1] - A[1]));
      double area = area1 + area2;
      if (area > 1e-14) {
        double h1 = 2.0 * area / lenAB, h2 = 2.0 * area / lenBC, h3 = 2.0 * area / lenCD, h4 = 2.0 * area / lenDA;
        elem_min_height = fmin(fmin(h1, h2), fmin(h3, h4));
        min_height = fmin(min_height, elem_min_height);
      }
      d->m_elem_length[e] = elem_min_height;
    }
  }
  d->m_min_length = min_len;
  d->m_min_height = min_height;
}

/* calcNodalPressureFromElemental (Mechanical.C:1187-1212) */
static void calcNodalPressureFromElemental(wfo_domain *d) {
  double *acc = calloc(d->nn ? d->nn : 1, sizeof(double));
  for (int n = 0; n < d->nn; n++) d->p_node[n] = 0.0;
  for (int e = 0; e < d->ne; e++)
    for (int a = 0; a < d->k; ++a) {
      int nid = d->m_elnod[(size_t)e * d->k + a];
      d->p_node[nid] += d->p[e] * d->vol[e];
      acc[nid] += d->vol[e];
    }
  for (int n = 0; n < d->nn; n++)
    if (acc[n] > 0.0) d->p_node[n] /= acc[n];
  free(acc);
}

typedef struct { double xx, xy, xz, yx, yy, yz, zx, zy, zz; } t3;
/* tensor3 operator* (include/common/Tensor3.C:290-304): NOT a matrix product for
 * non-symmetric operands; coded from its nine expressions. */
static t3 t3mul(t3 a, t3 b) {
  t3 r;
  r.xx = a.xx * b.xx + a.xy * b.yx + a.xz * b.zx;
  r.xy = a.xx * b.yx + a.xy * b.yy + a.xz * b.yz;
  r.xz = a.xx * b.zx + a.xy * b.zy + a.xz * b.zz;
  r.yx = a.yx * b.xx + a.yy * b.yx + a.yz * b.zx;
  r.yy = a.yx * b.yx + a.yy * b.yy + a.yz * b.yz;
  r.yz = a.yx * b.zx + a.yy * b.zy + a.yz * b.zz;
  r.zx = a.zx * b.xx + a.zy * b.yx + a.zz * b.zx;
  r.zy = a.zx * b.yx + a.zy * b.yy + a.zz * b.yz;
  r.zz = a.zx * b.zx + a.zy * b.zy + a.zz * b.zz;
  return r;
}
static t3 t3sym(const double *f) { t3 r = {f[0], f[3], f[5], f[3], f[1], f[4], f[5], f[4], f[2]}; return r; }
static t3 t3anti(const double *f) { t3 r = {f[0], f[3], f[5], -f[3], f[1], f[4], -f[5], -f[4], f[2]}; return r; }
static t3 t3trans(t3 m) { t3 r = {m.xx, m.yx, m.zx, m.xy, m.yy, m.zy, m.xz, m.yz, m.zz}; return r; }
static t3 t3scale(t3 m, double f) { t3 r = {m.xx * f, m.xy * f, m.xz * f, m.yx * f, m.yy * f, m.yz * f, m.zx * f, m.zy * f, m.zz * f}; return r; }
static t3 t3add(t3 a, t3 b) { t3 r = {a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yx + b.yx, a.yy + b.yy, a.yz + b.yz, a.zx + b.zx, a.zy + b.zy, a.zz + b.zz}; return r; }
static t3 t3sub(t3 a, t3 b) { t3 r = {a.xx - b.xx, a.xy - b.xy, a.xz - b.xz, a.yx - b.yx, a.yy - b.yy, a.yz - b.yz, a.zx - b.zx, a.zy - b.zy, a.zz - b.zz}; return r; }
static t3 t3ident(void) { t3 r = {1., 0., 0., 0., 1., 0., 0., 0., 1.}; return r; }
static double t3trace(t3 m) { return (m.xx + m.yy + m.zz); }
static void t3flat(t3 m, double *f) { f[0] = m.xx; f[1] = m.yy; f[2] = m.zz; f[3] = m.xy; f[4] = m.yz; f[5] = m.xz; }

/* CalcHollomonYieldStress / CalcHollomonTangentModulus (Material.cuh:353-364, 389-395) */
static double hollomon_sy(const wfo_domain *d, double strain) {
  if (strain + d->eps0 > d->eps1) return d->Kh * pow(strain + d->eps0, d->mh);
  return d->sy0;
}
static double hollomon_et(const wfo_domain *d, double strain) {
  if (strain + d->eps0 > d->eps1) return d->Kh * d->mh * pow(strain + d->eps0, (d->mh - 1.0));
  return 0.;
}

/* CalcJohnsonCookYieldStress / TangentModulus (Material.cuh:377-387, 397-412) on the public Material_ fields */
static double jc_sy(const wfo_domain *d, double strain, double strain_rate, double temp) {
  const double *q = d->mq; /* A B n C eps_0 m T_m T_t */
  double T_h = (temp - q[7]) / (q[6] - q[7]);
  double sr = strain_rate;
  if (strain_rate == 0.0) sr = 1.e-5;
  return (q[0] + q[1] * pow(strain, q[2])) * (1.0 + q[3] * log(sr / q[4])) * (1.0 - pow(T_h, q[5]));
}
static double jc_et(const wfo_domain *d, double plstrain, double strain_rate, double temp) {
  const double *q = d->mq;
  double T_h = (temp - q[7]) / (q[6] - q[7]);
  if (plstrain > 0.) return q[2] * q[1] * pow(plstrain, q[2] - 1.) * (1.0 + q[3] * log(strain_rate / q[4])) * (1.0 - pow(T_h, q[5]));
  return d->E * 0.1;
}
/* CalcGMTYieldStress / TangentModulus (Material.cuh:418-483) */
static void gmt_clamp(const wfo_domain *d, double *e, double *er, double *T) {
  const double *q = d->mq;
  if (*e < q[8]) *e = q[8]; else if (*e > q[9]) *e = q[9];
  if (*er < q[10]) *er = q[10]; else if (*er > q[11]) *er = q[11];
  if (*T < q[12]) *T = q[12]; else if (*T > q[13]) *T = q[13];
}
static double gmt_sy(const wfo_domain *d, double strain, double strain_rate, double temp) {
  const double *q = d->mq; /* n1 n2 C1 C2 m1 m2 I1 I2 */
  double e = strain, er = strain_rate, T = temp;
  gmt_clamp(d, &e, &er, &T);
  return q[2] * exp(q[3] * T) * pow(e, q[0] * T + q[1]) * exp((q[6] * T + q[7]) / e) * pow(er, q[4] * T + q[5]);
}
static double gmt_et(const wfo_domain *d, double plstrain, double strain_rate, double temp) {
  const double *q = d->mq;
  double e = plstrain, er = strain_rate, T = temp;
  gmt_clamp(d, &e, &er, &T);
  return q[2] * exp(q[3] * T) * pow(er, q[4] * T + q[5]) *
         pow(e, T * q[0] + q[1] - 2.0) * (-q[6] * T - q[7] + e * (q[0] * T + q[1])) * exp((q[6] * T + q[7]) / e);
}

/* CalcStressStrain (Mechanical.C:1664-1839), Hardening plasticity, thermal off */
static void CalcStressStrain(wfo_domain *d, double dt) {
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    size_t ot = 6 * (size_t)e;
    t3 Tau = t3sym(d->m_tau + ot), D = t3sym(d->m_str_rate + ot), W = t3anti(d->m_rot_rate + ot);
    t3 Eps = t3sym(d->m_eps + ot), Epl = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    t3 SRT = t3mul(Tau, t3trans(W));
    t3 RS = t3mul(W, Tau);
    t3 devD = t3sub(D, t3scale(t3ident(), 1.0 / 3.0 * t3trace(D)));
    Tau = t3add(Tau, t3scale(t3add(t3add(t3scale(devD, 2.0 * d->G), SRT), RS), dt));
    t3 Strial = t3add(t3scale(t3ident(), -d->p[e]), Tau);
    t3 s = t3sub(Strial, t3scale(t3ident(), (1.0 / 3.0) * t3trace(Strial)));
    double J2 = 0.5 * (s.xx * s.xx + 2.0 * s.xy * s.xy + 2.0 * s.xz * s.xz + s.yy * s.yy + 2.0 * s.yz * s.yz + s.zz * s.zz);
    double sig_trial = sqrt(3.0 * J2);
    double eff_strain_rate = sqrt(0.5 * ((D.xx - D.yy) * (D.xx - D.yy) + (D.yy - D.zz) * (D.yy - D.zz) + (D.zz - D.xx) * (D.zz - D.xx)) +
                                  3.0 * (D.xy * D.xy + D.yz * D.yz + D.zx * D.zx));
    if (d->model == HOLLOMON) d->sigma_y[e] = hollomon_sy(d, d->pl_strain[e]);
    /* the reference passes T[e]: the NODAL temperature array indexed by the element id (Mechanical.C:1731); defined
     * while e < node count, the uniform set_material_ext temperature stands in beyond that */
    const double temp_e = (d->thermal && e < d->nn) ? d->T[e] : d->temp;
    if (d->model == JOHNSON_COOK) d->sigma_y[e] = jc_sy(d, d->pl_strain[e], eff_strain_rate, temp_e);
    else if (d->model == GMT) d->sigma_y[e] = gmt_sy(d, d->pl_strain[e], eff_strain_rate, temp_e);
    double dep = 0.0;
    eff_strain_rate = eff_strain_rate < d->max_edot ? eff_strain_rate : d->max_edot; /* min(eff_strain_rate, m_max_edot), :1739 */
    if (d->sigma_y[e] < sig_trial) {
      double Et = 0.0; /* BILINEAR: uninitialised in the reference (UB); H = 0 here */
      if (d->model == HOLLOMON) Et = hollomon_et(d, d->pl_strain[e]);
      else if (d->model == JOHNSON_COOK) Et = jc_et(d, d->pl_strain[e], eff_strain_rate, temp_e);
      else if (d->model == GMT) Et = gmt_et(d, d->pl_strain[e], eff_strain_rate, temp_e);
      double H = Et, G = d->G;
      double dgamma = (sig_trial - d->sigma_y[e]) / (3.0 * G + H);
      double factor = 1.0 - (3.0 * G * dgamma) / sig_trial;
      Tau = t3scale(s, factor);
      d->pl_strain[e] += dgamma;
      dep = dgamma;
    }
    t3 Sigma = t3add(t3scale(t3ident(), -d->p[e]), Tau);
    if (d->thermal) d->m_q_plheat[e] = 0.0;
    if (dep > 0.0) {
      double f = dep / d->sigma_y[e];
      Epl.xx = f * (Sigma.xx - 0.5 * (Sigma.yy + Sigma.zz));
      Epl.yy = f * (Sigma.yy - 0.5 * (Sigma.xx + Sigma.zz));
      Epl.zz = f * (Sigma.zz - 0.5 * (Sigma.xx + Sigma.yy));
      Epl.xy = Epl.yx = 1.5 * f * (Sigma.xy);
      Epl.xz = Epl.zx = 1.5 * f * (Sigma.xz);
      Epl.yz = Epl.zy = 1.5 * f * (Sigma.yz);
      if (d->thermal) { /* plastic work rate, Mechanical.C:1798-1818 */
        t3 depdt = t3scale(Epl, 1. / dt);
        d->m_q_plheat[e] = d->plheatfraction * (Sigma.xx * depdt.xx + 2.0 * Sigma.xy * depdt.yx + 2.0 * Sigma.xz * depdt.zx +
                                                Sigma.yy * depdt.yy + 2.0 * Sigma.yz * depdt.yz + Sigma.zz * depdt.zz);
      }
    }
    Eps = t3add(Eps, t3scale(D, dt));
    t3flat(Sigma, d->m_sigma + ot);
    t3flat(Tau, d->m_tau + ot);
    t3flat(Eps, d->m_eps + ot);
    t3flat(Epl, d->m_strain_pl_incr + ot); /* reference stores an uninitialised tensor when dep == 0 */
  }
}

/* calcThermalExpansion (Thermal.C:151-166): the element-node array m_dTedt is read with NODE ids, as the reference does */
static void calcThermalExpansion(wfo_domain *d) {
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double dTdt_gp = 0.0;
    for (int i = 0; i < d->k; ++i) {
      int node = d->m_elnod[(size_t)e * d->k + i];
      dTdt_gp += 1.0 / d->k * d->m_dTedt[node];
    }
    double *sr = d->m_str_rate + 6 * (size_t)e;
    double f = d->exp_T * dTdt_gp;
    sr[0] = sr[0] - f * 1.; sr[1] = sr[1] - f * 1.; sr[2] = sr[2] - f * 1.;
    sr[3] = sr[3] - f * 0.; sr[4] = sr[4] - f * 0.; sr[5] = sr[5] - f * 0.;
  }
}
/* ThermalCalcs (Thermal.C:28-127) */
static void ThermalCalcs(wfo_domain *d) {
  const int k = d->k, dim = d->dim;
  const double w = gauss_w(d);
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double Kt[8][8], Te[8], dTde[8];
    for (int i = 0; i < k; ++i)
      for (int j = 0; j < k; ++j) {
        double kk = 0.0;
        for (int c = 0; c < dim; ++c) kk += DH(c, e, i) * DH(c, e, j);
        Kt[i][j] = kk * d->k_T / d->m_detJ[e] * w;
      }
    for (int i = 0; i < k; ++i) Te[i] = d->T[d->m_elnod[(size_t)e * k + i]];
    double heat = 0.9 * d->m_q_plheat[e];
    double elem_pow = heat * d->vol[e];
    double pow_per_node = elem_pow / k;
    for (int i = 0; i < k; ++i) { dTde[i] = 0.0; for (int j = 0; j < k; ++j) dTde[i] += Kt[i][j] * Te[j]; }
    for (int i = 0; i < k; ++i) {
      int node_id = d->m_elnod[(size_t)e * k + i];
      double m_inv = 1.0 / d->m_mdiag[node_id];
      d->m_dTedt[(size_t)e * k + i] = -m_inv * dTde[i];
      d->m_dTedt[(size_t)e * k + i] += pow_per_node / (d->m_mdiag[node_id] * d->cp_T);
    }
  }
#pragma omp parallel for
  for (int n = 0; n < d->nn; n++) {
    double dTdt = 0;
    for (int e = 0; e < d->m_nodel_count[n]; e++) {
      int eglob = d->m_nodel[d->m_nodel_offset[n] + e], ne = d->m_nodel_loc[d->m_nodel_offset[n] + e];
      dTdt += d->m_dTedt[(size_t)eglob * k + ne];
    }
    d->T[n] += (dTdt + d->q_cont_conv[n] * 1.0 / (d->m_mdiag[n] * d->cp_T)) * d->dt;
  }
}
void wfo_thermal_on(wfo_domain *d, double k_T, double cp_T, double exp_T, double plheatfrac, double T0) {
  d->thermal = 1;
  for (int n = 0; n < d->nn; n++) d->T[n] = T0; /* setTemp, Domain_d.h:678-685 */
  d->k_T = k_T; d->cp_T = cp_T; d->exp_T = exp_T; d->plheatfraction = plheatfrac;
}
void wfo_set_contact_heat(wfo_domain *d, double heat_cond, double T_const) { d->tm_heat_cond = heat_cond; d->tm_T_const = T_const; }

/* calcArtificialViscosity (Mechanical.C:1948-1977) */
static void calcArtificialViscosity(wfo_domain *d) {
  double alpha = d->av[0], beta = d->av[1], q_max = 1e9;
  for (int e = 0; e < d->ne; e++) {
    double c = sqrt(d->Kbulk / d->rho[e]);
    const double *sr = d->m_str_rate + 6 * (size_t)e;
    double eps_v = (sr[0] + sr[1] + sr[2]);
    if (fabs(eps_v) > 1e-12) {
      double l = pow(d->vol[e], 1.0 / 3.0);
      double q = alpha * c * fabs(eps_v) * l + beta * pow(eps_v * l, 2);
      q = q < q_max ? q : q_max;
      double q_signed = (eps_v > 0) ? -q : q;
      d->m_sigma[6 * (size_t)e + 0] += q_signed;
      d->m_sigma[6 * (size_t)e + 1] += q_signed;
      d->m_sigma[6 * (size_t)e + 2] += q_signed;
    }
  }
}

static const int SYM[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}}; /* Domain_d.h:602 */
#define SIG(e, i, j) d->m_sigma[6 * (size_t)(e) + SYM[i][j]]

/* calcElemForces (Mechanical.C:375-481) */
static void calcElemForces(wfo_domain *d) {
  const int dim = d->dim, k = d->k;
  const double w = gauss_w(d);
#pragma omp parallel for
  for (int e = 0; e < d->ne; e++) {
    double *fe = d->m_f_elem + (size_t)e * k * dim;
    for (int i = 0; i < k * dim; i++) fe[i] = 0.0;
    double fc = 1.0;
    if (dim == 2 && d->domtype == DOM_AXISYMM && d->vol_weight) fc = d->m_radius[e];
    for (int n = 0; n < k; n++) {
      for (int c = 0; c < dim; c++) fe[n * dim + c] += DH(c, e, n) * SIG(e, c, c) * fc;
      if (dim == 2) {
        if (d->domtype != DOM_AXISYMM) {
          fe[n * dim] += DH(1, e, n) * SIG(e, 0, 1);
          fe[n * dim + 1] += DH(0, e, n) * SIG(e, 0, 1);
        } else {
          double r_gp = d->m_radius[e], detJ_gp = d->m_detJ[e];
          double sigma_rr = SIG(e, 0, 0), sigma_tt = SIG(e, 2, 2), sigma_rz = SIG(e, 0, 1);
          double f = detJ_gp / (k);
          if (d->vol_weight) {
            fe[n * dim] += DH(1, e, n) * sigma_rz * r_gp + (sigma_rr - sigma_tt) * f;
            fe[n * dim + 1] += DH(0, e, n) * sigma_rz * r_gp + sigma_rz * f;
          } else {
            double fa = f / r_gp;
            fe[n * dim] += DH(1, e, n) * sigma_rz - (sigma_rr - sigma_tt) * fa;
            fe[n * dim + 1] += DH(0, e, n) * sigma_rz - sigma_rz * fa;
          }
        }
      } else {
        fe[n * dim] += DH(1, e, n) * SIG(e, 0, 1) + DH(2, e, n) * SIG(e, 0, 2);
        fe[n * dim + 1] += DH(0, e, n) * SIG(e, 0, 1) + DH(2, e, n) * SIG(e, 1, 2);
        fe[n * dim + 2] += DH(1, e, n) * SIG(e, 1, 2) + DH(0, e, n) * SIG(e, 0, 2);
      }
    }
    for (int i = 0; i < k * dim; i++) fe[i] *= w;
  }
}

/* calcElemHourglassForces: 2D quads as shipped (Mechanical.C:1842-1943); 3D hexa viscous form
 * restated from f90_ver/src/Mechanical.f90:241-344 (absent from the C++ at this commit). */
static void calcElemHourglassForces(wfo_domain *d) {
  const int dim = d->dim, k = d->k;
  if (dim == 2 && k == 4) {
    static const double sg[4] = {1, -1, 1, -1};
    double Sig[4];
    for (int n = 0; n < 4; n++) Sig[n] = 0.25 * sg[n];
    for (int e = 0; e < d->ne; e++) { /* serial: the reference shares hmod across threads */
      double hmod[2] = {0.0, 0.0};
      double *fh = d->m_f_elem_hg + (size_t)e * 8;
      for (int i = 0; i < 8; i++) fh[i] = 0.0;
      for (int c = 0; c < 2; c++)
        for (int n = 0; n < 4; n++) hmod[c] += VEL(e, n, c) * Sig[n];
      for (int c = 0; c < 2; c++) d->m_hg_q[(size_t)e * 2 + c] += d->dt * hmod[c];
      double k_h = d->stab[11] * d->Kbulk * d->vol[e];
      double c_h = d->stab[10] * d->rho[e] * d->cs0 * pow(d->vol[e], 1.0 / 3.0);
      for (int c = 0; c < 2; c++)
        for (int n = 0; n < 4; n++) {
          double sig = Sig[n];
          double f_visc = -c_h * hmod[c] * sig;
          double f_el = -k_h * d->m_hg_q[(size_t)e * 2 + c] * sig;
          fh[n * 2 + c] += f_visc + f_el;
        }
    }
  } else if (dim == 3 && k == 8 && d->hexa_hg != 0.0) {
    static const double Sig[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1}, {1, -1, -1, 1, -1, 1, 1, -1},
                                     {1, -1, 1, -1, 1, -1, 1, -1}, {-1, 1, -1, 1, 1, -1, 1, -1}};
#pragma omp parallel for
    for (int e = 0; e < d->ne; e++) {
      double hmod[3][4] = {{0}}, f[8][3];
      for (int j = 0; j < 4; j++)
        for (int n = 0; n < 8; n++)
          for (int c = 0; c < 3; c++) hmod[c][j] = hmod[c][j] + VEL(e, n, c) * Sig[j][n];
      for (int n = 0; n < 8; n++) {
        for (int c = 0; c < 3; c++) f[n][c] = 0.0;
        for (int j = 0; j < 4; j++)
          for (int c = 0; c < 3; c++) f[n][c] = f[n][c] - hmod[c][j] * Sig[j][n];
      }
      double c_h = d->hexa_hg * pow(d->vol[e], 0.6666666) * d->rho[e] * 0.25 * d->cs0;
      for (int n = 0; n < 8; n++)
        for (int c = 0; c < 3; c++) d->m_f_elem_hg[(size_t)e * 24 + n * 3 + c] = f[n][c] * c_h;
    }
  }
}

/* assemblyForces (Matrices.C:42-87): node-centred gather, element forces first, then
 * hourglass forces subtracted, both in nodel list order. */
static void assemblyForces(wfo_domain *d) {
  const int dim = d->dim, k = d->k;
#pragma omp parallel for
  for (int n = 0; n < d->nn; n++) {
    for (int c = 0; c < dim; c++) d->m_fi[n * dim + c] = 0.0;
    for (int e = 0; e < d->m_nodel_count[n]; e++) {
      int eg = d->m_nodel[d->m_nodel_offset[n] + e], ln = d->m_nodel_loc[d->m_nodel_offset[n] + e];
      size_t off = (size_t)eg * k * dim;
      for (int c = 0; c < dim; c++) d->m_fi[n * dim + c] += d->m_f_elem[off + ln * dim + c];
    }
    for (int e = 0; e < d->m_nodel_count[n]; e++) {
      int eg = d->m_nodel[d->m_nodel_offset[n] + e], ln = d->m_nodel_loc[d->m_nodel_offset[n] + e];
      size_t off = (size_t)eg * k * dim;
      for (int c = 0; c < dim; c++) d->m_fi[n * dim + c] -= d->m_f_elem_hg[off + ln * dim + c];
    }
  }
}


/* ---- contact with rigid surfaces ------------------------------------------------------------------- */
typedef struct { double x, y, z; } v3;
static v3 V3(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vmul(v3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
static v3 vdiv(v3 a, double s) { double inv = 1.0 / s; return vmul(a, inv); } /* double3_c.h:95-99 */
static double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 vcross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static double vlen(v3 a) { return sqrt(vdot(a, a)); }
static v3 ld3(const double *p, int i) { return V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static void st3(double *p, int i, v3 a) { p[3 * i] = a.x; p[3 * i + 1] = a.y; p[3 * i + 2] = a.z; }
static v3 node3(const wfo_domain *d, const double *q, int n) { /* getPosVec3 / getVelVec / getAccVec, Domain_d.h:447-481 */
  return V3(q[d->dim * n], q[d->dim * n + 1], d->dim == 3 ? q[d->dim * n + 2] : 0.0);
}

/* face tables, Domain_d.h:172-193; the constructor selects tetra faces (:246-249), set2DFacesValues quad edges (:785-791) */
static const int TETRA_FACES[4][4] = {{0, 1, 2, -1}, {0, 1, 3, -1}, {1, 2, 3, -1}, {0, 2, 3, -1}};
static const int QUAD_EDGES[4][4] = {{0, 1, -1, -1}, {1, 2, -1, -1}, {2, 3, -1, -1}, {3, 0, -1, -1}};

static int cmp_face(const void *a, const void *b) {
  const int *p = (const int *)a, *q = (const int *)b;
  for (int i = 0; i < 4; i++) if (p[i] != q[i]) return p[i] < q[i] ? -1 : 1;
  return p[4] < q[4] ? -1 : (p[4] > q[4]);
}
/* CalcExtFaceAreas (Domain_d.C:210-315): nodal areas from the faces that occur once, in faceList order */
static void CalcExtFaceAreas(wfo_domain *d) {
  unsigned char *elem_flags = zalloc(d->ne, 1);
  for (int i = 0; i < d->nn; i++) d->node_area[i] = 0.0;
  for (int i = 0; i < d->ne; i++) d->m_elem_area[i] = 0.0;
  for (int i = 0; i < d->faceCount; i++) {
    if (d->face_count[i] != 1) continue;
    const int *fn = d->face_nodes + 4 * i;
    int elem_id = d->face_elem[i];
    if (d->dim == 2) {
      double dx = d->x[2 * fn[1]] - d->x[2 * fn[0]], dy = d->x[2 * fn[1] + 1] - d->x[2 * fn[0] + 1];
      double length = sqrt(dx * dx + dy * dy);
      double share = 0.5 * length;
      d->node_area[fn[0]] += share;
      d->node_area[fn[1]] += share;
      d->m_elem_area[elem_id] += length;
    } else {
      v3 p0 = node3(d, d->x, fn[0]), p1 = node3(d, d->x, fn[1]), p2 = node3(d, d->x, fn[2]);
      v3 cr = vcross(vsub(p1, p0), vsub(p2, p0));
      double area = 0.5 * sqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
      double area_share = area / 3.0;
      d->node_area[fn[0]] += area_share;
      d->node_area[fn[1]] += area_share;
      d->node_area[fn[2]] += area_share;
      if (!elem_flags[elem_id]) { d->m_elem_area[elem_id] = area; elem_flags[elem_id] = 1; }
      else if (area > d->m_elem_area[elem_id]) d->m_elem_area[elem_id] = area;
    }
  }
  free(elem_flags);
}
/* SearchExtNodes (Domain_d.C:110-205).  The reference matches every new face against the whole list (O(F^2));
 * the list it ends with holds each distinct node set once, in order of first occurrence (element, then local
 * face), with its multiplicity and first element.  Restated with a sort: same list, same order. */
static int SearchExtNodes(wfo_domain *d) {
  const int (*tab)[4];
  int elfac, facenod;
  if (d->dim == 3 && d->k == 4) { tab = TETRA_FACES; elfac = 4; facenod = 3; }
  else if (d->dim == 2 && d->k == 4) { tab = QUAD_EDGES; elfac = 4; facenod = 2; }
  else return -1; /* the reference's face tables cover tets (3D) and quads (2D) only */
  size_t nf = (size_t)d->ne * elfac;
  int *rec = malloc(sizeof(int) * 6 * (nf ? nf : 1)); /* sorted node set [4], serial, pad */
  int *raw = malloc(sizeof(int) * 4 * (nf ? nf : 1));
  for (int e = 0; e < d->ne; e++)
    for (int f = 0; f < elfac; f++) {
      size_t i = (size_t)e * elfac + f;
      int q[4] = {-1, -1, -1, -1};
      for (int n = 0; n < facenod; n++) q[n] = raw[4 * i + n] = (int)d->m_elnod[(size_t)e * d->k + tab[f][n]];
      for (int n = facenod; n < 4; n++) raw[4 * i + n] = -1;
      for (int a = 1; a < facenod; a++) for (int b = a; b > 0 && q[b - 1] > q[b]; b--) { int t = q[b]; q[b] = q[b - 1]; q[b - 1] = t; }
      memcpy(rec + 6 * i, q, sizeof(q));
      rec[6 * i + 4] = (int)i;
      rec[6 * i + 5] = 0;
    }
  qsort(rec, nf, 6 * sizeof(int), cmp_face);
  int *first = malloc(sizeof(int) * (nf ? nf : 1)); /* serial -> multiplicity if first of its group, else 0 */
  memset(first, 0, sizeof(int) * nf);
  for (size_t i = 0; i < nf;) {
    size_t j = i + 1;
    while (j < nf && memcmp(rec + 6 * i, rec + 6 * j, 4 * sizeof(int)) == 0) j++;
    first[rec[6 * i + 4]] = (int)(j - i);
    i = j;
  }
  free(d->face_nodes); free(d->face_count); free(d->face_elem);
  d->face_nodes = malloc(sizeof(int) * 4 * (nf ? nf : 1));
  d->face_count = malloc(sizeof(int) * (nf ? nf : 1));
  d->face_elem = malloc(sizeof(int) * (nf ? nf : 1));
  d->faceCount = 0;
  for (size_t i = 0; i < nf; i++)
    if (first[i]) {
      memcpy(d->face_nodes + 4 * d->faceCount, raw + 4 * i, 4 * sizeof(int));
      d->face_count[d->faceCount] = first[i];
      d->face_elem[d->faceCount] = (int)(i / elfac);
      d->faceCount++;
    }
  free(rec); free(raw); free(first);
  for (int n = 0; n < d->nn; n++) d->ext_nodes[n] = 0;
  for (int i = 0; i < d->faceCount; i++)
    if (d->face_count[i] == 1)
      for (int j = 0; j < facenod; j++) d->ext_nodes[d->face_nodes[4 * i + j]] = 1;
  CalcExtFaceAreas(d);
  return 0;
}

/* TriMesh_d::CalcCentroids / CalcNormals / UpdatePlaneCoeff / CalcSpheres / Move (Mesh.h:228-328) */
static void tm_CalcCentroids(wfo_domain *d) {
  int nen = d->tm.dimension == 3 ? 3 : 2;
  for (int e = 0; e < d->tm.elemcount; e++) {
    v3 c = V3(0.0, 0.0, 0.0);
    for (int ne = 0; ne < nen; ne++) c = vadd(c, ld3(d->tm.node, d->tm.elnode[nen * e + ne]));
    st3(d->tm.centroid, e, vdiv(c, nen));
  }
}
static void tm_CalcNormals(wfo_domain *d) {
  if (d->tm.dimension == 3) {
    for (int e = 0; e < d->tm.elemcount; e++) {
      v3 u = vsub(ld3(d->tm.node, d->tm.elnode[3 * e + 1]), ld3(d->tm.node, d->tm.elnode[3 * e]));
      v3 v = vsub(ld3(d->tm.node, d->tm.elnode[3 * e + 2]), ld3(d->tm.node, d->tm.elnode[3 * e]));
      v3 w = vcross(u, v);
      st3(d->tm.normal, e, vdiv(w, vlen(w)));
    }
  } else {
    for (int e = 0; e < d->tm.elemcount; e++) {
      v3 u = vsub(ld3(d->tm.node, d->tm.elnode[2 * e + 1]), ld3(d->tm.node, d->tm.elnode[2 * e]));
      v3 v = V3(-u.y, u.x, 0.0);
      st3(d->tm.normal, e, vdiv(v, vlen(v)));
    }
  }
}
static void tm_UpdatePlaneCoeff(wfo_domain *d) {
  for (int e = 0; e < d->tm.elemcount; e++)
    d->tm.pplane[e] = vdot(ld3(d->tm.node, d->tm.elnode[d->tm.dimension * e + d->tm.nfar[e]]), ld3(d->tm.normal, e));
}
static void tm_CalcSpheres(wfo_domain *d) { /* nfar ends as the last local node whatever the distances (Mesh.h:308-316) */
  for (int e = 0; e < d->tm.elemcount; e++) d->tm.nfar[e] = d->tm.dimension - 1;
  tm_UpdatePlaneCoeff(d);
}
/* Solver_explicit.C:981-1005 */
static void move_trimesh(wfo_domain *d) {
  const double RAMP_FRACTION = 1.0e-2; /* Solver_explicit.C:309 */
  double f = 1.0;
  if (d->time < RAMP_FRACTION * d->end_t) f = pow(d->time / (RAMP_FRACTION * d->end_t), 0.5);
  for (int n = 0; n < d->tm.nodecount; n++) st3(d->tm.node_v, n, vmul(ld3(d->tm.v_orig, n), f));
  for (int n = 0; n < d->tm.nodecount; n++) st3(d->tm.node, n, vadd(ld3(d->tm.node, n), vmul(ld3(d->tm.node_v, n), d->dt)));
  tm_CalcCentroids(d);
  tm_CalcNormals(d);
  tm_UpdatePlaneCoeff(d);
}

static void tm_free(wfo_domain *d) {
  free(d->tm.elnode); free(d->tm.ele_mesh_id); free(d->tm.nfar); free(d->tm.node); free(d->tm.node_v);
  free(d->tm.v_orig); free(d->tm.centroid); free(d->tm.normal); free(d->tm.pplane);
  memset(&d->tm, 0, sizeof(d->tm));
}
static void tm_grow(wfo_domain *d, int nn, int ne, int dimension) {
  int nen = dimension == 3 ? 3 : 2;
  d->tm.dimension = dimension;
  d->tm.node = realloc(d->tm.node, sizeof(double) * 3 * nn);
  d->tm.node_v = realloc(d->tm.node_v, sizeof(double) * 3 * nn);
  d->tm.v_orig = realloc(d->tm.v_orig, sizeof(double) * 3 * nn);
  d->tm.elnode = realloc(d->tm.elnode, sizeof(int) * nen * ne);
  d->tm.ele_mesh_id = realloc(d->tm.ele_mesh_id, sizeof(int) * ne);
  d->tm.nfar = realloc(d->tm.nfar, sizeof(int) * ne);
  d->tm.centroid = realloc(d->tm.centroid, sizeof(double) * 3 * ne);
  d->tm.normal = realloc(d->tm.normal, sizeof(double) * 3 * ne);
  d->tm.pplane = realloc(d->tm.pplane, sizeof(double) * ne);
}

/* TriMesh_d::AxisPlaneMesh (Mesh.C:48-283) appended to the surface set the way main.C:672-708 (first body: every
 * node gets the body velocity) and TriMesh_d::AddMesh (Mesh.C:438-539; further bodies: node ids offset, velocity =
 * m_v of the new mesh) do.  As in the reference the nodes always lie in a z = const (3D) / y = const (2D) plane
 * spanned from p1 with spacing dl = (p2-p1).x / dens whatever `axis` says; `axis` only picks the initial normal. */
void wfo_add_plane(wfo_domain *d, int dimension, int id, int axis, int positaxisorent, const double *p1, const double *p2,
                   int dens, const double *vel) {
  int nn0 = d->tm.nodecount, ne0 = d->tm.elemcount;
  int nn = dimension == 3 ? (dens + 1) * (dens + 1) : dens + 1;
  int ne = dimension == 3 ? dens * dens * 2 : dens;
  tm_grow(d, nn0 + nn, ne0 + ne, dimension);
  double px = p2[0] - p1[0];
  double x2 = p1[1], x3 = p1[2];
  double dl = px / dens;
  int vi = nn0, test = dimension == 2 ? 1 : dens + 1;
  for (int j = 0; j < test; j++) {
    double x1 = p1[0];
    for (int i = 0; i < dens + 1; i++) {
      st3(d->tm.node, vi, V3(x1, x2, x3));
      st3(d->tm.node_v, vi, V3(vel[0], vel[1], vel[2]));
      vi++;
      x1 += dl;
    }
    x2 += dl;
  }
  int el = ne0;
  if (dimension == 3) {
    for (int j = 0; j < dens; j++)
      for (int i = 0; i < dens; i++) {
        int n[4];
        n[0] = (dens + 1) * j + i; n[1] = n[0] + 1; n[2] = (dens + 1) * (j + 1) + i; n[3] = n[2] + 1;
        int elcon[2][3];
        if (positaxisorent) { elcon[0][0] = n[0]; elcon[0][1] = n[1]; elcon[0][2] = n[2]; elcon[1][0] = n[1]; elcon[1][1] = n[3]; elcon[1][2] = n[2]; }
        else { elcon[0][0] = n[0]; elcon[0][1] = n[2]; elcon[0][2] = n[1]; elcon[1][0] = n[1]; elcon[1][1] = n[2]; elcon[1][2] = n[3]; }
        for (int e = 0; e < 2; e++) {
          for (int q = 0; q < 3; q++) d->tm.elnode[3 * el + q] = elcon[e][q] + nn0;
          el++;
        }
      }
  } else {
    for (int i = 0; i < dens; i++) {
      int n0 = i, n1 = i + 1;
      d->tm.elnode[2 * el] = (positaxisorent ? n0 : n1) + nn0;
      d->tm.elnode[2 * el + 1] = (positaxisorent ? n1 : n0) + nn0;
      el++;
    }
  }
  double f = positaxisorent ? 1. : -1.;
  for (int e = ne0; e < ne0 + ne; e++) {
    v3 nrm = V3(0.0, 0.0, 0.0);
    if (dimension == 3) { if (axis == 0) nrm.x = f; else if (axis == 1) nrm.y = f; else nrm.z = f; }
    else { if (axis == 0) nrm.x = f; else if (axis == 1) nrm.y = f; }
    st3(d->tm.normal, e, nrm);
    d->tm.ele_mesh_id[e] = id;
  }
  d->tm.nodecount = nn0 + nn;
  d->tm.elemcount = ne0 + ne;
  tm_CalcCentroids(d);
}

void wfo_set_trimesh(wfo_domain *d, int dimension, int nn, int ne, const double *node, const double *node_v,
                     const int *elnode, const double *normal, const int *mesh_id) {
  tm_free(d);
  tm_grow(d, nn, ne, dimension);
  int nen = dimension == 3 ? 3 : 2;
  memcpy(d->tm.node, node, sizeof(double) * 3 * nn);
  memcpy(d->tm.node_v, node_v, sizeof(double) * 3 * nn);
  memcpy(d->tm.elnode, elnode, sizeof(int) * nen * ne);
  memcpy(d->tm.normal, normal, sizeof(double) * 3 * ne);
  memcpy(d->tm.ele_mesh_id, mesh_id, sizeof(int) * ne);
  d->tm.nodecount = nn; d->tm.elemcount = ne;
  tm_CalcCentroids(d);
}

/* main.C:716-725, :842-847 */
void wfo_contact_on(wfo_domain *d, double mu_sta, double mu_dyn, double pf, double end_time) {
  d->tm.mu_sta = mu_sta; d->tm.mu_dyn = mu_dyn;
  if (pf > -1.0) d->m_contPF = pf;
  tm_CalcSpheres(d);
  d->contact = 1;
  d->end_t = end_time;
}
void wfo_trimesh_counts(wfo_domain *d, int *out) { out[0] = d->tm.dimension; out[1] = d->tm.nodecount; out[2] = d->tm.elemcount; }

/* CalcContactForces (Contact.C:31-336), serial order (the reference's 2D friction reset touches the next node) */
static void CalcContactForces(wfo_domain *d) {
  const int dim = d->dim, nen = d->tm.dimension == 3 ? 3 : 2;
  const double dt = d->dt;
  for (int i = 0; i < d->nn; i++) d->m_mesh_in_contact[i] = -1;
  for (int i = 0; i < d->nn; i++) {
    if (!d->ext_nodes[i]) continue;
    int j = 0, end = d->tm.elemcount == 0;
    while (!end) {
      v3 nj = ld3(d->tm.normal, j);
      v3 xi = node3(d, d->x, i), vi = node3(d, d->v, i), ai = node3(d, d->a, i);
      double dist = vdot(nj, xi) - d->tm.pplane[j];
      v3 x_pred = vadd(vadd(xi, vmul(vi, dt)), vdiv(vmul(vmul(ai, dt), dt), 2.0));
      if (dist < 0) {
        v3 Qj = vsub(xi, vmul(nj, dist)); /* d * normal -> operator*(double, double3) = a*s */
        int inside = 1;
        if (d->tm.dimension == 3) {
          int l = 0, n;
          while (l < 3 && inside) {
            n = l + 1; if (n > 2) n = 0;
            v3 nl = ld3(d->tm.node, d->tm.elnode[nen * j + l]);
            double crit = vdot(vcross(vsub(ld3(d->tm.node, d->tm.elnode[nen * j + n]), nl), vsub(Qj, nl)), nj);
            if (crit < 0.0) inside = 0;
            l++;
          }
        } else {
          int l = 0, n;
          while (l < 2 && inside) {
            n = l + 1; if (n > 1) n = 0;
            v3 nl = ld3(d->tm.node, d->tm.elnode[nen * j + l]);
            double crit = vdot(vsub(ld3(d->tm.node, d->tm.elnode[nen * j + n]), nl), vsub(Qj, nl));
            if (crit < 0.0) inside = 0;
            l++;
          }
        }
        if (inside) {
          double nodlen = 0.0;
          for (int e = 0; e < d->m_nodel_count[i]; e++) nodlen += d->m_elem_length[d->m_nodel[d->m_nodel_offset[i] + e]];
          nodlen /= d->m_nodel_count[i];
          v3 v_rel = vi;
          double v_reln = vdot(v_rel, nj);
          double kcont_geo = d->E * d->node_area[i] / nodlen;
          double kcont_mass = 0.2 * d->m_mdiag[i] / (dt * dt);
          double kcont = kcont_geo < kcont_mass ? kcont_geo : kcont_mass; /* std::min(geo, mass) */
          double damping_ratio = 0.2;
          double omega = sqrt(kcont / d->m_mdiag[i]);
          double ccrit = 2.0 * d->m_mdiag[i] * omega;
          double F_damp = damping_ratio * ccrit * v_reln;
          double F_normal = d->m_contPF * kcont * dist;
          v3 cf = vmul(nj, -(F_normal + F_damp));
          d->contforce[dim * i] = cf.x; d->contforce[dim * i + 1] = cf.y;
          if (dim == 3) d->contforce[dim * i + 2] = cf.z;
          d->m_mesh_in_contact[i] = d->tm.ele_mesh_id[j];
          v3 v_tan = vsub(v_rel, vmul(nj, vdot(v_rel, nj)));
          v3 du_tangent = vmul(v_tan, dt);
          if (dim == 2) du_tangent.z = 0.0;
          double utz = 0.0;
          if (dim == 3) utz = d->ut_prev[3 * i + 2];
          v3 ut_acc = V3(d->ut_prev[dim * i], d->ut_prev[dim * i + 1], utz);
          ut_acc = vadd(ut_acc, du_tangent);
          d->ut_prev[dim * i] += du_tangent.x;
          d->ut_prev[dim * i + 1] += du_tangent.y; d->ut_prev[dim * i + 2] += du_tangent.z;
          v3 Ft_trial = vmul(ut_acc, -kcont);
          double Ft_mag = vlen(Ft_trial);
          v3 Fn = vmul(nj, vdot(cf, nj));
          double normFn = sqrt(Fn.x * Fn.x + Fn.y * Fn.y + Fn.z * Fn.z);
          v3 Ft;
          double Ft_max_static = d->tm.mu_sta * normFn;
          if (Ft_mag <= Ft_max_static) Ft = Ft_trial;
          else {
            double Ft_max_dynamic = d->tm.mu_dyn * normFn;
            double nvt = sqrt(v_tan.x * v_tan.x + v_tan.y * v_tan.y + v_tan.z * v_tan.z);
            Ft = vdiv(vmul(v_tan, -Ft_max_dynamic), nvt);
            for (int c = 0; c < 3; c++) d->ut_prev[dim * i + c] = 0;
          }
          d->contforce[dim * i + 0] += Ft.x;
          d->contforce[dim * i + 1] += Ft.y;
          if (dim == 3) d->contforce[dim * i + 2] += Ft.z;
          d->q_cont_conv[i] = d->tm_heat_cond * d->node_area[i] * (d->tm_T_const - d->T[i]); /* Contact.C:309 */
          end = 1;
        }
      }
      j++;
      if (j == d->tm.elemcount) end = 1;
    }
  }
}

static void scrub_nonfinite(wfo_domain *d) { /* Solver_explicit.C:779-784 */
  for (int i = 0; i < d->nn * d->dim; ++i)
    if (!isfinite(d->m_fi[i])) d->m_fi[i] = 0.0;
}

/* calcAccel (Mechanical.C:321-341) */
static void calcAccel(wfo_domain *d) {
#pragma omp parallel for
  for (int n = 0; n < d->nn; n++) {
    for (int c = 0; c < d->dim; c++) {
      int i = n * d->dim + c;
      d->a[i] = (d->m_fe[i] - d->m_fi[i]) / d->m_mdiag[n];
    }
    if (d->contact)
      for (int c = 0; c < d->dim; c++) {
        int i = n * d->dim + c;
        d->a[i] += d->contforce[i] / d->m_mdiag[n];
      }
  }
}

/* axis constraint (Solver_explicit.C:953-969) */
static void axis_constraint(wfo_domain *d) {
  if (d->domtype != DOM_AXISYMM) return;
  double xmin = 1000.0;
  for (int i = 0; i < d->nn; i++)
    if (d->x[d->dim * i] < xmin) xmin = d->x[d->dim * i];
  for (int i = 0; i < d->nn; i++)
    if (d->x[d->dim * i] <= xmin + 1.e-6) { d->a[d->dim * i] = 0.0; d->v[d->dim * i] = 0.0; }
}

/* computeEnergies (Mechanical.C:2145-2185) */
void wfo_energies(wfo_domain *d, double *ekin, double *deint) {
  double Ekin = 0.0;
  for (int n = 0; n < d->nn; ++n) {
    double vx = d->v[d->dim * n], vy = d->v[d->dim * n + 1], vz = 0.0;
    if (d->dim == 3) vz = d->v[3 * n + 2];
    Ekin += 0.5 * d->m_mdiag[n] * (vx * vx + vy * vy + vz * vz);
  }
  double Edot = 0.0;
  for (int e = 0; e < d->ne; ++e) {
    const double *s = d->m_sigma + 6 * (size_t)e, *r = d->m_str_rate + 6 * (size_t)e;
    double sdot = s[0] * r[0] + s[1] * r[1] + s[2] * r[2] + s[3] * r[3] + s[4] * r[4] + s[5] * r[5];
    Edot += sdot * d->vol[e];
  }
  *ekin = Ekin;
  *deint = Edot * d->dt;
}

/* initialisation: Solver_explicit.C:115-292 (CPU branch) */
void wfo_init(wfo_domain *d, double dt) {
  d->dt = dt;
  for (int e = 0; e < d->ne; e++) { d->pl_strain[e] = 0.0; d->sigma_y[e] = d->sy0; } /* InitValues, Domain_d.C:393-415 */
  for (int c = 0; c < d->dim; c++) { /* only the last dimension's BCs survive (:176-190) */
    for (int n = 0; n < d->nn * d->dim; n++) d->v[n] = d->a[n] = d->u[n] = 0.0;
    ImposeBCV(d, c);
  }
  double rho_b = 0.818200;
  d->alpha = (2.0 * rho_b - 1.0) / (1.0 + rho_b);
  d->beta = (5.0 - 3.0 * rho_b) / ((1.0 + rho_b) * (1.0 + rho_b) * (2.0 - rho_b));
  d->gamma = 1.5 - d->alpha;
  calcElemJAndDerivatives(d);
  if (d->dim == 2 && d->domtype == DOM_AXISYMM) Calc_Element_Radius(d);
  CalcElemInitialVol(d);
  CalcElemVol(d);
  calcElemDensity(d);
  CalcNodalVol(d);
  CalcNodalMassFromVol(d);
  for (int n = 0; n < d->nn * d->dim; n++) d->ut_prev[n] = 0.0; /* Solver_explicit.C:286-289 */
  d->time = 0.0;
  d->step_count = 0;
  if (d->contact) memcpy(d->tm.v_orig, d->tm.node_v, sizeof(double) * 3 * d->tm.nodecount); /* m_v_orig, :168-173 */
}

/* one step: Solver_explicit.C:524-978, rows 1-22 of SURVEY.md §3.3 */
static void step_once(wfo_domain *d) {
  if (d->dim > 2 && d->faceCount > 0 && d->step_count % 10 == 0) CalcExtFaceAreas(d); /* Solver_explicit.C:445-450 */
  UpdatePrediction(d);
  for (int c = 0; c < d->dim; c++) ImposeBCV(d, c);
  calcElemJAndDerivatives(d);
  if (d->dim == 2 && d->domtype == DOM_AXISYMM) Calc_Element_Radius(d);
  CalcElemVol(d);
  CalcNodalVol(d);
  CalcNodalMassFromVol(d);
  calcElemStrainRates(d);
  if (d->thermal) calcThermalExpansion(d); /* Solver_explicit.C:719-720 */
  pressure(d);
  calcNodalPressureFromElemental(d);
  CalcStressStrain(d, d->dt);
  calcArtificialViscosity(d);
  calcElemForces(d);
  calcElemHourglassForces(d);
  if (d->contact) CalcContactForces(d); /* Solver_explicit.C:769-770 */
  assemblyForces(d);
  scrub_nonfinite(d);
  calcAccel(d);
  for (int c = 0; c < d->dim; c++) ImposeBCA(d, c);
  UpdateCorrectionAccVel(d);
  for (int c = 0; c < d->dim; c++) ImposeBCV(d, c);
  axis_constraint(d);
  UpdateCorrectionPos(d);
  if (d->contact) move_trimesh(d);
  if (d->thermal) ThermalCalcs(d); /* Solver_explicit.C:1008-1012 */
  d->time += d->dt;
  d->step_count++;
}

void wfo_step(wfo_domain *d, int n) { for (int i = 0; i < n; i++) step_once(d); }
double wfo_time_steps(wfo_domain *d, int n) {
  double t0 = omp_get_wtime();
  for (int i = 0; i < n; i++) step_once(d);
  return omp_get_wtime() - t0;
}

int wfo_call(wfo_domain *d, const char *f, double arg) {
#define IS(s) (strcmp(f, s) == 0)
  if (IS("UpdatePrediction")) UpdatePrediction(d);
  else if (IS("ImposeBCV")) ImposeBCV(d, (int)arg);
  else if (IS("ImposeBCVAllDim")) { for (int c = 0; c < d->dim; c++) ImposeBCV(d, c); }
  else if (IS("ImposeBCA")) ImposeBCA(d, (int)arg);
  else if (IS("ImposeBCAAllDim")) { for (int c = 0; c < d->dim; c++) ImposeBCA(d, c); }
  else if (IS("calcElemJAndDerivatives")) calcElemJAndDerivatives(d);
  else if (IS("Calc_Element_Radius")) Calc_Element_Radius(d);
  else if (IS("CalcElemVol")) CalcElemVol(d);
  else if (IS("CalcElemInitialVol")) CalcElemInitialVol(d);
  else if (IS("calcElemDensity")) calcElemDensity(d);
  else if (IS("CalcNodalVol")) CalcNodalVol(d);
  else if (IS("CalcNodalMassFromVol")) CalcNodalMassFromVol(d);
  else if (IS("calcElemStrainRates")) calcElemStrainRates(d);
  else if (IS("calcElemPressure")) pressure(d);
  else if (IS("calcNodalPressureFromElemental")) calcNodalPressureFromElemental(d);
  else if (IS("SetDT")) d->dt = arg; /* Domain_d::SetDT, Domain_d.h:635 */
  else if (IS("calcMinEdgeLength")) { if (d->dim == 2 && d->k != 4) return -1; calcMinEdgeLength(d); }
  else if (IS("CalcStressStrain")) CalcStressStrain(d, arg);
  else if (IS("calcArtificialViscosity")) calcArtificialViscosity(d);
  else if (IS("calcElemForces")) calcElemForces(d);
  else if (IS("calcElemHourglassForces")) calcElemHourglassForces(d);
  else if (IS("assemblyForces")) assemblyForces(d);
  else if (IS("calcAccel")) calcAccel(d);
  else if (IS("UpdateCorrectionAccVel")) UpdateCorrectionAccVel(d);
  else if (IS("AxisConstraint")) axis_constraint(d);
  else if (IS("UpdateCorrectionPos")) UpdateCorrectionPos(d);
  else if (IS("calcThermalExpansion")) calcThermalExpansion(d);
  else if (IS("ThermalCalcs")) ThermalCalcs(d);
  else if (IS("SearchExtNodes")) return SearchExtNodes(d);
  else if (IS("CalcExtFaceAreas")) CalcExtFaceAreas(d);
  else if (IS("CalcContactForces")) CalcContactForces(d);
  else if (IS("MoveTriMesh")) move_trimesh(d);
  else return -1;
#undef IS
  return 0;
}

typedef struct { void *ptr; size_t bytes; } view_t;
static view_t view(wfo_domain *d, const char *nm) {
  size_t nd = 8 * (size_t)d->nn * d->dim, nn = 8 * (size_t)d->nn, ne = 8 * (size_t)d->ne, nk = ne * d->k;
  view_t z = {NULL, 0};
#define V(s, p, b) if (strcmp(nm, s) == 0) { view_t r = {(void *)(p), (b)}; return r; }
  V("x", d->x, nd) V("v", d->v, nd) V("a", d->a, nd) V("u", d->u, nd) V("u_dt", d->u_dt, nd)
  V("prev_a", d->prev_a, nd) V("m_fi", d->m_fi, nd) V("m_fe", d->m_fe, nd)
  V("m_mdiag", d->m_mdiag, nn) V("m_voln", d->m_voln, nn) V("p_node", d->p_node, nn)
  V("m_voln_0", d->m_voln_0, nn) V("m_Jn", d->m_Jn, nn)
  V("m_dH_detJ_dx", d->dHx, nk) V("m_dH_detJ_dy", d->dHy, nk) V("m_dH_detJ_dz", d->dHz, nk)
  V("m_detJ", d->m_detJ, ne) V("vol", d->vol, ne) V("vol_0", d->vol_0, ne) V("rho", d->rho, ne)
  V("rho_0", d->rho_0, ne) V("p", d->p, ne) V("pl_strain", d->pl_strain, ne) V("sigma_y", d->sigma_y, ne)
  V("m_radius", d->m_radius, ne)
  V("m_str_rate", d->m_str_rate, 6 * ne) V("m_rot_rate", d->m_rot_rate, 6 * ne) V("m_sigma", d->m_sigma, 6 * ne)
  V("m_tau", d->m_tau, 6 * ne) V("m_eps", d->m_eps, 6 * ne)
  V("m_f_elem", d->m_f_elem, nk * d->dim) V("m_f_elem_hg", d->m_f_elem_hg, nk * d->dim)
  V("m_hg_q", d->dim == 2 ? d->m_hg_q : NULL, d->dim == 2 ? nk * d->dim : 0)
  V("m_elem_length", d->m_elem_length, d->m_elem_length ? ne : 0)
  V("T", d->T, nn) V("m_dTedt", d->m_dTedt, nk) V("m_q_plheat", d->m_q_plheat, ne) V("q_cont_conv", d->q_cont_conv, nn)
  V("contforce", d->contforce, nd) V("ut_prev", d->ut_prev, nd) V("node_area", d->node_area, nn)
  V("m_elem_area", d->m_elem_area, ne) V("ext_nodes", d->ext_nodes, (size_t)d->nn)
  V("m_mesh_in_contact", d->m_mesh_in_contact, sizeof(int) * (size_t)d->nn)
  if (d->tm.nodecount) {
    size_t tn = 24 * (size_t)d->tm.nodecount, te = (size_t)d->tm.elemcount;
    V("trimesh.node", d->tm.node, tn) V("trimesh.node_v", d->tm.node_v, tn)
    V("trimesh.normal", d->tm.normal, 24 * te) V("trimesh.pplane", d->tm.pplane, 8 * te)
    V("trimesh.elnode", d->tm.elnode, sizeof(int) * te * (d->tm.dimension == 3 ? 3 : 2))
    V("trimesh.ele_mesh_id", d->tm.ele_mesh_id, sizeof(int) * te)
  }
  V("bcx_val", d->bc_val[0], sizeof(double) * (size_t)d->bc_count[0]) /* Domain_d.h:901 */
  V("bcy_val", d->bc_val[1], sizeof(double) * (size_t)d->bc_count[1])
  V("bcz_val", d->bc_val[2], sizeof(double) * (size_t)d->bc_count[2])
  V("m_elnod", d->m_elnod, sizeof(unsigned) * (size_t)d->ne * d->k)
  V("m_nodel", d->m_nodel, sizeof(int) * (size_t)d->nodel_tot)
  V("m_nodel_loc", d->m_nodel_loc, sizeof(int) * (size_t)d->nodel_tot)
  V("m_nodel_offset", d->m_nodel_offset, sizeof(int) * (size_t)d->nn)
  V("m_nodel_count", d->m_nodel_count, sizeof(int) * (size_t)d->nn)
#undef V
  return z;
}

long wfo_get(wfo_domain *d, const char *name, void *dst, long cap) {
  view_t w = view(d, name);
  if (!w.ptr) return -1;
  if ((long)w.bytes > cap) return -(long)w.bytes;
  memcpy(dst, w.ptr, w.bytes);
  return (long)w.bytes;
}
long wfo_set(wfo_domain *d, const char *name, const void *src, long bytes) {
  view_t w = view(d, name);
  if (!w.ptr || (long)w.bytes != bytes) return -1;
  memcpy(w.ptr, src, w.bytes);
  return bytes;
}
void wfo_info(wfo_domain *d, int *out) {
  out[0] = d->dim; out[1] = d->k; out[2] = d->nn; out[3] = d->ne;
  out[4] = d->bc_count[0]; out[5] = d->bc_count[1]; out[6] = d->bc_count[2]; out[7] = d->domtype;
}
void wfo_consts(wfo_domain *d, double *out) {
  out[0] = d->alpha; out[1] = d->beta; out[2] = d->gamma; out[3] = d->dt; out[4] = d->time;
  out[5] = d->m_min_length; out[6] = d->m_min_height;
}
