/* TEST INFRASTRUCTURE (oracle) — not product code.
 *
 * Plain-C restatement of the reference's CPU algorithm for the explicit
 * Chung-Hulbert step of Domain_d (luchete80/WeldFormFEM @ c68e50e).  Every
 * function cites the reference file:line it follows.  PARITY PINNED: this file
 * is checked bit-for-bit against the compiled, unmodified reference
 * (oracle/_ref/libwf_ref.so, built by oracle/Makefile where /root/reference
 * exists) in tests/test_oracle_vs_ref.py, and against the committed golden
 * vectors in tests/golden/ (generated from that reference build by
 * tests/golden/make_golden.py) everywhere else.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the engine never does.
 */
#ifndef WF_ORACLE_H
#define WF_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wfo_domain wfo_domain;

wfo_domain *wfo_new(void);
void wfo_free(wfo_domain *);
void wfo_set_threads(int n);
int wfo_max_threads(void);
void wfo_set_domtype(wfo_domain *, int domtype, int vol_weight);
void wfo_box(wfo_domain *, const double *V, const double *L, double r, int tritet);
void wfo_set_mesh(wfo_domain *, int dim, int k, int nn, int ne, const double *x, const int *elnod);
void wfo_set_material(wfo_domain *, double E, double nu, double rho0, int model, double sy0, double K, double m);
void wfo_set_material_ext(wfo_domain *, double E, double nu, double rho0, int model, double sy0, const double *q, double temp);
void wfo_set_max_edot(wfo_domain *, double v);
/* thermal coupling: setThermalOn + setTemp + k_T / cp_T / exp_T + plHeatFrac (main.C:218, 436-441, 567-570); contact heat */
void wfo_thermal_on(wfo_domain *, double k_T, double cp_T, double exp_T, double plheatfrac, double T0);
void wfo_set_contact_heat(wfo_domain *, double heat_cond, double T_const);
void wfo_set_stab(wfo_domain *, const double *s12);
void wfo_set_options(wfo_domain *, int press_variant, double av_alpha, double av_beta, double hexa_hg_coeff);
void wfo_add_bc(wfo_domain *, int node, int dim, double val);
void wfo_allocate_bcs(wfo_domain *);
/* contact with rigid surfaces: TriMesh_d::AxisPlaneMesh / AddMesh, raw arrays, main.C:716-725 + :842-847 */
void wfo_add_plane(wfo_domain *, int dimension, int id, int axis, int positaxisorent, const double *p1,
                   const double *p2, int dens, const double *vel);
void wfo_set_trimesh(wfo_domain *, int dimension, int nn, int ne, const double *node, const double *node_v,
                     const int *elnode, const double *normal, const int *mesh_id);
void wfo_contact_on(wfo_domain *, double mu_sta, double mu_dyn, double penalty_factor, double end_time);
void wfo_trimesh_counts(wfo_domain *, int *out3);
void wfo_init(wfo_domain *, double dt);
void wfo_step(wfo_domain *, int n);
double wfo_time_steps(wfo_domain *, int n);
int wfo_call(wfo_domain *, const char *fn, double arg);
long wfo_get(wfo_domain *, const char *name, void *dst, long cap);
long wfo_set(wfo_domain *, const char *name, const void *src, long bytes);
void wfo_info(wfo_domain *, int *out8);
void wfo_consts(wfo_domain *, double *out5);
void wfo_energies(wfo_domain *, double *ekin, double *deint);

#ifdef __cplusplus
}
#endif
#endif
