/* wf_engine.h — C ABI of the B200-native explicit time-step engine for WeldFormFEM.
 *
 * The reference (luchete80/WeldFormFEM @ c68e50e) has no plugin ABI: the seam is
 * the MetFEM::Domain_d object — its public setters, its flat SOA arrays
 * (include/common/Domain_d.h:837-1044) and the fixed call sequence of
 * Domain_d::SolveChungHulbert() (src/explicit/Solver_explicit.C:115-292 init,
 * :524-978 one step).  Every entry point below names the reference interface it
 * replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; wf_last_error()
 *     returns the message.  The engine has NO CPU fallback: without a CUDA
 *     device wf_create() fails.
 *   - host arrays passed in / out use the REFERENCE layouts and member names
 *     ("xyzxyz" nodal vectors, 6-vectors [xx,yy,zz,xy,yz,xz] per element,
 *     m_f_elem[e][local node][dim], ...).  Device-side layouts are private.
 *   - indices are 32-bit like the reference (unsigned m_elnod, int m_nodel*).
 */
#ifndef WF_ENGINE_H
#define WF_ENGINE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wf_engine wf_engine; /* opaque; owns all device memory of one GPU */

/* dom_type, include/common/Domain_d.h:101 */
enum { WF_PLANE_STRAIN = 0, WF_PLANE_STRESS = 1, WF_AXISYMM = 2, WF_3D = 3 };
/* Material_model, include/common/Material.cuh:9-13 */
enum { WF_BILINEAR = 0, WF_HOLLOMON = 1, WF_JOHNSON_COOK = 2, WF_GMT = 3 };
/* pressure law selected per step (Solver_explicit.C:735-746):
 *   0 = m_press_algorithm 0: calcElemPressure (3D, Mechanical.C:691) / calcElemPressureLocal (2D, :1165)
 *   1 = m_press_algorithm 1: calcElemPressureANP as shipped (Mechanical.C:1220; accumulates)
 *   3 = calcElemPressureANP_Nodal (Mechanical.C:1253; corrected ANP, not wired in the reference solver) */
enum { WF_PRESS_DEFAULT = 0, WF_PRESS_ANP_SHIPPED = 1, WF_PRESS_ANP_NODAL = 3 };
/* numerics flavour */
enum {
  WF_STRICT = 1, /* operation-for-operation with the reference CPU path, no FMA contraction,
                    separate m_f_elem_hg and two-pass assembly (Matrices.C:51-75) */
  WF_FAST = 0    /* same algorithm, FMA + algebraically equivalent regrouping, hourglass force
                    folded into m_f_elem; meets the 1e-10 / 1e-6 tolerances of BASELINE.json */
};

/* Material_ / Elastic_ (include/common/Material.cuh:15-156) as read by the hot path */
typedef struct wf_material {
  int model;     /* WF_BILINEAR | WF_HOLLOMON */
  double E, nu;  /* Elastic_(E,nu): K = E/(3(1-2nu)), G = E/(2(1+nu)) */
  double rho0;   /* setDensity(), Domain_d.C:951 */
  double sy0;    /* yieldStress0 */
  double K, m;   /* Hollomon constants; eps0 = sy0/E, eps1 = pow(sy0/K,1/m) (Material.cuh:90-104) */
  /* WF_JOHNSON_COOK: q[0..7] = A B n C eps_0 m T_m T_t (Material_::Init_JohnsonCook, Material.cuh:105-116);
   * WF_GMT: q[0..13] = n1 n2 C1 C2 m1 m2 I1 I2 e_min e_max er_min er_max T_min T_max (GMT, Material.cuh:237-262).
   * These are the PUBLIC Material_ fields the free functions CalcJohnsonCook* / CalcGMT* (Material.cuh:377-483) read. */
  double q[14];
  double temp;     /* temperature seen by the flow stress: the reference reads T[e] (Mechanical.C:1731), uniform with thermal off */
  double max_edot; /* m_max_edot (Domain_d.h:824, config "maxStrRate"); <= 0 selects the default 1e6 */
} wf_material;

/* StabilizationParams (include/common/Domain_d.h:140-153) + the hexa hourglass coefficient of
 * f90_ver/src/Mechanical.f90:307 (0.06; 0 = off = behaviour of the C++ at this commit) */
typedef struct wf_stab {
  double alpha_free, alpha_contact, hg_coeff_free, hg_coeff_contact, av_coeff_div, av_coeff_bulk;
  double log_factor, pspg_scale, p_pspg_bulkfac, J_min, hg_visc, hg_stiff;
  double hexa_hg_coeff;
} wf_stab;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* new Domain_d + setAxiSymm/m_domtype (Domain_d.h:241-353, :666); device = CUDA ordinal */
int wf_create(wf_engine **out, int dim, int nodxelem, int domtype, int device);
void wf_destroy(wf_engine *);
const char *wf_last_error(wf_engine *); /* engine may be NULL: last error of wf_create */
/* run all engine work on a caller-owned CUDA stream (cudaStream_t as void*); default stream if never set */
int wf_set_stream(wf_engine *, void *cuda_stream);
int wf_get_stream(wf_engine *, void **cuda_stream);
int wf_synchronize(wf_engine *);

/* ---- mesh ---------------------------------------------------------------------------------------- */
/* SetDimension + copy x + setNodElem (Domain_d::CreateFromLSDyna, Domain_d.C:1647-1699; setNodElem :1508-1611) */
int wf_set_mesh(wf_engine *, int n_nodes, int n_elems, const double *x /*N_n*dim*/, const unsigned *elnod /*N_e*k*/);
/* Domain_d::AddBoxLength (Domain_d.C:1136-1504): same nel, numbering, coordinates by accumulation, tet/tri split */
int wf_gen_box(wf_engine *, const double V[3], const double L[3], double r, int tritet);
int wf_get_counts(wf_engine *, int *n_nodes, int *n_elems, int *nodel_total);
/* Internal element order used by the NEXT wf_set_mesh / wf_gen_box / wf_set_mesh_partition (see wf_host_elem_order):
 * 0 = the caller's numbering, 1 = Morton order.  Default: 1 for hexahedra, 0 for every other element type (measured:
 * tetrahedra and quadrilaterals gather per element and lose from the reordering); the environment variable
 * WF_ELEM_ORDER overrides the default.  Every array that crosses this ABI stays in the caller's numbering whatever the mode; wf_device_ptr
 * of an element array exposes the internal order ("elem_perm" via wf_get_array gives perm[internal] = user). */
int wf_set_elem_order(wf_engine *, int mode);
/* Thread layout of the brick form of the hexahedron passes on the current mesh: n_cta = CTAs of 128 thread slots (0: the
 * brick form is not in use), plan = 1 when the slots follow the cells of the mesh (wf_host_brick_plan: every CTA a
 * clipped 8x4x4 brick), 0 for the compact numbering.  The environment variable WF_BRICK_PLAN=0 forces the latter. */
int wf_brick_info(wf_engine *, int *n_cta, int *plan);
int wf_set_axisymm_vol_weight(wf_engine *, int on); /* setAxiSymm(vol_weight), Domain_d.h:666 */

/* ---- material / options / boundary conditions -------------------------------------------------- */
int wf_set_material(wf_engine *, const wf_material *); /* main.C:460-581 + AssignMaterial (Domain_d.C:903) */
int wf_set_stab(wf_engine *, const wf_stab *);          /* m_stab, main.C:84-120 */
/* m_press_algorithm, m_artifvisc[0..1] (main.C:211-347), numerics flavour */
int wf_set_options(wf_engine *, int press_algorithm, double av_alpha, double av_beta, int strict_reference);
/* optional per-step products the reference always keeps: bit0 = integrate m_eps (Mechanical.C:1822),
 * bit1 = store m_sigma every step instead of rebuilding it on wf_get_array (identical values) */
int wf_set_tracking(wf_engine *, int flags);
/* thermal coupling (config "thermal"): setThermalOn + setTemp(T0) (main.C:436-441), Material_::k_T / cp_T / exp_T
 * (thermalCond / thermalHeatCap / thermalExp, main.C:567-570) and m_plheatfraction (plHeatFrac, main.C:218).  Per step:
 * calcThermalExpansion (Thermal.C:151-166), the plastic work rate m_q_plheat (Mechanical.C:1784-1818) and ThermalCalcs
 * (Thermal.C:28-127), fused into the element and node passes.  Arrays: "T", "m_q_plheat", "m_dTedt", "q_cont_conv". */
int wf_set_thermal(wf_engine *, double k_T, double cp_T, double exp_T, double plheatfrac, double T0);
int wf_add_bc_vel(wf_engine *, int node, int dim, double val); /* AddBCVelNode, Domain_d.C:1057 */
int wf_add_bc_vel_array(wf_engine *, int count, const int *node, const int *dim, const double *val);
int wf_allocate_bcs(wf_engine *);                              /* AllocateBCs, Domain_d.C:1063 */
/* overwrite bcx_val / bcy_val / bcz_val (Domain_d.h:901) of one dimension, insertion order, from host memory;
 * asynchronous on the engine's stream (time-dependent prescribed velocities) */
int wf_set_bc_values(wf_engine *, int dim, int count, const double *vals);

/* ---- contact with rigid tool surfaces (tetrahedra in 3D, quadrilaterals in 2D like the reference) --- */
/* Domain_d::SearchExtNodes (Domain_d.C:110-205): external faces / nodes + CalcExtFaceAreas; call after the mesh is set */
int wf_SearchExtNodes(wf_engine *);
int wf_CalcExtFaceAreas(wf_engine *); /* Domain_d.C:210-315; the step recomputes it every 10th step in 3D (Solver_explicit.C:445-450) */
/* Domain_d::setTriMesh with the flattened TriMesh_d (include/common/Mesh.h:72-140) the reference holds after
 * AxisPlaneMesh / AddMesh (main.C:672-708, :775-828): node and node_v as xyz triples, elnode with 3 (3D) or 2 (2D) node
 * ids per facet, initial facet normals, ele_mesh_id.  The engine moves the surfaces itself every step
 * (velocity ramp, Move, CalcNormals, UpdatePlaneCoeff; Solver_explicit.C:981-1005). */
int wf_set_trimesh(wf_engine *, int dimension, int n_nodes, int n_elems, const double *node, const double *node_v,
                   const int *elnode, const double *normal, const int *ele_mesh_id);
/* friction coefficients mu_sta[0] / mu_dyn[0], setContactPF (penalty_factor <= -1 keeps the default 0.1), CalcSpheres +
 * setContactOn (main.C:716-725, :842-847) and SetEndTime (the surface velocity ramps over the first 1 % of end_time);
 * needs wf_calcMinEdgeLength (m_elem_length, main.C:862) before wf_init */
int wf_set_contact(wf_engine *, double mu_sta, double mu_dyn, double penalty_factor, double end_time);
/* TriMesh_d::heat_cond / T_const (heatCondCoeff, dieTemp; main.C:718-719): heat flow q_cont_conv = heat_cond * node_area *
 * (T_const - T) into every node in contact (Contact.C:309); needs wf_set_thermal */
int wf_set_contact_heat(wf_engine *, double heat_cond, double T_const);
int wf_CalcContactForces(wf_engine *); /* Contact.C:31-336, unfused entry point */
int wf_MoveTriMesh(wf_engine *);       /* Solver_explicit.C:981-1005, unfused entry point */
int wf_get_trimesh_counts(wf_engine *, int *dimension, int *n_nodes, int *n_elems);

/* ---- solve --------------------------------------------------------------------------------------- */
int wf_init(wf_engine *, double dt);  /* Solver_explicit.C:115-292 incl. SetDT; CH constants rho_b = 0.8182 */
int wf_step(wf_engine *, int nsteps); /* fused rows 1-22 of the step, Solver_explicit.C:524-978 */
/* wf_step for a host loop that exchanges data with the engine EVERY step (the reference's loop rewrites bcx_val / reads
 * energies per step, Solver_explicit.C:524-1170): same steps, but the engine is left in predicted state — the next
 * step's UpdatePrediction + ImposeBCV (Solver_explicit.C:524-540) have already run inside the last node pass — so
 * single-step calls keep the fused schedule of a batch.  In that state only wf_set_bc_values (patches the predicted
 * velocities of the prescribed components), wf_monitor_async / wf_monitor_wait, wf_step, wf_step_open and
 * wf_step_close may be called.  wf_step_close undoes the prediction so that every array can be read; fast flavour. */
int wf_step_open(wf_engine *, int nsteps);
int wf_step_close(wf_engine *);
/* set after a step that scrubbed a non-finite internal force (Solver_explicit.C:779-784); clears the flag */
int wf_nonfinite_flag(wf_engine *, int *flag);
int wf_energies(wf_engine *, double *Ekin, double *dEint); /* computeEnergies, Mechanical.C:2145 */
int wf_get_time(wf_engine *, double *time, long *step_count);
/* restart / remesh hand-off: continue the clock of a previous engine (Domain_d::Time, step_count; the velocity ramp of
 * the rigid surfaces and the CalcExtFaceAreas cadence depend on them, Solver_explicit.C:309, :445-450) */
int wf_set_time(wf_engine *, double time, long step_count);
/* diagnostics computed on the device: calcMinEdgeLength (Domain_d.C:2224; also fills "m_elem_length"), max |v| and the
 * variable step of the explicit loop dt = cfl * min_length / (cs + max|v|) (Solver_explicit.C:579-598); wf_set_dt
 * applies a new step between batches ("p_node", calcNodalPressureFromElemental Mechanical.C:1187, is produced by
 * wf_get_array on request) */
int wf_calcMinEdgeLength(wf_engine *, double *min_length, double *min_height);
int wf_max_velocity(wf_engine *, double *vmax);
int wf_cfl_dt(wf_engine *, double cfl_factor, double *dt);
int wf_set_dt(wf_engine *, double dt);
/* step monitor that does not drain the stream: enqueue {kinetic energy, non-finite flag} -> pinned host memory;
 * wf_monitor_wait returns the oldest pending result (at most two pending) */
int wf_monitor_async(wf_engine *);
int wf_monitor_wait(wf_engine *, double *Ekin, int *nonfinite);
/* profiling / tuning hooks (no reference counterpart): wf_step with a CUDA event after every launch,
 * ms[0..4] += device time of predictor, E1 (element volume), N1 (nodal sums), E2 (main element pass),
 * N2 (assembly + integration; partitioned mesh: the nodes this rank does not share); partitioned mesh, peer transport:
 * ms[5] += nodal sums of the shared nodes, ms[6] += N2 of the shared nodes (both start by waiting for the neighbours, so
 * they show a neighbour's lateness), ms[7] += halo kernels of their own (WF_HALO_FOLD=0).  ms has 8 entries.
 * wf_set_variant: selection of an alternative implementation of one of the four kernels */
int wf_step_timed(wf_engine *, int nsteps, float *ms8);
int wf_set_variant(wf_engine *, int kernel /*0..3 = E1,N1,E2,N2*/, int variant);

/* 1:1 unfused entry points for parity bisecting; names = Domain_d members */
int wf_UpdatePrediction(wf_engine *);          /* Domain_d.C:961 */
int wf_ImposeBCV(wf_engine *, int d);          /* Domain_d.C:1109 */
int wf_ImposeBCA(wf_engine *, int d);          /* Domain_d.C:1123 */
int wf_calcElemJAndDerivatives(wf_engine *);   /* Domain_d.C:1701 */
int wf_Calc_Element_Radius(wf_engine *);       /* Domain_d.C:2140 */
int wf_CalcElemVol(wf_engine *);               /* Mechanical.C:264 */
int wf_CalcNodalVol(wf_engine *);              /* Mechanical.C:1555 */
int wf_CalcNodalMassFromVol(wf_engine *);      /* Mechanical.C:1576 */
int wf_calcElemStrainRates(wf_engine *);       /* Mechanical.C:41 */
int wf_calcElemPressure(wf_engine *);          /* Solver_explicit.C:735-746 dispatch */
int wf_CalcStressStrain(wf_engine *, double dt); /* Mechanical.C:1664 */
int wf_calcArtificialViscosity(wf_engine *);   /* Mechanical.C:1948 */
int wf_calcElemForces(wf_engine *);            /* Mechanical.C:375 */
int wf_calcElemHourglassForces(wf_engine *);   /* Mechanical.C:1842 (+ f90_ver hexa form) */
int wf_assemblyForces(wf_engine *);            /* Matrices.C:42 (+ non-finite scrub, Solver_explicit.C:779) */
int wf_calcAccel(wf_engine *);                 /* Mechanical.C:321 */
int wf_UpdateCorrectionAccVel(wf_engine *);    /* Domain_d.C:981 */
int wf_AxisConstraint(wf_engine *);            /* Solver_explicit.C:953-969 */
int wf_UpdateCorrectionPos(wf_engine *);       /* Domain_d.C:1005 */

/* ---- state access (checkpoint / parity dump); names = Domain_d member names --------------------- */
int wf_get_array(wf_engine *, const char *name, void *host_dst, size_t bytes);
int wf_set_array(wf_engine *, const char *name, const void *host_src, size_t bytes);
size_t wf_array_bytes(wf_engine *, const char *name); /* 0 if unknown */
/* raw device pointer of a private SoA array (for zero-copy interop, e.g. torch tensors); NULL if unknown */
void *wf_device_ptr(wf_engine *, const char *name, size_t *pitch_elems);

/* ---- multi-GPU: one engine per rank/GPU (one process per GPU) ------------------------------------ */
/* Canonical partition (no reference exists): elements sorted by id, rank p owns the
 * contiguous block [floor(p*Ne/P), floor((p+1)*Ne/P)); a node is local to every rank owning an element
 * that references it; local node order = ascending global id; halo list per neighbour = shared global
 * ids ascending.  Host-side, deterministic, no GPU needed. */
typedef struct wf_partition wf_partition;
int wf_partition_build(wf_partition **out, int nranks, int rank, int nodxelem, int n_nodes, int n_elems,
                       const unsigned *elnod);
/* same, for the AddBoxLength mesh without materialising the global connectivity */
int wf_partition_build_box(wf_partition **out, int nranks, int rank, const double V[3], const double L[3],
                           double r, int tritet);
void wf_partition_free(wf_partition *);
int wf_partition_info(const wf_partition *, int *elem_begin, int *elem_end, int *n_local_nodes, int *n_neigh);
const int *wf_partition_node_l2g(const wf_partition *);            /* [n_local_nodes] */
const unsigned *wf_partition_local_elnod(const wf_partition *);    /* [(elem_end-elem_begin)*k] */
const int *wf_partition_neigh_ranks(const wf_partition *);         /* [n_neigh] ascending */
const int *wf_partition_halo_offset(const wf_partition *);         /* [n_neigh+1] */
const int *wf_partition_halo_nodes(const wf_partition *);          /* local node ids, grouped by neighbour */
/* load the local part into an engine (replaces wf_set_mesh / wf_gen_box on that rank).  From here on the node
 * ids given to wf_add_bc_vel are GLOBAL ids (nodes of other ranks are skipped) and wf_get_array / wf_set_array
 * move LOCAL arrays (local node order = ascending global id, wf_halo_info -> node_l2g). */
int wf_set_mesh_partition(wf_engine *, const wf_partition *, const double *x_local /*n_local*dim or NULL for box*/);
/* Axisymmetric domains: the axis constraint (Solver_explicit.C:953-969) needs the minimum radial coordinate of the WHOLE
 * mesh.  A partitioned axisymmetric engine keeps the rank-local minimum, which is the global one as long as the rank
 * owns a node on the axis (those nodes never move radially) — true for every rank of an AddBoxLength box cut into
 * contiguous element blocks.  Call this with the global minimum BEFORE wf_set_mesh_partition; the partition is refused
 * when the rank has no node within 1e-6 of it (or when this was not called). */
int wf_set_axis_xmin(wf_engine *, double global_xmin);
int wf_halo_info(wf_engine *, int *rank, int *nranks, int *n_neigh, const int **neigh_ranks, const int **halo_offset,
                 const int **node_l2g);

/* Halo exchange, default transport = peer memory over NVLink: every engine owns one "comm block"
 * [flags | receive regions]; a neighbour's send kernel stores its partial nodal sums straight into the region
 * reserved for it and then publishes the exchange's sequence number in its flag slot; the consumer side waits on
 * the flags with a one-CTA kernel.  After the neighbours are connected, wf_init / wf_step of a partitioned engine
 * enqueue the complete distributed step (E1, N1, send, wait, finish, E2, send, wait, N2) with no host involvement.
 * Two exchanges per step: 2 doubles per shared node after N1 (sum of element volumes), dim doubles after E2
 * (internal force), summed in ascending rank order on every sharer so all copies of a shared node stay bit-identical. */
int wf_halo_comm_block(wf_engine *, void **dev_base, size_t *bytes);
/* where neighbour index i of THIS engine must write: offsets inside this engine's comm block */
int wf_halo_slot_offsets(wf_engine *, int neigh_idx, size_t *flag_byte_offset, size_t *region_byte_offset, size_t *region_bytes);
/* one process per GPU: export the comm block (cudaIpcGetMemHandle, 64 bytes) / map a neighbour's block */
int wf_halo_ipc_export(wf_engine *, void *handle64);
int wf_halo_ipc_open(wf_engine *, const void *handle64, void **mapped_base);
/* tell this engine where neighbour neigh_idx receives: the neighbour's comm block as mapped in this process and
 * the offsets the NEIGHBOUR reports from wf_halo_slot_offsets(neighbour, index of this rank in its list) */
int wf_halo_connect(wf_engine *, int neigh_idx, void *peer_comm_base, size_t peer_flag_offset, size_t peer_region_offset);
int wf_halo_status(wf_engine *, int *error); /* non-zero error: a wait timed out (WF_HALO_TIMEOUT_S, default 30 s) */
/* all ranks inside one process (one host thread drives every GPU; also how two ranks are tested on one GPU) */
int wf_connect_all(wf_engine **ranks, int nranks);
int wf_init_all(wf_engine **ranks, int nranks, double dt);
int wf_step_all(wf_engine **ranks, int nranks, int nsteps);

/* Alternative transport, host-driven (NCCL send/recv issued by the caller between phases): sends are packed into
 * a local staging block; after each of phases 0 and 1 the caller moves, for every neighbour i, n_doubles from
 * send_ptr to the neighbour's recv_ptr (wf_halo_exchange_ptrs on both sides describe the same exchange).
 *   wf_init_phase 0 | exchange | 1 | exchange | 2        wf_step_phase 0 | exchange | 1 | exchange | 2 */
int wf_halo_set_transport(wf_engine *, int host_driven);
int wf_halo_exchange_ptrs(wf_engine *, int neigh_idx, void **send_ptr, void **recv_ptr, size_t *n_doubles);
int wf_init_phase(wf_engine *, int phase, double dt);
int wf_step_phase(wf_engine *, int phase, int last_step);

/* ---- host-side helpers (no GPU needed; used by tests against the oracle) ------------------------- */
int wf_host_box_counts(const double L[3], double r, int tritet, int *dim, int *nodxelem, int *n_nodes, int *n_elems);
int wf_host_gen_box(const double V[3], const double L[3], double r, int tritet, double *x, unsigned *elnod);
int wf_host_nodel(int n_nodes, int n_elems, int nodxelem, const unsigned *elnod, int *nodel_offset,
                  int *nodel_count, int *nodel /*N_e*k*/, int *nodel_loc /*N_e*k*/);
/* SearchExtNodes, integer part: ext_nodes[n_nodes] flags, m_faceCount, and the external faces in faceList order
 * (ext_face_nodes: 3 (tets) / 2 (quads) ids per face, capacity 4*n_elems faces; ext_face_elem: owning element) */
int wf_host_ext_faces(int dim, int nodxelem, int n_nodes, int n_elems, const unsigned *elnod, unsigned char *ext_nodes,
                      int *n_faces_total, int *n_ext_faces, int *ext_face_nodes, int *ext_face_elem);
/* TriMesh_d::AxisPlaneMesh (Mesh.C:48-283): one rigid plane (3D: 2*dens^2 triangles) or line (2D: dens segments) */
int wf_host_axis_plane_counts(int dimension, int dens, int *n_nodes, int *n_elems);
int wf_host_axis_plane_mesh(int dimension, int mesh_id, int axis, int positaxisorent, const double p1[3], const double p2[3],
                            int dens, double *node, int *elnode, double *normal, int *ele_mesh_id);
/* Tables of the tile-reduced force path (tile = 32 consecutive elements): call
 * with NULL buffers for info[7] = {usable, n_tiles, stride, tpitch, n_slices, n_slot_entries, rounds}, then with buffers
 * tidx[n_elems*k] (position of (e, ln) in its tile's ascending unique-node list), ptr[n_slices+1] / slots[n_slot_entries]
 * (sliced-ELL of the tile entries of every node: tile*dim*stride + position, ascending tile, ~0u padding) and, for
 * every element type but the hexahedron, tab[n_tiles*tpitch] (per-tile incidence: ptr[stride+1] then inc[32k], uint8).  Integer artefacts: the
 * tests compare them bit for bit with a numpy restatement. */
int wf_host_force_tiles(int n_nodes, int n_elems, int nodxelem, int dim, const unsigned *elnod, long long *info,
                        unsigned char *tidx, long long *ptr, unsigned *slots, unsigned char *tab);
/* Internal element order of the engine: perm[internal] = user element id.  mode 0 = identity,
 * 1 = Morton order of the quantised element centroids (ascending (key, user id)).  Integer artefact, bit-exact
 * against the numpy restatement in tests/test_host_mesh.py. */
int wf_host_elem_order(int dim, int nodxelem, int n_nodes, int n_elems, const double *x, const unsigned *elnod, int mode,
                       int *perm);
/* Same, also returning the sort key of every element in the resulting internal order (mode 1; keys may be NULL). */
int wf_host_elem_order_keys(int dim, int nodxelem, int n_nodes, int n_elems, const double *x, const unsigned *elnod, int mode,
                            int *perm, unsigned long long *keys);
/* Thread slots of the brick passes of the hexahedron path (csrc/wf_mesh.cpp): with the keys of wf_host_elem_order_keys
 * strictly ascending (one element per cell), CTA = rank of key >> 7 (an 8x4x4 group of cells), thread = key & 127:
 * slot_elem[cta * 128 + thread] = internal element, -1 = no such cell.  slot_elem = NULL returns n_cta only.  Returns 1
 * when two elements share a key (the engine then keeps the compact numbering 128 cta + thread). */
int wf_host_brick_plan(int n_elems, const unsigned long long *keys, int *n_cta, int *slot_elem);
/* Shared-memory slots of a sorted node list for the brick form of the hexa main pass (csrc/wf_mesh.cpp): run r of
 * consecutive ids starts at the first free slot congruent to 12 r (mod 16).  Returns the slot count. */
int wf_host_run_slots(int n, const int *sorted_ids, int *slots);
const char *wf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WF_ENGINE_H */
