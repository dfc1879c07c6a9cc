// wf_dev.h — private device-side view of one engine (passed by value to every kernel).
//
// HBM layout (private; the C ABI converts to/from the reference layouts):
//   * every array is SoA ("component-major") with a pitch that is a multiple of 32
//     elements, so lane i of a warp touches element i of a 256 B-aligned run:
//       nodal vectors     q[c * np + n]           c < dim      (reference: q[dim*n + c])
//       element 6-vectors t[c * ep + e]           c < 6        (reference: t[6*e + c])
//       element-node recs f[(ln*dim + c) * ep + e]             (reference: f[(e*k + ln)*dim + c])
//       connectivity      elnod[ln * ep + e]                   (reference: m_elnod[e*k + ln])
//   * node->element connectivity (m_nodel / m_nodel_loc, Domain_d.h:1004-1007) is kept on the
//     device as ONE packed array of "slots" slot = e*k + ln in sliced-ELL form: slice s = nodes
//     [32s, 32s+32), width = max list length in the slice, entry j of node n at
//     sell_ptr[s] + 32*j + (n & 31), -1 = padding.  List order == reference order (ascending
//     element id), so gathers sum in exactly the reference's order and a warp reads 128 B rows.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#define WF_MAXK 8
#define WF_EBLK 128 /* elements per CTA of the element passes */
#define WF_BRICK_STRIDE 304 /* compile-time pitches of k_elem_main_hex_brick: slots of a CTA's node copy (8x4x4 brick of a */
#define WF_BRICK_WS 176     /* structured mesh: 297) and of a tile's accumulators (4x4x2 brick: 173) */
#define WF_HALO_NC 3 /* doubles per shared node and exchange (max: dim force components / init triple) */

struct WfHaloNb {
  int offset, count;            /* slice of halo_nodes */
  double *dst;                  /* receive region for this rank in the neighbour's memory (peer pointer), or the local staging region */
  unsigned long long *flag;     /* this rank's flag slot in the neighbour's memory */
  unsigned *counter;            /* local block-completion counter */
};

struct WfDev {
  int nn, ne, nslices;
  int dim, k, domtype, vol_weight;
  long long np, ep; /* pitches */

  /* nodes */
  double *x, *v, *prev_a, *u, *u_dt; /* [dim][np] persistent state */
  double *fe;                        /* [dim][np] external force, NULL == 0 (m_fe) */
  double *voln_sum;                  /* [np] sum of vol over nodel(n) (no /k) */
  double *voln0_sum;                 /* [np] sum of vol_0 (press 0/1) or of vol_0/4.0 (press 3) */
  double *nodal_p;                   /* [np] press 0: voln/voln_0 ; press 1/3: nodal pressure pn */
  double *rhobar;                    /* [np] mean rho of the elements around the node (WF_FAST mass) */
  int *nodel_count;                  /* [np] */
  int *bc_index;                     /* [np] -1 or row of bc_mask / bc_vals */
  const unsigned char *bc_mask;      /* [nbc] bit c: dim c prescribed */
  const double *bc_vals;             /* [nbc*3] */
  const long long *sell_ptr;         /* [nslices+1] */
  const int *sell_slots;
  /* halo (multi-GPU, NULL / 0 on a single GPU): shared nodes whose partial nodal sums are exchanged.
   * Send side: halo_nodes grouped by neighbour, described by nb[]; every neighbour owns one receive
   * region [parity][WF_HALO_NC][count] in the peer's memory, written with plain stores over NVLink.
   * Receive side: one record per UNIQUE shared node, listing its sharers in ascending rank order. */
  const int *halo_nodes;             /* [n_halo] local node ids, grouped by neighbour */
  const struct WfHaloNb *nb;         /* [n_neigh] */
  const int *halo_slot;              /* [np] index into hu_* or -1 */
  const int *hu_node;                /* [n_uniq] local node id */
  const int *hu_ptr;                 /* [n_uniq+1] */
  const int2 *hu_ent;                /* .x = index of (parity 0, comp 0) in recv, or -1 = this rank's own partial; .y = count of that neighbour */
  const double *recv;                /* receive regions of this rank */
  unsigned long long *flags;         /* [n_neigh] sequence number of the last exchange completed by that neighbour */
  int *comm_error;                   /* [1] set when a wait timed out */
  int n_halo, n_uniq, n_neigh;

  /* elements (arrays in the engine's INTERNAL element order, see wf_host_elem_order) */
  const int *elnod;                  /* [k][ep] */
  const int *e_user;                 /* [ep] user id of internal element e, NULL = identity (only the reference quirks
                                      * that index a NODAL array with an element id read it) */
  double *tau;                       /* [6][ep] */
  double *sigma;                     /* [6][ep] (optional per step) */
  double *eps;                       /* [6][ep] (optional) */
  double *p, *pl_strain, *sigma_y, *vol, *vol_0, *rho, *rho_0; /* [ep] */
  double *hg_q;                      /* [2][ep] 2D quads (m_hg_q) */
  /* element-block node tables: CTA b of the element passes owns elements [b*WF_EBLK, (b+1)*WF_EBLK); its
   * UNIQUE nodes (ascending id) are blk_nodes[blk_off[b] .. blk_off[b+1]) and lidx gives every element-node its
   * index in that list, so the CTA stages each node's data in shared memory once instead of once per element. */
  const int *blk_off;                /* [nblk+1] */
  const int *blk_nodes;
  const unsigned short *lidx;        /* [k][ep] */
  int blk_umax;                      /* longest unique list */
  int blk_pitch;                     /* pitch of blk_pad and of the staged copies in shared memory: blk_umax rounded up to 32 */
  const int *blk_pad;                /* [nblk][blk_pitch] the same lists at a fixed pitch, -1 padded (no blk_off round trip) */
  const int *pos;                    /* [k][ep] offset of (e, ln) in fsell: dim*q - (dim-1)*(n&31), q = sell index */
  double *fsell, *fsell_hg;          /* [slice][j][dim][32] node-ordered element (and hourglass) forces */
  /* tile-reduced forces (hexa, WF_FAST only; NULL when the mesh does not qualify): instead of one force record per
   * element node, every WARP of the main element pass (32 thread slots = one "force tile") sums the
   * contributions of its elements per UNIQUE node in shared memory — one round per local corner, each round a
   * conflict-free read-modify-write, so the order of additions is fixed — and writes one partial per (tile, unique
   * node): ftile[(tile*3 + c) * tf_stride + u], u = tf_idx[corner][e] = index of the node in the tile's ascending
   * unique list.  Node n then adds its tile partials in ascending tile order: entry j of node n is
   * tf_slots[tf_ptr[n >> 5] + 32*j + (n & 31)] = offset of component 0, ~0u = padding. */
  double *ftile;
  const long long *tf_ptr;           /* [nslices+1] */
  const unsigned *tf_slots;
  const unsigned char *tf_idx;       /* [k][ep] (hexahedra) */
  /* brick form of the hexa passes (NULL when the bank-aware layouts do not fit the compile-time pitches).  The brick
   * kernels run n_bcta CTAs of 128 THREAD SLOTS; slot s = cta * 128 + thread works on element brick_elem[s] (>= 0), or
   * idles (< 0: ~brick_elem[s] is an element of the same CTA whose tables the idle thread reads).  brick_plan = 1: the
   * slots follow the cells of the mesh (wf_host_brick_plan: every CTA a clipped 8x4x4 brick), 0: slot s = element s.
   * Node list of CTA b at blk_pad_b + b * WF_BRICK_STRIDE indexed by shared-memory SLOT (wf_host_run_slots, -1 = hole);
   * lidx_pk[s] = the slots of the element's eight nodes in the CTA copy (16 bit each); tile_pk[s] = {x, y: its eight
   * 8-bit slots in the tile's accumulators; z: byte j = accumulator slot of the tile's unique node of rank
   * lane + 32 j (0xff = none: at most 128 unique nodes), read when the partial sums are written out in rank order;
   * w = brick_elem[s]} */
  const int *blk_pad_b;
  const uint4 *lidx_pk;              /* [n_bcta * 128] */
  const uint4 *tile_pk;              /* [n_bcta * 128] */
  const int *brick_elem;             /* [n_bcta * 128] */
  int n_bcta, brick_plan;
  int cta_lookahead;                 /* resident CTAs of the main element pass on the device (L2 look-ahead distance) */
  int sm_count;                      /* multiprocessors of the device (look-ahead distance of the generic element pass) */
  /* tetrahedra: corners of different elements of a tile DO share nodes, so the tile sum is pulled instead: every
   * element drops its k*dim force values in shared memory and lane u adds up the entries of unique node u listed in
   * the tile's incidence table (ascending element, then corner: a fixed order).  Table of tile w at
   * tf_tab + w * tf_tpitch: ptr[tf_stride + 1] then inc[32 * k], all uint8; inc = local element * k + corner. */
  const unsigned char *tf_tab;
  int tf_tpitch;
  int tf_stride;                     /* longest unique list of a force tile, rounded up to a multiple of 4 */
  double *f_elem;                    /* [k*dim][ep]  (unfused path only) */
  double *f_elem_hg;                 /* [k*dim][ep]  (unfused path only) */

  /* unfused-path intermediates (allocated on first use) */
  double *dH;                        /* [dim][k][ep] dN/dX * detJ (m_dH_detJ_dx/dy/dz) */
  double *detJ, *radius;             /* [ep] */
  double *str_rate, *rot_rate;       /* [6][ep] */
  double *pl_incr;                   /* [6][ep] m_strain_pl_incr */
  double *a, *fi;                    /* [dim][np] */
  double *mdiag, *voln, *p_node;     /* [np] */

  /* contact (NULL when contact is off): contact force per node (persists between steps like the reference's
   * contforce, Contact.C:211-213) and the "has a non-zero contact force" flag read by calcElemPressure */
  double *contforce;                 /* [dim][np] */
  unsigned char *cflag;              /* [np] */

  /* thermal coupling (NULL when off; Thermal.C): nodal temperature, node-ordered element contributions m_dTedt
   * (one double per (e, ln), same ordering as fsell), plastic heat source, and the first n_nodes entries of the
   * reference's flat m_dTedt[e*k+ln] of the previous / current step (calcThermalExpansion reads that array with
   * NODE ids, Thermal.C:157-160) */
  double *T;                         /* [np] */
  double *tsell;                     /* [sell_total] */
  double *q_plheat;                  /* [ep] m_q_plheat */
  double *dtedt_low[2];              /* [np] each */
  double *q_cont_conv;               /* [np] contact heat flow (Contact.C:309), persists like contforce */

  /* flags / reductions */
  int *nonfinite;                    /* [1] */
  unsigned long long *xmin_key;      /* [2] ordered-key of min x_r (axisymmetric axis constraint) */
  double *red;                       /* [8] energy reductions */
  double *ekin_acc;                  /* NULL, or where this step's node pass adds 1/2 m |v|^2 of the corrected velocities (step monitor) */
};

struct WfPar {
  /* material (Material.cuh) */
  int model;
  double Kbulk, G, sy0, Kh, mh, eps0, eps1, cs0;
  int thermal, dtedt_cur;               /* thermal coupling on; which dtedt_low buffer this step writes */
  double k_T, cp_T, exp_T, plheatfrac;  /* Material_::k_T / cp_T / exp_T (Material.cuh:79-81), m_plheatfraction */
  double young, mq[14], temp, max_edot; /* Johnson-Cook / GMT constants (wf_material::q), uniform temperature */
  /* StabilizationParams + hexa hourglass coefficient */
  double alpha_contact, hg_coeff_contact; /* used instead of the _free values by elements touching a contact node */
  double alpha_free, hg_coeff_free, av_coeff_div, av_coeff_bulk, log_factor, pspg_scale, p_pspg_bulkfac, J_min;
  double hg_visc, hg_stiff, hexa_hg;
  int stab_simple; /* all pressure-stabilisation terms are zero -> p = -K (J_avg - 1) */
  int press;
  double av_alpha, av_beta;
  int track_eps, store_sigma, strict;
  /* time integration (Solver_explicit.C:193-197) */
  double dt, alpha, beta, gamma;
  double w; /* Gauss weight, Mechanical.C:269-282 */
  int xmin_cur; /* which xmin_key slot holds min x_r of the current coordinates */
  int halo_parity; /* multi-GPU: which half of the receive regions the current exchange uses */
  /* multi-GPU, peer transport: halo work folded into the node passes.  send_ctas > 0: the first send_ctas CTAs of the
   * launch send this rank's partial sums to the neighbours (send_chunks CTAs per neighbour, exchange number send_seq)
   * instead of a separate k_halo_send launch; wait_seq > 0: every CTA of the consumer first waits until all neighbours
   * have published exchange wait_seq (instead of a separate k_halo_wait launch). */
  int send_ctas, send_chunks;
  unsigned long long send_seq, wait_seq, wait_timeout_ns;
  int variant[4]; /* tuning: kernel variant for E1, N1, E2, N2 (0 = default) */
};

/* Contact with rigid surfaces (Contact.C, Mesh.h): everything the contact kernels need besides WfDev. */
struct WfContact {
  /* external nodes (SearchExtNodes, Domain_d.C:110-205), ascending node id */
  int n_ext;
  const int *ext_nodes;              /* [n_ext] */
  double *nodlen;                    /* [n_ext] mean m_elem_length of the elements around the node (Contact.C:146-153) */
  double *node_area;                 /* [np] CalcExtFaceAreas */
  double *ut_prev;                   /* [dim][np + 32] accumulated tangential slip */
  int *mesh_in_contact;              /* [np] m_mesh_in_contact */
  double *rec;                       /* [10][n_ext] 2D only: per-node records handed to the serial friction pass */
  int *cand_count;                   /* [1] number of candidate nodes of this step's contact search */
  int2 *cand;                        /* [n_ext] (external-node index, first facet not ruled out by the fp32 filter) */
  /* external faces in faceList order (= ascending element, then local face) */
  int n_xf, facenod;
  const int *xf_nodes;               /* [facenod][n_xf] */
  const int *xf_elem;                /* [n_xf] */
  double *xf_area;                   /* [n_xf] face area (3D) / edge length (2D) */
  const int *xn_ptr;                 /* [n_ext+1] faces of external node t: xn_faces[xn_ptr[t] .. xn_ptr[t+1]) ascending */
  const int *xn_faces;
  double *elem_area;                 /* [ep] m_elem_area */
  /* rigid surfaces: TriMesh_d, arrays of double3 as xyzxyz */
  int tm_dim, tm_nn, tm_ne;
  double *tm_node, *tm_node_v, *tm_normal, *tm_pplane;
  const double *tm_v_orig;
  const int *tm_elnode, *tm_mesh_id;
  double mu_sta, mu_dyn, contPF, young;
  double heat_cond, T_const;         /* TriMesh_d::heat_cond / T_const (main.C:718-719) */
};
