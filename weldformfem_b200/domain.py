"""Host-side mirror of the reference's ``MetFEM::Domain_d`` for the explicit step, over the C ABI.

Method names and argument meaning follow the reference (include/common/Domain_d.h): ``AddBoxLength``,
``AddBCVelNode``, ``AllocateBCs``, ``SetDT`` ..., plus the snake_case aliases the CPU checkers under
``oracle/`` use so that a parity test can drive engine and checker with the same calls.  All work is
done by hand-written CUDA kernels in ``libwf_b200.so``; there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import STAB_FIELDS, wf_material, wf_stab

PLANE_STRAIN, AXISYMM, DOM_3D = 0, 2, 3
BILINEAR, HOLLOMON, JOHNSON_COOK, GMT = 0, 1, 2, 3
STRICT, FAST = 1, 0

_INT_ARRAYS = {"m_nodel": np.int32, "m_nodel_loc": np.int32, "m_nodel_offset": np.int32, "m_nodel_count": np.int32,
               "m_elnod": np.uint32, "ext_nodes": np.uint8, "m_mesh_in_contact": np.int32, "elem_perm": np.int32}


def axis_plane_mesh(dimension, mesh_id, axis, positaxisorent, p1, p2, dens):
    """TriMesh_d::AxisPlaneMesh (src/common/Mesh.C:48-283) on the host: (node[n,3], elnode[e,nen], normal[e,3], mesh_id[e])."""
    lib = _lib.load()
    nn, ne = C.c_int(), C.c_int()
    if lib.wf_host_axis_plane_counts(int(dimension), int(dens), C.byref(nn), C.byref(ne)):
        raise WfError("bad plane mesh parameters")
    nen = 3 if dimension == 3 else 2
    node = np.zeros((nn.value, 3)); elnode = np.zeros((ne.value, nen), dtype=np.int32)
    normal = np.zeros((ne.value, 3)); mid = np.zeros(ne.value, dtype=np.int32)
    a3 = lambda q: (C.c_double * 3)(*[float(t) for t in q])
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    rc = lib.wf_host_axis_plane_mesh(int(dimension), int(mesh_id), int(axis), int(bool(positaxisorent)), a3(p1), a3(p2),
                                     int(dens), node.ctypes.data_as(dp), elnode.ctypes.data_as(ip),
                                     normal.ctypes.data_as(dp), mid.ctypes.data_as(ip))
    if rc:
        raise WfError("wf_host_axis_plane_mesh failed")
    return node, elnode, normal, mid


class WfError(RuntimeError):
    pass


class Domain_d:
    """One explicit-dynamics domain on one GPU."""

    def __init__(self, device: int = 0, strict: bool = False, elem_order: int | None = None):
        """elem_order: internal element order of the engine (None = default = Morton, 0 = the caller's numbering);
        arrays always cross the ABI in the caller's numbering."""
        self._elem_order = elem_order
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._device = device
        self._strict = bool(strict)
        self._domtype = DOM_3D
        self._vol_weight = False
        self._pending_bcs = []
        self._dt = None
        self.dim = None
        self.nodxelem = None
        self._tracking = 0

    # ---- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.wf_last_error(self._h if self._h else None)
            raise WfError(msg.decode() if msg else f"error {rc}")

    def _create(self, dim, k):
        if self._h:
            raise WfError("mesh already created")
        domtype = DOM_3D if dim == 3 else self._domtype
        if dim == 2 and domtype == DOM_3D:
            domtype = PLANE_STRAIN
        h = C.c_void_p()
        rc = self._lib.wf_create(C.byref(h), dim, k, domtype, self._device)
        if rc != 0:
            raise WfError(self._lib.wf_last_error(None).decode())
        self._h = h
        if self._elem_order is not None:
            self._ck(self._lib.wf_set_elem_order(self._h, int(self._elem_order)))
        self.dim, self.nodxelem, self._domtype = dim, k, domtype
        if domtype == AXISYMM and self._vol_weight:
            self._ck(self._lib.wf_set_axisymm_vol_weight(self._h, 1))

    def close(self):
        if self._h:
            self._lib.wf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.wf_set_stream(self._h, C.c_void_p(cuda_stream)))

    def get_stream(self) -> int:
        s = C.c_void_p()
        self._ck(self._lib.wf_get_stream(self._h, C.byref(s)))
        return s.value or 0

    def synchronize(self):
        self._ck(self._lib.wf_synchronize(self._h))

    # ---- setup: names of the reference ------------------------------------------------------------
    def setAxiSymm(self, vol_weight: bool = False):           # Domain_d.h:666
        self._domtype, self._vol_weight = AXISYMM, bool(vol_weight)

    def set_domtype(self, domtype: int, vol_weight: bool = False):
        self._domtype, self._vol_weight = int(domtype), bool(vol_weight)

    def AddBoxLength(self, V, L, r, red_int: bool = True, tritetra: bool = False):   # Domain_d.C:1136
        if not red_int:
            raise WfError("full integration is not implemented by the reference step (Domain_d.C:1906-2008)")
        dim = 3 if L[2] > 0.0 else 2
        k = (4 if tritetra else 8) if dim == 3 else (3 if tritetra else 4)
        self._create(dim, k)
        Vd, Ld = (C.c_double * 3)(*V), (C.c_double * 3)(*L)
        self._ck(self._lib.wf_gen_box(self._h, Vd, Ld, float(r), int(tritetra)))

    def box(self, V, L, r, tritet=False):
        self.AddBoxLength(V, L, r, True, bool(tritet))

    def set_mesh(self, dim, k, x, elnod):                     # CreateFromLSDyna path, Domain_d.C:1647
        x = np.ascontiguousarray(x, dtype=np.float64)
        el = np.ascontiguousarray(elnod, dtype=np.uint32)
        self._create(dim, k)
        self._ck(self._lib.wf_set_mesh(self._h, x.size // dim, el.size // k, x.ctypes.data_as(C.POINTER(C.c_double)),
                                       el.ctypes.data_as(C.POINTER(C.c_uint))))

    def set_material(self, E, nu, rho0, model=BILINEAR, sy0=1.0e10, K=0.0, m=1.0):   # main.C:460-581
        mat = wf_material(int(model), float(E), float(nu), float(rho0), float(sy0), float(K), float(m))
        self._ck(self._lib.wf_set_material(self._h, C.byref(mat)))

    def set_material_ext(self, E, nu, rho0, model, sy0, params, temp=20.0, max_edot=None):
        """Johnson-Cook (params = A B n C eps_0 m T_m T_t) / GMT (n1 n2 C1 C2 m1 m2 I1 I2 e_min e_max er_min er_max T_min
        T_max): the public Material_ fields read by CalcJohnsonCook* / CalcGMT* (Material.cuh:377-483); ``temp`` is the
        uniform temperature the flow stress sees with thermal coupling off; ``max_edot`` = m_max_edot (Domain_d.h:824)."""
        mat = wf_material(int(model), float(E), float(nu), float(rho0), float(sy0), 0.0, 1.0)
        for i, v in enumerate(params):
            mat.q[i] = float(v)
        mat.temp = float(temp)
        mat.max_edot = float(max_edot) if max_edot is not None else 0.0
        self._ck(self._lib.wf_set_material(self._h, C.byref(mat)))

    def thermal_on(self, k_T, cp_T, exp_T=0.0, plheatfrac=0.9, T0=20.0):
        """setThermalOn + setTemp(T0) + thermalCond / thermalHeatCap / thermalExp + plHeatFrac (main.C:218, 436-441, 567-570)."""
        self._ck(self._lib.wf_set_thermal(self._h, float(k_T), float(cp_T), float(exp_T), float(plheatfrac), float(T0)))

    def set_contact_heat(self, heat_cond, T_const):            # heatCondCoeff / dieTemp, main.C:718-719
        self._ck(self._lib.wf_set_contact_heat(self._h, float(heat_cond), float(T_const)))

    def set_stab(self, **kw):                                  # m_stab, main.C:84-120
        self._stab_kw = {k: float(v) for k, v in kw.items()}
        self._push_stab()

    def _push_stab(self, hexa_hg=None):
        kw = dict(getattr(self, "_stab_kw", {}))
        if hexa_hg is not None:
            kw["hexa_hg_coeff"] = hexa_hg
            self._stab_kw = kw
        st = wf_stab(*[float(kw.get(f, 0.0)) for f in STAB_FIELDS])
        self._ck(self._lib.wf_set_stab(self._h, C.byref(st)))

    def set_options(self, press=0, av_alpha=0.0, av_beta=0.0, hexa_hg=0.0, strict=None):
        if strict is not None:
            self._strict = bool(strict)
        self._push_stab(hexa_hg=float(hexa_hg))
        self._ck(self._lib.wf_set_options(self._h, int(press), float(av_alpha), float(av_beta), int(self._strict)))

    def set_tracking(self, eps: bool = False, sigma: bool = False):
        self._tracking = (1 if eps else 0) | (2 if sigma else 0)
        self._ck(self._lib.wf_set_tracking(self._h, self._tracking))

    def AddBCVelNode(self, node, dim, val):                    # Domain_d.C:1057
        self._ck(self._lib.wf_add_bc_vel(self._h, int(node), int(dim), float(val)))

    add_bc = AddBCVelNode

    def add_bcs(self, nodes, dims, vals):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self._ck(self._lib.wf_add_bc_vel_array(self._h, nodes.size, nodes.ctypes.data_as(C.POINTER(C.c_int)),
                                               dims.ctypes.data_as(C.POINTER(C.c_int)),
                                               vals.ctypes.data_as(C.POINTER(C.c_double))))

    def AllocateBCs(self):                                     # Domain_d.C:1063
        self._ck(self._lib.wf_allocate_bcs(self._h))

    allocate_bcs = AllocateBCs

    def set_bc_values(self, dim, vals):
        """New values for bcx_val / bcy_val / bcz_val of one dimension (insertion order), uploaded asynchronously."""
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self._ck(self._lib.wf_set_bc_values(self._h, int(dim), vals.size, vals.ctypes.data_as(C.POINTER(C.c_double))))

    # ---- contact with rigid tool surfaces (src/explicit/main.C:636-848) ---------------------------
    def SearchExtNodes(self):                                  # Domain_d.C:110
        self._ck(self._lib.wf_SearchExtNodes(self._h))

    def add_plane(self, dimension, mesh_id, axis, positaxisorent, p1, p2, dens, vel=(0.0, 0.0, 0.0)):
        """One more rigid body: AxisPlaneMesh, every node moving with ``vel`` (main.C:672-708 for the first body,
        TriMesh_d::AddMesh, Mesh.C:438-539, for the others: node ids offset by the nodes already present)."""
        node, elnode, normal, mid = axis_plane_mesh(dimension, mesh_id, axis, positaxisorent, p1, p2, dens)
        tm = getattr(self, "_trimesh", None)
        if tm is None:
            tm = self._trimesh = dict(dimension=int(dimension), node=[], node_v=[], elnode=[], normal=[], mesh_id=[], nn=0)
        if tm["dimension"] != int(dimension):
            raise WfError("all rigid bodies must have the same dimension")
        tm["node"].append(node)
        tm["node_v"].append(np.tile(np.asarray(vel, dtype=np.float64), (node.shape[0], 1)))
        tm["elnode"].append(elnode + tm["nn"])
        tm["normal"].append(normal)
        tm["mesh_id"].append(mid)
        tm["nn"] += node.shape[0]

    def set_trimesh(self, dimension, node, node_v, elnode, normal, mesh_id):   # Domain_d::setTriMesh, Domain_d.h:464
        node = np.ascontiguousarray(node, dtype=np.float64)
        node_v = np.ascontiguousarray(node_v, dtype=np.float64)
        elnode = np.ascontiguousarray(elnode, dtype=np.int32)
        normal = np.ascontiguousarray(normal, dtype=np.float64)
        mesh_id = np.ascontiguousarray(mesh_id, dtype=np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        self._ck(self._lib.wf_set_trimesh(self._h, int(dimension), node.size // 3, mesh_id.size, node.ctypes.data_as(dp),
                                          node_v.ctypes.data_as(dp), elnode.ctypes.data_as(ip), normal.ctypes.data_as(dp),
                                          mesh_id.ctypes.data_as(ip)))
        self._trimesh = None

    def contact_on(self, mu_sta=0.0, mu_dyn=0.0, penalty_factor=-1.0, end_time=1.0):
        """fricCoeffStatic / fricCoeffDynamic / penaltyFactor (main.C:716-725), CalcSpheres + setContactOn (:842-847),
        SetEndTime.  Rigid bodies collected by add_plane are handed to the engine here."""
        tm = getattr(self, "_trimesh", None)
        if tm:
            self.set_trimesh(tm["dimension"], np.concatenate(tm["node"]), np.concatenate(tm["node_v"]),
                             np.concatenate(tm["elnode"]), np.concatenate(tm["normal"]), np.concatenate(tm["mesh_id"]))
        self._ck(self._lib.wf_set_contact(self._h, float(mu_sta), float(mu_dyn), float(penalty_factor), float(end_time)))

    def trimesh_counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.wf_get_trimesh_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(dimension=a.value, nodecount=b.value, elemcount=c.value)

    def SetDT(self, dt):                                       # Domain_d.h:629
        self._dt = float(dt)

    # ---- solve ------------------------------------------------------------------------------------
    def init(self, dt=None):
        """Initialisation part of SolveChungHulbert (Solver_explicit.C:115-292)."""
        if dt is not None:
            self._dt = float(dt)
        if self._dt is None:
            raise WfError("SetDT first")
        self._ck(self._lib.wf_init(self._h, self._dt))

    def step(self, n=1):
        """n fused time steps (rows 1-22 of the loop body, Solver_explicit.C:524-978)."""
        self._ck(self._lib.wf_step(self._h, int(n)))

    def step_open(self, n=1):
        """n steps, leaving the engine in predicted state (wf_step_open): for loops that call the engine once per step."""
        self._ck(self._lib.wf_step_open(self._h, int(n)))

    def step_close(self):
        self._ck(self._lib.wf_step_close(self._h))

    def step_timed(self, n=1):
        """wf_step with per-launch CUDA-event timing; returns ms for [predictor, E1, N1, E2, N2, shared-node sums,
        shared-node N2, halo kernels of their own] (the last three only on a partitioned mesh)."""
        ms = (C.c_float * 8)()
        self._ck(self._lib.wf_step_timed(self._h, int(n), ms))
        return list(ms)

    def set_variant(self, kernel: int, variant: int):
        self._ck(self._lib.wf_set_variant(self._h, int(kernel), int(variant)))

    def SolveChungHulbert(self, end_t):
        """Run the explicit loop up to end_t with the fixed step set by SetDT (while Time < end_t)."""
        self.init()
        t, n = 0.0, 0
        while t < end_t:
            t += self._dt
            n += 1
        self.step(n)
        return n

    def call(self, fn, arg=0.0):
        """Unfused entry points, names = Domain_d members (parity bisecting)."""
        if fn in ("ImposeBCV", "ImposeBCA"):
            self._ck(getattr(self._lib, "wf_" + fn)(self._h, int(arg)))
        elif fn in ("ImposeBCVAllDim", "ImposeBCAAllDim"):
            for d in range(self.dim):
                self._ck(getattr(self._lib, "wf_" + fn[:9])(self._h, d))
        elif fn == "CalcStressStrain":
            self._ck(self._lib.wf_CalcStressStrain(self._h, float(arg)))
        elif fn == "calcMinEdgeLength":
            self.calcMinEdgeLength()
        elif fn in _lib.UNFUSED:
            self._ck(getattr(self._lib, "wf_" + fn)(self._h))
        else:
            raise KeyError(fn)

    def nonfinite_flag(self) -> bool:
        f = C.c_int(0)
        self._ck(self._lib.wf_nonfinite_flag(self._h, C.byref(f)))
        return bool(f.value)

    def energies(self):
        ek, de = C.c_double(), C.c_double()
        self._ck(self._lib.wf_energies(self._h, C.byref(ek), C.byref(de)))
        return ek.value, de.value

    def calcMinEdgeLength(self):                               # Domain_d.C:2224
        a, b = C.c_double(), C.c_double()
        self._ck(self._lib.wf_calcMinEdgeLength(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def max_velocity(self):
        a = C.c_double()
        self._ck(self._lib.wf_max_velocity(self._h, C.byref(a)))
        return a.value

    def cfl_dt(self, cfl_factor):                              # Solver_explicit.C:579-598
        a = C.c_double()
        self._ck(self._lib.wf_cfl_dt(self._h, float(cfl_factor), C.byref(a)))
        return a.value

    def set_dt(self, dt):
        self._ck(self._lib.wf_set_dt(self._h, float(dt)))
        self._dt = float(dt)

    def monitor_async(self):
        self._ck(self._lib.wf_monitor_async(self._h))

    def monitor_wait(self):
        ek, f = C.c_double(), C.c_int()
        self._ck(self._lib.wf_monitor_wait(self._h, C.byref(ek), C.byref(f)))
        return ek.value, bool(f.value)

    def time(self):
        t, n = C.c_double(), C.c_long()
        self._ck(self._lib.wf_get_time(self._h, C.byref(t), C.byref(n)))
        return t.value, n.value

    def set_time(self, time, step_count):
        """Continue the clock of a previous engine (restart / remesh hand-off)."""
        self._ck(self._lib.wf_set_time(self._h, float(time), int(step_count)))

    # ---- state ------------------------------------------------------------------------------------
    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.wf_get_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def brick_info(self):
        """(CTAs of the brick form of the hexa passes, 1 if their thread slots follow the mesh cells) — wf_brick_info."""
        a, b = C.c_int(), C.c_int()
        self._ck(self._lib.wf_brick_info(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def info(self):
        nn, ne, _ = self.counts()
        return dict(dim=self.dim, nodxelem=self.nodxelem, n_nodes=nn, n_elems=ne, domtype=self._domtype)

    def get(self, name: str) -> np.ndarray:
        nbytes = self._lib.wf_array_bytes(self._h, name.encode())
        dt = _INT_ARRAYS.get(name, np.float64)
        if nbytes == 0:
            # distinguish "unknown" from "known but empty"
            self._ck(self._lib.wf_get_array(self._h, name.encode(), C.c_void_p(1), 0))
        out = np.empty(nbytes // np.dtype(dt).itemsize, dtype=dt)
        self._ck(self._lib.wf_get_array(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), nbytes))
        return out

    def set(self, name: str, arr):
        dt = _INT_ARRAYS.get(name, np.float64)
        arr = np.ascontiguousarray(arr, dtype=dt)
        self._ck(self._lib.wf_set_array(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def device_ptr(self, name: str):
        pitch = C.c_size_t()
        p = self._lib.wf_device_ptr(self._h, name.encode(), C.byref(pitch))
        return p, pitch.value
