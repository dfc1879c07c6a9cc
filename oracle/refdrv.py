"""TEST INFRASTRUCTURE (oracle) — ctypes driver for the two CPU checkers.

* ``RefDomain``    -> oracle/_ref/libwf_ref.so : the UNMODIFIED reference CPU sources driven
                      member-by-member (oracle/ref_harness.cpp).  Exists only where it was built
                      from /root/reference (it travels to the GPU box as a prebuilt .so).
* ``OracleDomain`` -> oracle/_build/libwf_oracle.so : the plain-C restatement (oracle/wf_oracle.c).

Both expose the same Python surface (the Domain_d call sequence) so a parity test reads the same
whichever checker it uses.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libwf_ref.so")
PORT_SO = os.path.join(HERE, "_build", "libwf_oracle.so")

_DBL = ("x v a u u_dt prev_a m_fi m_fe m_mdiag m_voln p_node m_dH_detJ_dx m_dH_detJ_dy m_dH_detJ_dz "
        "m_detJ vol vol_0 rho rho_0 p pl_strain sigma_y m_radius m_str_rate m_rot_rate m_sigma m_tau "
        "m_eps m_f_elem m_f_elem_hg m_hg_q m_voln_0 m_Jn bcx_val bcy_val bcz_val m_elem_length "
        "T m_dTedt m_q_plheat q_cont_conv contforce ut_prev node_area m_elem_area trimesh.node trimesh.node_v trimesh.normal trimesh.pplane").split()
_INT = ("m_nodel m_nodel_loc m_nodel_offset m_nodel_count m_mesh_in_contact trimesh.elnode "
        "trimesh.ele_mesh_id").split()
_UINT = ["m_elnod"]
_BYTE = ["ext_nodes"]


def _dtype_of(name):
    if name in _DBL:
        return np.float64
    if name in _UINT:
        return np.uint32
    if name in _BYTE:
        return np.uint8
    return np.int32

HOLLOMON = 1
BILINEAR = 0
JOHNSON_COOK = 2
GMT = 3

STAB_FIELDS = ("alpha_free alpha_contact hg_coeff_free hg_coeff_contact av_coeff_div av_coeff_bulk "
               "log_factor pspg_scale p_pspg_bulkfac J_min hg_visc hg_stiff").split()


def build(target: str = "all", quiet: bool = True) -> None:
    """Build the checkers (building the checker is not using it)."""
    targets = []
    if target in ("all", "port"):
        targets.append("port")
    if target in ("all", "ref") and os.path.isdir("/root/reference"):
        targets.append("ref")
    for t in targets:
        subprocess.run(["make", "-C", HERE, t], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_port() -> bool:
    return os.path.exists(PORT_SO)


class _Base:
    _prefix = ""
    _lib = None

    @classmethod
    def _load(cls, path):
        lib = C.CDLL(path)
        p = cls._prefix
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        sig = {
            "new": (vp, []), "free": (None, [vp]),
            "set_domtype": (None, [vp, C.c_int, C.c_int]),
            "box": (None, [vp, dp, dp, C.c_double, C.c_int]),
            "set_mesh": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, ip]),
            "set_material": (None, [vp] + [C.c_double] * 3 + [C.c_int] + [C.c_double] * 3),
            "set_material_ext": (None, [vp] + [C.c_double] * 3 + [C.c_int, C.c_double, dp, C.c_double]),
            "set_max_edot": (None, [vp, C.c_double]),
            "thermal_on": (None, [vp] + [C.c_double] * 5),
            "set_contact_heat": (None, [vp, C.c_double, C.c_double]),
            "set_stab": (None, [vp, dp]),
            "set_options": (None, [vp, C.c_int, C.c_double, C.c_double, C.c_double]),
            "add_bc": (None, [vp, C.c_int, C.c_int, C.c_double]),
            "add_plane": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp]),
            "set_trimesh": (None, [vp, C.c_int, C.c_int, C.c_int, dp, dp, ip, dp, ip]),
            "contact_on": (None, [vp] + [C.c_double] * 4),
            "trimesh_counts": (None, [vp, ip]),
            "allocate_bcs": (None, [vp]),
            "init": (None, [vp, C.c_double]),
            "step": (None, [vp, C.c_int]),
            "time_steps": (C.c_double, [vp, C.c_int]),
            "call": (C.c_int, [vp, C.c_char_p, C.c_double]),
            "get": (C.c_long, [vp, C.c_char_p, vp, C.c_long]),
            "set": (C.c_long, [vp, C.c_char_p, vp, C.c_long]),
            "info": (None, [vp, ip]),
            "consts": (None, [vp, dp]),
            "energies": (None, [vp, dp, dp]),
            "set_threads": (None, [C.c_int]),
            "max_threads": (C.c_int, []),
        }
        for name, (res, args) in sig.items():
            f = getattr(lib, p + name)
            f.restype, f.argtypes = res, args
        return lib

    def __init__(self):
        self.h = self._f("new")()

    def _f(self, name):
        return getattr(type(self)._lib, self._prefix + name)

    def close(self):
        if self.h:
            self._f("free")(self.h)
            self.h = None

    # ---- setup (same order as src/explicit/main.C) -------------------------------------------
    def set_domtype(self, domtype: int, vol_weight: bool = False):
        self._f("set_domtype")(self.h, domtype, int(vol_weight))

    def box(self, V, L, r, tritet=False):
        V = (C.c_double * 3)(*V)
        L = (C.c_double * 3)(*L)
        self._f("box")(self.h, V, L, r, int(tritet))

    def set_mesh(self, dim, k, x, elnod):
        x = np.ascontiguousarray(x, dtype=np.float64)
        el = np.ascontiguousarray(elnod, dtype=np.int32)
        nn, ne = x.size // dim, el.size // k
        self._f("set_mesh")(self.h, dim, k, nn, ne, x.ctypes.data_as(C.POINTER(C.c_double)),
                            el.ctypes.data_as(C.POINTER(C.c_int)))

    def set_material(self, E, nu, rho0, model=BILINEAR, sy0=1.0e10, K=0.0, m=1.0):
        self._f("set_material")(self.h, E, nu, rho0, model, sy0, K, m)

    def set_material_ext(self, E, nu, rho0, model, sy0, params, temp=20.0, max_edot=None):
        """Johnson-Cook (params = A B n C eps_0 m T_m T_t) or GMT (n1 n2 C1 C2 m1 m2 I1 I2 e_min e_max er_min er_max
        T_min T_max) through the public Material_ fields the free functions of Material.cuh:377-483 read."""
        q = (C.c_double * 14)(*([float(v) for v in params] + [0.0] * (14 - len(params))))
        self._f("set_material_ext")(self.h, E, nu, rho0, int(model), sy0, q, float(temp))
        if max_edot is not None:
            self._f("set_max_edot")(self.h, float(max_edot))

    def thermal_on(self, k_T, cp_T, exp_T=0.0, plheatfrac=0.9, T0=20.0):
        """setThermalOn + setTemp(T0) + thermalCond / thermalHeatCap / thermalExp + plHeatFrac (main.C:218, 436-441, 567-570)."""
        self._f("thermal_on")(self.h, float(k_T), float(cp_T), float(exp_T), float(plheatfrac), float(T0))

    def set_contact_heat(self, heat_cond, T_const):
        """heatCondCoeff / dieTemp of the rigid surfaces (main.C:718-719); call after contact_on."""
        self._f("set_contact_heat")(self.h, float(heat_cond), float(T_const))

    def set_stab(self, **kw):
        vals = [float(kw.get(k, 0.0)) for k in STAB_FIELDS]
        self._f("set_stab")(self.h, (C.c_double * 12)(*vals))

    def set_options(self, press=0, av_alpha=0.0, av_beta=0.0, hexa_hg=0.0):
        self._f("set_options")(self.h, press, av_alpha, av_beta, hexa_hg)

    def add_bc(self, node, dim, val):
        self._f("add_bc")(self.h, int(node), int(dim), float(val))

    def allocate_bcs(self):
        self._f("allocate_bcs")(self.h)

    # ---- contact with rigid surfaces (src/explicit/main.C:636-848) ------------------------------
    def add_plane(self, dimension, mesh_id, axis, positaxisorent, p1, p2, dens, vel=(0.0, 0.0, 0.0)):
        """TriMesh_d::AxisPlaneMesh (+ AddMesh for every body after the first) with node velocity ``vel``."""
        a3 = lambda q: (C.c_double * 3)(*[float(t) for t in q])
        self._f("add_plane")(self.h, int(dimension), int(mesh_id), int(axis), int(bool(positaxisorent)), a3(p1),
                             a3(p2), int(dens), a3(vel))

    def set_trimesh(self, dimension, node, node_v, elnode, normal, mesh_id):
        node = np.ascontiguousarray(node, dtype=np.float64)
        node_v = np.ascontiguousarray(node_v, dtype=np.float64)
        elnode = np.ascontiguousarray(elnode, dtype=np.int32)
        normal = np.ascontiguousarray(normal, dtype=np.float64)
        mesh_id = np.ascontiguousarray(mesh_id, dtype=np.int32)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        self._f("set_trimesh")(self.h, int(dimension), node.size // 3, mesh_id.size, node.ctypes.data_as(dp),
                               node_v.ctypes.data_as(dp), elnode.ctypes.data_as(ip), normal.ctypes.data_as(dp),
                               mesh_id.ctypes.data_as(ip))

    def contact_on(self, mu_sta=0.0, mu_dyn=0.0, penalty_factor=-1.0, end_time=1.0):
        """friction + penalty factor (main.C:716-725), CalcSpheres + setContactOn (:842-847), SetEndTime."""
        self._f("contact_on")(self.h, float(mu_sta), float(mu_dyn), float(penalty_factor), float(end_time))

    def trimesh_counts(self):
        out = (C.c_int * 3)()
        self._f("trimesh_counts")(self.h, out)
        return dict(zip("dimension nodecount elemcount".split(), list(out)))

    def init(self, dt):
        self._f("init")(self.h, dt)

    def step(self, n=1):
        self._f("step")(self.h, n)

    def time_steps(self, n):
        return self._f("time_steps")(self.h, n)

    def call(self, fn, arg=0.0):
        rc = self._f("call")(self.h, fn.encode(), float(arg))
        if rc != 0:
            raise KeyError(fn)

    @classmethod
    def _ensure(cls):
        if cls._lib is None:
            cls().close()

    @classmethod
    def set_threads(cls, n):
        cls._ensure()
        getattr(cls._lib, cls._prefix + "set_threads")(int(n))

    @classmethod
    def max_threads(cls):
        cls._ensure()
        return getattr(cls._lib, cls._prefix + "max_threads")()

    # ---- state ---------------------------------------------------------------------------------
    def info(self):
        out = (C.c_int * 8)()
        self._f("info")(self.h, out)
        keys = "dim nodxelem n_nodes n_elems bcx bcy bcz domtype".split()
        return dict(zip(keys, list(out)))

    def consts(self):
        out = (C.c_double * 7)()
        self._f("consts")(self.h, out)
        return dict(zip("alpha beta gamma dt time min_length min_height".split(), list(out)))

    def energies(self):
        ek, de = C.c_double(), C.c_double()
        self._f("energies")(self.h, C.byref(ek), C.byref(de))
        return ek.value, de.value

    def get(self, name):
        dt = _dtype_of(name)
        need = self._f("get")(self.h, name.encode(), None, 0)
        if need == -1:
            raise KeyError(name)
        nbytes = -need if need < 0 else need
        out = np.empty(nbytes // np.dtype(dt).itemsize, dtype=dt)
        if nbytes:
            got = self._f("get")(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), nbytes)
            assert got == nbytes, (name, got, nbytes)
        return out

    def set(self, name, arr):
        dt = _dtype_of(name)
        arr = np.ascontiguousarray(arr, dtype=dt)
        rc = self._f("set")(self.h, name.encode(), arr.ctypes.data_as(C.c_void_p), arr.nbytes)
        if rc != arr.nbytes:
            raise ValueError(f"set({name}) size mismatch")


class RefDomain(_Base):
    """The unmodified reference, compiled from /root/reference (oracle/_ref)."""
    _prefix = "wfref_"

    def __init__(self):
        if RefDomain._lib is None:
            if not have_ref():
                raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
            RefDomain._lib = self._load(REF_SO)
        super().__init__()

    @classmethod
    def from_deck(cls, deck_path):
        """Set a domain up by running the reference's own front-end (src/explicit/main.C, compiled unmodified into the
        harness) on a JSON deck; the process changes into the deck's directory for the call because main.C resolves
        `fileName` and its `.out` log relative to the current directory.  Returns (domain, dt, end_time); the domain
        still needs init(dt)."""
        import os
        cls._ensure()
        lib = cls._lib
        lib.wfref_load_deck.restype = C.c_void_p
        lib.wfref_load_deck.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
        out = (C.c_double * 2)()
        cwd = os.getcwd()
        os.chdir(os.path.dirname(os.path.abspath(deck_path)))
        try:
            h = lib.wfref_load_deck(os.path.basename(deck_path).encode(), out)
        finally:
            os.chdir(cwd)
        if not h:
            raise RuntimeError("the reference front-end did not reach SolveChungHulbert for " + deck_path)
        self = cls.__new__(cls)
        self.h = h
        return self, out[0], out[1]


class OracleDomain(_Base):
    """The plain-C restatement (oracle/wf_oracle.c)."""
    _prefix = "wfo_"

    def __init__(self):
        if OracleDomain._lib is None:
            if not have_port():
                build("port")
            OracleDomain._lib = self._load(PORT_SO)
        super().__init__()
